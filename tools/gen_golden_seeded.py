"""Golden vectors for the seeded / ensemble entry points from the UNMODIFIED reference (oracle/_ref):
noise factors of build_tree_kmeans_noisy (lib/src/tlrng.c generator), resolve_run_params (lib/src/ensemble.c:55,
through oracle/ref_ensemble_params.c) and alignments of kalign_run_seeded (lib/src/aln_wrap.c:133) with noisy guide
trees, scaled gap penalties, dist_scale and use_seq_weights -> tests/golden/seeded.npz.
Run once in the build container:  python tools/gen_golden_seeded.py"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import kbind  # noqa: E402
from kalign_b200 import synth  # noqa: E402

NOISE = [(43, 0.2), (1, 0.35), (2 ** 40 + 7, 0.15), (99, 5.0)]
BASES = [(5.5, 2.0, 1.0), (217.0, 39.4, 292.6)]


def families():
    return {
        "protein": (synth.family(60, 100, synth.PROTEIN, seed=61), 8),
        "rna": (synth.family(48, 180, synth.RNA, seed=62), 2),
        "dna": (synth.family(150, 120, synth.DNA, seed=63, sub=0.08, ins=0.01, dele=0.01), 0),
    }


# name -> (family, kwargs of kalign_run_seeded)
SEEDED = {
    "protein_noise": ("protein", dict(tree_seed=43, tree_noise=0.2)),
    "protein_noise_gaps": ("protein", dict(tree_seed=44, tree_noise=0.35, gpo=3.0, gpe=3.0, tgpe=0.5, consistency=5)),
    "rna_noise_consistency": ("rna", dict(tree_seed=7, tree_noise=0.3, consistency=5)),
    "dna_noise_many_bisections": ("dna", dict(tree_seed=2 ** 33, tree_noise=0.25)),
    "rna_seed_without_noise": ("rna", dict(tree_seed=5, tree_noise=0.0)),
    "protein_dist_scale_weights": ("protein", dict(tree_seed=9, tree_noise=0.2, dist_scale=0.5, use_seq_weights=1.0, vsm_amax=1.0)),
}
ENSEMBLE = {"protein": (5, 42), "rna": (4, 1234)}          # family -> (n_runs, seed); consistency 0


def main():
    rec = {}
    for i, (seed, sigma) in enumerate(NOISE):
        f = kbind.ref_tree_noise(seed, sigma, 100000)
        rec["noise%d_head" % i] = f[:64]
        rec["noise%d_sha" % i] = hashlib.sha256(f.tobytes()).hexdigest()
    rp = []
    for base in BASES:
        for k in range(26):
            rp.append(kbind.ref_resolve_run_params(*base, k, 42))
    rec["run_params"] = np.array(rp, dtype=np.float64)
    fam = families()
    for name, (fk, kw) in SEEDED.items():
        seqs, type_ = fam[fk]
        rec["rows_" + name] = np.array(kbind.ref_run_seeded(seqs, n_threads=2, type_=type_, **kw))
    for fk, (n_runs, seed) in ENSEMBLE.items():
        seqs, type_ = fam[fk]
        run = kbind.RefRun(seqs, n_threads=2, type_=type_, consistency=0, weight=2.0)
        _, gp = run.params()
        run.close()
        rec["ens_base_" + fk] = np.array(gp[:3], dtype=np.float32)
        for k in range(n_runs):
            g, e, t, ts, nz = kbind.ref_resolve_run_params(float(gp[0]), float(gp[1]), float(gp[2]), k, seed)
            rec["ens_%s_%d" % (fk, k)] = np.array(kbind.ref_run_seeded(seqs, n_threads=2, type_=type_, gpo=g, gpe=e, tgpe=t, tree_seed=ts,
                                                                      tree_noise=nz, use_seq_weights=0.0))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "seeded.npz"), **rec)
    print("seeded:", len(SEEDED), "alignments,", sum(v[0] for v in ENSEMBLE.values()), "ensemble runs")
    # the noisy trees must actually differ from the plain ones, or the fixtures pin nothing
    for name, (fk, kw) in SEEDED.items():
        if kw.get("tree_noise", 0) > 0:
            seqs, type_ = fam[fk]
            kw0 = dict(kw, tree_seed=0, tree_noise=0.0)
            plain = kbind.ref_run_seeded(seqs, n_threads=2, type_=type_, **kw0)
            print(name, "differs from the plain tree:", plain != [str(x) for x in rec["rows_" + name]])


if __name__ == "__main__":
    main()
