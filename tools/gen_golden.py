#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REAL reference (oracle/_ref, built from /root/reference).
Run in the build container; the fixtures are committed so that the parity of the oracle and of
the GPU path can be checked where /root/reference does not exist (the GPU box).

  pairs.npz : seq-seq / profile-seq / profile-profile jobs -> raw Hirschberg path, top score,
              margin sum/count (aln_runner, lib/src/aln_controller.c:21)
  bpm.npz   : (text, pattern) pairs -> bpm_block (lib/src/bpm.c:356)
  msa_*.npz : whole kalign_run_seeded runs (lib/src/aln_wrap.c:133) on small synthetic families:
              input sequences, guide-tree tasks, seq_distances, anchor ids, gaps, aligned rows
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import kbind  # noqa: E402
from kalign_b200 import synth  # noqa: E402
from test_oracle_vs_ref import merged_profile, mutate, pfasum_like  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def gen_pairs():
    rng = np.random.default_rng(2024)
    rec = {}
    n = 0
    for name, A, gpo, gpe, tgpe in (("protein", 20, 7.0, 1.25, 1.0), ("rna", 4, 217.0, 39.4, 292.6)):
        subm = pfasum_like(rng) if name == "protein" else np.ascontiguousarray(
            rng.integers(-4, 6, size=(23, 23)).astype(np.float32) * 40.0)
        for la in (1, 7, 33, 64, 150):
            s1 = rng.integers(0, A, size=la).astype(np.uint8)
            s2 = mutate(rng, s1, A)
            if len(s2) < len(s1):
                s1, s2 = s2, s1
            soff = 0.5 if name == "protein" else 0.0
            p, st = kbind.ref_align(0, len(s1), len(s2), subm, gpo, gpe, tgpe, soff=soff, seq1=s1, seq2=s2)
            _, sc = kbind.ref_align(0, len(s1), len(s2), subm, gpo, gpe, tgpe, soff=soff, seq1=s1, seq2=s2, score_only=True)
            rec.update({"j%d_kind" % n: 0, "j%d_subm" % n: subm, "j%d_gp" % n: np.array([gpo, gpe, tgpe, soff], np.float32),
                        "j%d_seq1" % n: s1, "j%d_seq2" % n: s2, "j%d_path" % n: p,
                        "j%d_score" % n: np.float32(sc["score"]), "j%d_margin" % n: np.array([st["margin_sum"], st["margin_count"]], np.float64)})
            n += 1
        for L in (12, 45, 90):
            p1, l1, n1 = merged_profile(rng, A, L, subm, gpo, gpe, tgpe, depth=2)
            p2, l2, n2 = merged_profile(rng, A, L, subm, gpo, gpe, tgpe, depth=1)
            if l1 >= l2:
                p1, l1, n1, p2, l2, n2 = p2, l2, n2, p1, l1, n1
            a, b = p1.copy(), p2.copy()
            kbind.refh().refh_set_gap_penalties(a, l1, n2)
            kbind.refh().refh_set_gap_penalties(b, l2, n1)
            p, st = kbind.ref_align(2, l1, l2, subm, gpo, gpe, tgpe, prof1=a, prof2=b)
            _, sc = kbind.ref_align(2, l1, l2, subm, gpo, gpe, tgpe, prof1=a, prof2=b, score_only=True)
            rec.update({"j%d_kind" % n: 2, "j%d_subm" % n: subm, "j%d_gp" % n: np.array([gpo, gpe, tgpe, 0.0], np.float32),
                        "j%d_prof1" % n: a, "j%d_prof2" % n: b, "j%d_len" % n: np.array([l1, l2], np.int32), "j%d_path" % n: p,
                        "j%d_score" % n: np.float32(sc["score"]), "j%d_margin" % n: np.array([st["margin_sum"], st["margin_count"]], np.float64)})
            n += 1
            s = mutate(rng, rng.integers(0, A, size=L).astype(np.uint8), A)
            p, st = kbind.ref_align(1, l1, len(s), subm, gpo, gpe, tgpe, prof1=a, seq2=s, sip=n1)
            _, sc = kbind.ref_align(1, l1, len(s), subm, gpo, gpe, tgpe, prof1=a, seq2=s, sip=n1, score_only=True)
            rec.update({"j%d_kind" % n: 1, "j%d_subm" % n: subm, "j%d_gp" % n: np.array([gpo, gpe, tgpe, 0.0], np.float32),
                        "j%d_prof1" % n: a, "j%d_seq2" % n: s, "j%d_len" % n: np.array([l1, len(s), n1], np.int32), "j%d_path" % n: p,
                        "j%d_score" % n: np.float32(sc["score"]), "j%d_margin" % n: np.array([st["margin_sum"], st["margin_count"]], np.float64)})
            n += 1
    rec["njobs"] = n
    np.savez_compressed(os.path.join(OUT, "pairs.npz"), **rec)
    print("pairs:", n)


def gen_bpm():
    rng = np.random.default_rng(77)
    ref = kbind.ref()
    rec = {}
    n = 0
    for A in (4, 13):
        for m in (1, 5, 63, 64, 65, 200, 700, 1024, 1100):
            t = rng.integers(0, A, size=m + int(rng.integers(0, 200))).astype(np.uint8)
            p = mutate(rng, t, A, sub=0.25, indel=0.03)[:m]
            if len(p) > len(t):
                t, p = p, t
            rec["t%d" % n] = t
            rec["p%d" % n] = p
            rec["d%d" % n] = np.int32(ref.bpm_block(t, p, len(t), len(p)))
            n += 1
    rec["n"] = n
    np.savez_compressed(os.path.join(OUT, "bpm.npz"), **rec)
    print("bpm:", n)


def gen_msa(tag, seqs, type_, consistency):
    run = kbind.RefRun(seqs, n_threads=2, type_=type_, consistency=consistency, weight=2.0)
    rec = {"seqs": np.array(seqs), "type": type_, "consistency": consistency,
           "tasks": run.tasks(), "seq_distances": run.seq_distances(), "rank": run.rank, "lens": run.lens,
           "aligned": np.array(run.aligned())}
    subm, gp = run.params()
    rec["subm"] = subm
    rec["gp"] = gp
    if consistency:
        rec["anchor_ids"] = run.anchor_ids()
    for i in range(run.n):
        rec["gaps%d" % i] = run.gaps(i)
        rec["codes%d" % i] = run.codes(i)
    run.close()
    np.savez_compressed(os.path.join(OUT, "msa_%s.npz" % tag), **rec)
    print("msa", tag, len(seqs))


if __name__ == "__main__":
    gen_pairs()
    gen_bpm()
    gen_msa("c1_protein_default", synth.config("C1"), 8, 5)          # BASELINE config 1 shape
    gen_msa("protein_fast", synth.family(40, 80, synth.PROTEIN, seed=31), 8, 0)
    gen_msa("rna_default", synth.family(30, 150, synth.RNA, seed=32), 2, 5)
    gen_msa("dna_default", synth.family(20, 120, synth.DNA, seed=33, sub=0.05, ins=0.005, dele=0.005), 0, 5)
