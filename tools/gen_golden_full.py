#!/usr/bin/env python
"""Full-size golden fixtures: run the UNMODIFIED reference (oracle/_ref, built from
/root/reference) ONCE on the BASELINE.json configurations at their stated sizes and store, per
configuration, under tests/golden/full_<tag>.npz:

  msa_sha256     SHA-256 of the aligned rows in input order, joined by "\n" (kalign_b200.msa_sha256)
  alnlen, n      alignment length, number of sequences
  tasks          the guide tree of build_tree_kmeans (lib/src/bisectingKmeans.c:177): (a, b, c) x N-1
  seq_distances  msa->seq_distances (bisectingKmeans.c:247-256)
  gaps_sha256    SHA-256 of all gaps[] arrays in sorted order (int32, concatenated)
  confidence     task->confidence of every task (aln_run.c:390-394)
  times          the reference's stage times on this container (dist+tree, anchors, tree alignment, total)

Only the hash of the MSA is stored (a C3 alignment is 10 000 x ~4 000 characters); the inputs are
regenerated from the seeded generator (kalign_b200/synth.py).

  python tools/gen_golden_full.py C2 C4 C3 C5r24 T3 ...      (minutes to ~15 min each on 8 cores)
"""
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import kbind  # noqa: E402
from kalign_b200 import synth  # noqa: E402
from kalign_b200.synth import msa_sha256  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# tag -> (synth config, n or None, type, consistency anchors, stop_after)
#   type: 0 dna, 2 rna, 8 protein (lib/include/kalign/kalign.h:18-26); consistency 0 = --fast
SPECS = {
    "C2": ("C2", None, 8, 5, 0),
    "C3": ("C3", None, 2, 5, 0),
    "C4": ("C4", None, 8, 0, 0),
    "C5r24": ("C5", 24, 0, 5, 0),
    "C5r8": ("C5", 8, 0, 5, 0),
    "C3r2000": ("C3", 2000, 2, 5, 0),
    "C2fast": ("C2", None, 8, 0, 0),
    # guide tree only (stop after build_tree_kmeans)
    "T3": ("C3", None, 2, 0, 1),
    "T4": ("C4", None, 8, 0, 1),
}


def run(tag, threads):
    cfg, n, type_, cons, stop = SPECS[tag]
    seqs = synth.config(cfg, n)
    t0 = time.time()
    r = kbind.RefRun(seqs, n_threads=threads, type_=type_, consistency=cons, weight=2.0, stop_after=stop)
    wall = time.time() - t0
    rec = {"config": cfg, "n": len(seqs), "type": type_, "consistency": cons, "stop_after": stop,
           "tasks": r.tasks(), "seq_distances": r.seq_distances(), "rank": r.rank, "lens": r.lens,
           "wall_s": wall, "threads": threads}
    tm = r.times()
    rec["times"] = np.array([tm["dist_tree"], tm["anchor"], tm["tree_aln"], tm["total"]])
    if stop == 0:
        rows = r.aligned()
        rec["msa_sha256"] = msa_sha256(rows)
        rec["alnlen"] = len(rows[0])
        h = hashlib.sha256()
        for i in range(r.n):
            h.update(r.gaps(i).astype(np.int32).tobytes())
        rec["gaps_sha256"] = h.hexdigest()
        rec["confidence"] = r.task_confidence()
    r.close()
    np.savez_compressed(os.path.join(OUT, "full_%s.npz" % tag), **rec)
    print(tag, "n=%d" % len(seqs), "wall %.1f s" % wall, rec.get("msa_sha256", "")[:16], flush=True)


if __name__ == "__main__":
    threads = int(os.environ.get("KB_REF_THREADS", "8"))
    for tag in sys.argv[1:]:
        run(tag, threads)
