#!/usr/bin/env python
"""Timing + sanity of the public call on the BASELINE shapes that are not the bench line:
C4 (100 000 x 300 aa, --fast) with host-stage trace, and a C5-shaped family (30 kb DNA genomes,
--type dna, default mode) at a reduced sequence count.  usage: python tools/shape_check.py [n_c5]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["KB200_TRACE"] = "1"
from kalign_b200 import _lib, synth
import bench

ctx = _lib.Context(0)
threads = bench.effective_cpus()
n5 = int(sys.argv[1]) if len(sys.argv) > 1 else 100
for name, seqs, type_, K in (("C4", synth.config("C4"), 8, 0),
                             ("C5 shape, %d x 30 kb" % n5, synth.config("C5", n5), 0, 5)):
    for it in range(2):
        s0 = ctx.stats()
        t0 = time.perf_counter()
        rows = ctx.kalign(seqs, n_threads=threads, type_=type_, consistency=K, weight=2.0)
        dt = time.perf_counter() - t0
        s1 = ctx.stats()
        cells = s1["dp_cells"] - s0["dp_cells"]
        ok = all(r.replace("-", "") == s for r, s in zip(rows, seqs))
        sys.stderr.write("[shape] %s call %d: %.3f s wall, %.3g cells -> %.1f G cells/s e2e, alignment length %d, residues preserved %s\n"
                         % (name, it, dt, cells, cells / dt / 1e9, len(rows[0]), ok))
