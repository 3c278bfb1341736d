#!/usr/bin/env python
"""Host-stage timing of the public call kb200_kalign on a BASELINE shape (KB200_TRACE=1 lines of
the host stages only).  usage: python tools/e2e_trace.py [C3] [n]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["KB200_TRACE"] = "1"
from kalign_b200 import _lib, synth
import bench

w = sys.argv[1] if len(sys.argv) > 1 else "C3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else None
cfg, type_, K, label = bench.WORKLOADS[w]
seqs = synth.config(cfg, n)
ctx = _lib.Context(0)
threads = bench.effective_cpus()
for it in range(3):
    t0 = time.perf_counter()
    rows = ctx.kalign(seqs, n_threads=threads, type_=type_, consistency=K, weight=2.0)
    sys.stderr.write("[e2e] call %d: %.1f ms wall (%d host threads)\n" % (it, 1e3 * (time.perf_counter() - t0), threads))
