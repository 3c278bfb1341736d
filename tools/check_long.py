#!/usr/bin/env python
"""One-off GPU check of the long-sequence regime (BASELINE config 5 shape, reduced N): a few
~30 kb DNA genomes, --type dna, fast mode; compares the GPU MSA with the reference byte for byte."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import kbind
from kalign_b200 import _lib, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
L = int(sys.argv[2]) if len(sys.argv) > 2 else 30000
K = int(sys.argv[3]) if len(sys.argv) > 3 else 0
seqs = synth.family(n, L, synth.DNA, seed=5, sub=0.01, ins=0.001, dele=0.001)
ctx = _lib.Context(0)
t0 = time.time(); got = ctx.kalign(seqs, n_threads=8, type_=0, consistency=K, weight=2.0); t1 = time.time()
st = ctx.stats()
print("gpu: %.2fs cells %.3g -> %.1f Gcells/s e2e" % (t1 - t0, st["dp_cells"], st["dp_cells"] / (t1 - t0) / 1e9))
t0 = time.time(); run = kbind.RefRun(seqs, n_threads=8, type_=0, consistency=K, weight=2.0); want = run.aligned(); t1 = time.time()
print("ref: %.2fs" % (t1 - t0), "identical:", got == want)
sys.exit(0 if got == want else 1)
