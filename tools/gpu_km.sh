#!/bin/bash
O=gpurun_out/km; mkdir -p $O
for w in C3 C4; do
KB200_TRACE=1 timeout 300 python - > $O/$w.out 2> $O/$w.err <<PY
import time
from kalign_b200 import _lib, synth
import bench
cfg, type_, K, _ = bench.WORKLOADS["$w"]
seqs = synth.config(cfg)
ctx = _lib.Context(0)
for it in range(2):
    t0 = time.perf_counter()
    m = _lib.Msa(ctx, seqs, n_threads=16, type_=type_, consistency=0)
    print("create %.3f s" % (time.perf_counter() - t0))
    m.close()
PY
echo "== $w"; cat $O/$w.out; grep "k-means\|guide tree" $O/$w.err | tail -22
done
