#!/bin/bash
# thin-regime rows per lane: K = 1 / 2 / 4 on the top tree levels of C3 and on a lone 6000 x 6000 profile pair
TAG=${1:-thinK}
O=gpurun_out/$TAG
mkdir -p $O
for k in 1 2 4; do
  KB200_THIN_K=$k KB200_TRACE=1 timeout 200 python tools/check_long.py 4 6000 5 > $O/pp4_k$k.out 2> $O/pp4_k$k.err
  echo "== pp4 K=$k"; cat $O/pp4_k$k.out | head -3; grep "jobs=1 round=[0-3] " $O/pp4_k$k.err | tail -4
  KB200_THIN_K=$k KB200_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/c3_k$k.json 2> $O/c3_k$k.err
  echo "== C3 K=$k"; python tools/trace_sum.py $O/c3_k$k.err; grep "tree level \(8\|12\|16\|21\):" $O/c3_k$k.err | tail -4
  KB200_THIN_K=$k timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > $O/c3b_k$k.json 2> $O/c3b_k$k.err
  grep -o '"ms_per_step": [0-9.]*\|"msa_identical_to_reference": [a-z]*' $O/c3b_k$k.json | tr '\n' ' '; echo
done
