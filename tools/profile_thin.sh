#!/bin/bash
# ncu source-level captures of the LONE-WARP (thin strip) regime of kb_sweep_kernel:
#   thin_ss : one seq-seq job 8000 x 8000 (2 sequences, --fast), round 0
#   thin_pp : the profile-profile root task of a 4-sequence family (6000 nt), round 0
# usage: gpurun -- 'bash tools/profile_thin.sh <tag>'
TAG=${1:-thin}
O=gpurun_out/$TAG
mkdir -p $O
NCU="ncu --set full --import-source on --clock-control none -f"
timeout 300 $NCU -k regex:kb_sweep_kernel -c 1 -o $O/thin_ss python tools/check_long.py 2 8000 0 > $O/thin_ss.log 2>&1
KB200_TRACE=1 timeout 120 python tools/check_long.py 4 6000 0 > $O/trace4.out 2> $O/trace4.err
skip=$(grep "round=" $O/trace4.err | grep -n "jobs=1 round=0" | head -1 | cut -d: -f1)
skip=$((skip-1))
echo "pp root task: skipping $skip sweep launches" | tee $O/skip.txt
timeout 300 $NCU -k regex:kb_sweep_kernel --launch-skip $skip -c 1 -o $O/thin_pp python tools/check_long.py 4 6000 0 > $O/thin_pp.log 2>&1
grep "round=\|small" $O/trace4.err | tail -14
ls -la $O
