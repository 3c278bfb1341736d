#!/bin/bash
# round-2 ncu evidence, sweep kernel families (demangled names so that the template argument can be matched)
TAG=${1:-r02ncu}
O=gpurun_out/$TAG
mkdir -p $O
B="python bench.py --workload C3 --steps 1 --warmup 0 --no-cpu-baseline"
NCU="timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
$NCU -k "regex:kb_sweep_kernel<\\(int\\)0>" -c 2 -o $O/sweep_none $B > $O/ncu_sweep_none.log 2>&1
$NCU -k "regex:kb_sweep_kernel<\\(int\\)1>" -c 1 -o $O/sweep_sparse_l1 $B > $O/ncu_sweep_sparse.log 2>&1
# level 2 (profile-sequence / profile-profile) round 0: every level enqueues 14 rounds
$NCU -k "regex:kb_sweep_kernel<\\(int\\)1>" --launch-skip 14 -c 1 -o $O/sweep_sparse_l2 $B > $O/ncu_sweep_sparse2.log 2>&1
# the root task (level 22 of 22), round 0: the lone-warp regime
$NCU -k "regex:kb_sweep_kernel<\\(int\\)1>" --launch-skip 294 -c 1 -o $O/sweep_sparse_root $B > $O/ncu_sweep_root.log 2>&1
ls -la $O | grep sweep
