#!/bin/bash
# round-2 ncu evidence on the bench command (C3, one step): launch list + full captures per kernel family
TAG=${1:-r02ncu}
O=gpurun_out/$TAG
mkdir -p $O
B="python bench.py --workload C3 --steps 1 --warmup 0 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches.csv $B > $O/ncu_launch.log 2>&1
NCU="timeout 900 ncu --set full --clock-control none --import-source on -f"
$NCU -k "regex:kb_sweep_kernel<\(int\)0>" -c 2 -o $O/sweep_none $B > $O/ncu_sweep_none.log 2>&1
$NCU -k "regex:kb_sweep_kernel<\(int\)1>" -c 2 -o $O/sweep_sparse_l1 $B > $O/ncu_sweep_sparse.log 2>&1
# level 2 (profile-sequence / profile-profile) round 0 of the sparse family: level 1 enqueues 14 rounds
$NCU -k "regex:kb_sweep_kernel<\(int\)1>" --launch-skip 14 -c 1 -o $O/sweep_sparse_l2 $B > $O/ncu_sweep_sparse2.log 2>&1
$NCU -k "regex:kb_small_kernel" -c 2 -o $O/small $B > $O/ncu_small.log 2>&1
$NCU -k "regex:kb_meetup_kernel" -c 1 -o $O/meetup $B > $O/ncu_meetup.log 2>&1
$NCU -k "regex:kb_bpm_kernel" -c 1 -o $O/bpm $B > $O/ncu_bpm.log 2>&1
$NCU -k "regex:kb_bonus_votes" --launch-skip 20 -c 2 -o $O/votes $B > $O/ncu_votes.log 2>&1
ls -la $O
