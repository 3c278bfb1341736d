#!/bin/bash
# 8-GPU box: strong scaling of C3 at N = 8 (and 1 for the same box), per-level trace of one rank, full C5 at N = 8
TAG=${1:-scale8}
O=gpurun_out/$TAG
mkdir -p $O
run() { # n workload steps warmup tag extra-env
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 295$1$1 bench.py --gpus $1 --workload $2 --steps $3 --warmup $4 --no-cpu-baseline 2> $O/$5.err | grep '^{' > $O/$5.json
  echo "$5 exit $?"; grep -o '"ms_per_step": [0-9.]*\|"seconds_per_call": [0-9.]*\|"msa_identical_to_reference": [a-z]*\|"msa_identical_on_all_ranks": [a-z]*' $O/$5.json | tr '\n' ' '; echo
}
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > $O/pytest_multi.log 2>&1; echo "pytest multi exit $?"; tail -2 $O/pytest_multi.log
run 8 C3 3 3 c3_n8
run 4 C3 3 3 c3_n4
run 2 C3 3 3 c3_n2
run 8 C5 1 1 c5_n8
timeout 600 python bench.py --workload C3 --no-cpu-baseline 2> $O/c3_n1.err | grep '^{' > $O/c3_n1.json; grep -o '"ms_per_step": [0-9.]*\|"seconds_per_call": [0-9.]*' $O/c3_n1.json | tr '\n' ' '; echo
KB200_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 8 --workload C3 --steps 1 --warmup 1 --no-cpu-baseline > $O/trace_n8.json 2> $O/trace_n8.err
grep "tree level" $O/trace_n8.err | tail -176 | awk 'NR%8==1' | tail -22
