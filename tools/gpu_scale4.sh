#!/bin/bash
# 4-GPU box: multi-GPU parity test, strong-scaling bench at N = 1 / 2 / 4, per-level trace of rank 0 at N = 4
TAG=${1:-scale4}
O=gpurun_out/$TAG
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > $O/pytest_multi.log 2>&1; echo "pytest multi exit $?"; tail -3 $O/pytest_multi.log
timeout 600 python bench.py --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err
for n in 2 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_n$n.json 2> $O/bench_n$n.err
  echo "N=$n exit $?"
done
for n in 1 2 4; do grep -o '"ms_per_step": [0-9.]*\|"seconds_per_call": [0-9.]*\|"msa_identical_to_reference": [a-z]*\|"msa_identical_on_all_ranks": [a-z]*' $O/bench_n$n.json | tr '\n' ' '; echo; done
KB200_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 1 --warmup 1 --no-cpu-baseline > $O/trace_n4.json 2> $O/trace_n4.err
grep "tree level" $O/trace_n4.err | tail -88 | awk 'NR%4==1' | tail -22
KB200_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/trace_n1.json 2> $O/trace_n1.err
python tools/trace_sum.py $O/trace_n1.err
