#!/bin/bash
# One gpurun call: GPU parity tests, bench line, per-level trace, ncu launch list + full capture of the sweep kernel.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [tag]'
TAG=${1:-run}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/pytest.log
tail -3 $O/pytest.log
timeout 600 python bench.py > $O/bench_c3.json 2> $O/bench_c3.err; echo "bench exit $?"
cat $O/bench_c3.json
if [ -z "$KB200_CHECK_SKIP_REF" ]; then timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_c3_reference.json 2> $O/bench_c3_reference.err; echo "reference arm exit $?"; fi
KB200_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/trace_bench.json 2> $O/trace_c3.log
python tools/trace_sum.py $O/trace_c3.log > $O/trace_sum.txt; cat $O/trace_sum.txt
for w in ${KB200_CHECK_SHAPES:-C2 C4}; do
  timeout 400 python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err; echo "bench $w exit $?"
  grep -o '"ms_per_step": [0-9.]*\|"seconds_per_call": [0-9.]*' $O/bench_$w.json | tr '\n' ' '; echo
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches.csv \
    python bench.py --workload C3 --n 1000 --steps 1 --warmup 0 --no-cpu-baseline > $O/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kb_sweep_kernel -c 3 -f -o $O/sweep_full \
    python bench.py --workload C3 --n 1000 --steps 1 --warmup 0 --no-cpu-baseline > $O/ncu_full.log 2>&1
ls -la $O
