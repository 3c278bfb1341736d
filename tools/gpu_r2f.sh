#!/bin/bash
TAG=${1:-r2f}
O=gpurun_out/$TAG
mkdir -p $O
bash tools/gpu_km.sh 2>&1 | grep "create\|on the device\|== "
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/pytest.log
tail -5 $O/pytest.log
for w in C3 C2 C4; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err; echo "bench $w exit $?"
  grep -o '"ms_per_step": [0-9.]*\|"seconds_per_call": [0-9.]*\|"msa_identical_to_reference": [a-z]*\|"create_seconds": [0-9.]*' $O/bench_$w.json | tr '\n' ' '; echo
done
