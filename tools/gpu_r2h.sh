#!/bin/bash
TAG=${1:-r2h}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_pipeline.py tests/test_cmake_package.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest exit $?"; tail -4 $O/pytest.log
for v in 128 64 32 16; do
  for w in C4 C2; do
    KB200_SMALL_ROWS_PROF=$v timeout 600 python bench.py --workload $w --steps 3 --warmup 2 --no-cpu-baseline > $O/bench_${w}_$v.json 2> $O/bench_${w}_$v.err
    echo -n "small_rows_prof=$v $w: "; grep -o '"ms_per_step": [0-9.]*\|"msa_identical_to_reference": [a-z]*' $O/bench_${w}_$v.json | tr '\n' ' '; echo
  done
done
