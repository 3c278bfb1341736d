#!/bin/bash
TAG=${1:-r2e}
O=gpurun_out/$TAG
mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_full.py tests/test_gpu_pipeline.py tests/test_gpu_edge.py -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/pytest.log
tail -6 $O/pytest.log
for w in C4 C3; do
  KB200_TRACE=1 timeout 600 python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline > $O/trace_$w.json 2> $O/trace_$w.err
  echo "== $w"; grep "guide tree\|kalign:" $O/trace_$w.err | tail -6
  grep -o '"seconds_per_call": [0-9.]*\|"msa_identical_to_reference": [a-z]*\|"create_seconds": [0-9.]*' $O/trace_$w.json | tr '\n' ' '; echo
done
