#!/bin/bash
TAG=${1:-r2d}
O=gpurun_out/$TAG
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/pytest.log
tail -12 $O/pytest.log
for w in C4 C3 C2; do
  KB200_TRACE=1 timeout 600 python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline > $O/trace_$w.json 2> $O/trace_$w.err
  echo "== $w"; grep "guide tree\|msa_create\|kalign:" $O/trace_$w.err | tail -8
  timeout 600 python bench.py --workload $w --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err; echo "bench $w exit $?"
  grep -o '"ms_per_step": [0-9.]*\|"seconds_per_call": [0-9.]*\|"msa_identical_to_reference": [a-z]*\|"create_seconds": [0-9.]*' $O/bench_$w.json | tr '\n' ' '; echo
done
python tools/trace_sum.py $O/trace_C3.err
grep "jobs=49995 round\|jobs=3964 round\|jobs=2193 round=0\|jobs=1 round=0" $O/trace_C3.err | tail -12
