#!/bin/bash
# round-2 final check: full GPU parity suite, smoke(), per-workload trace + bench line
O=gpurun_out/r2l; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?"; tail -1 $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?"; tail -5 $O/pytest.log
for w in C3 C4 C2; do
  KB200_TRACE=1 timeout 600 python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline > $O/trace_$w.json 2> $O/trace_$w.err
  grep "guide tree" $O/trace_$w.err | tail -3
  timeout 600 python bench.py --workload $w --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err; echo "bench $w exit $?"
  grep -o '"ms_per_step": [0-9.]*\|"seconds_per_call": [0-9.]*\|"msa_identical_to_reference": [a-z]*' $O/bench_$w.json | tr '\n' ' '; echo
done
python tools/trace_sum.py $O/trace_C3.err
