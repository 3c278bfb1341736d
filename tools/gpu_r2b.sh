#!/bin/bash
# GPU parity (incl. full-size hashes), bench C3 / C2 / C4, per-level trace
TAG=${1:-r2b}
O=gpurun_out/$TAG
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/pytest.log
tail -8 $O/pytest.log
timeout 600 python bench.py --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("$O/bench_c3.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","msa_identical_to_reference")}, d["e2e"]["seconds_per_call"], d["roofline"]["kernel_seconds_per_step"], d["roofline"]["small_box_kernel_seconds_per_step"])
PY
tail -3 $O/bench_c3.err
KB200_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/trace_bench.json 2> $O/trace_c3.log
python tools/trace_sum.py $O/trace_c3.log > $O/trace_sum.txt; cat $O/trace_sum.txt
grep "tree level" $O/trace_c3.log | tail -22
grep "jobs=1 round" $O/trace_c3.log | tail -10
for w in C2 C4; do
  timeout 400 python bench.py --workload $w --steps 3 --warmup 2 --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err; echo "bench $w exit $?"
  grep -o '"ms_per_step": [0-9.]*\|"seconds_per_call": [0-9.]*\|"msa_identical_to_reference": [a-z]*' $O/bench_$w.json | tr '\n' ' '; echo
done
