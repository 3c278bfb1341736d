#!/bin/bash
TAG=${1:-r2g}
O=gpurun_out/$TAG
mkdir -p $O
KB200_TRACE=1 timeout 200 python tools/check_long.py 4 6000 5 > $O/pp4.out 2> $O/pp4.err
cat $O/pp4.out | head -3; grep "jobs=1 round=[0-3] " $O/pp4.err | tail -4
timeout 600 python bench.py --workload C3 --no-cpu-baseline > $O/bench_C3.json 2> $O/bench_C3.err; echo "bench C3 exit $?"
grep -o '"ms_per_step": [0-9.]*\|"seconds_per_call": [0-9.]*\|"msa_identical_to_reference": [a-z]*' $O/bench_C3.json | tr '\n' ' '; echo
# BASELINE config 5 at full size: 1000 x 30 kb genomes, --type dna, default mode
timeout 1200 python bench.py --workload C5 --steps 1 --warmup 0 --no-cpu-baseline > $O/bench_C5.json 2> $O/bench_C5.err; echo "bench C5 exit $?"
python - <<PY
import json
d=json.load(open("$O/bench_C5.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","cells_per_step")}, d["e2e"], d["roofline"]["frac"], d["roofline"]["kernel_seconds_per_step"])
PY
tail -3 $O/bench_C5.err
