#!/bin/bash
O=gpurun_out/r2j; mkdir -p $O
for v in 0 1 2; do
  KB200_SMALL_VARIANT=$v timeout 600 python bench.py --workload C3 --steps 2 --warmup 2 --no-cpu-baseline > $O/c3_$v.json 2> $O/c3_$v.err
  echo -n "small variant=$v: "; python - <<PY
import json
d=json.load(open("$O/c3_$v.json")); print("ms/step %.1f small %.1f ms sweep %.1f ms identical %s" % (d["ms_per_step"], 1e3*d["roofline"]["small_box_kernel_seconds_per_step"], 1e3*d["roofline"]["kernel_seconds_per_step"], d["msa_identical_to_reference"]))
PY
done
