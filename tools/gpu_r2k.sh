#!/bin/bash
O=gpurun_out/r2k; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_pairs.py tests/test_gpu_pipeline.py tests/test_gpu_edge.py tests/test_golden.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest exit $?"; tail -4 $O/pytest.log
for v in 128 64 32 16; do
  KB200_SMALL_ROWS_SS=$v KB200_TRACE=1 timeout 600 python bench.py --workload C3 --steps 1 --warmup 1 --no-cpu-baseline > $O/c3t_$v.json 2> $O/c3t_$v.err
  KB200_SMALL_ROWS_SS=$v timeout 600 python bench.py --workload C3 --steps 2 --warmup 2 --no-cpu-baseline > $O/c3_$v.json 2> $O/c3_$v.err
  echo -n "small_rows_ss=$v: "; python - <<PY
import json
d=json.load(open("$O/c3_$v.json")); print("ms/step %.1f small %.1f ms sweep %.1f ms identical %s" % (d["ms_per_step"], 1e3*d["roofline"]["small_box_kernel_seconds_per_step"], 1e3*d["roofline"]["kernel_seconds_per_step"], d["msa_identical_to_reference"]))
PY
  grep "jobs=49995" $O/c3t_$v.err | tail -8 | cut -c14-110
done
