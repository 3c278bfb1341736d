#!/bin/bash
# widening rows: smoke(), the GPU parity suite (new: test_apair, ensemble / realign CLI cases,
# kb200_kalign_file), the default bench line with its 'apair' object, FASTA I/O timing on the box's host cores
O=gpurun_out/r2m; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?"; tail -2 $O/smoke.log
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?"; tail -15 $O/pytest.log
timeout 600 python bench.py > $O/bench_C3.json 2> $O/bench_C3.err; echo "bench exit $?"; tail -3 $O/bench_C3.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2m/bench_C3.json"))
print("C3", d["ms_per_step"], d["e2e"]["seconds_per_call"], d["roofline"]["frac"], d["msa_identical_to_reference"])
print("apair", json.dumps(d.get("apair")))
PY
timeout 300 python tools/bench_io.py C4 > $O/io_C4.json 2> $O/io_C4.err; echo "io exit $?"; cat $O/io_C4.json
