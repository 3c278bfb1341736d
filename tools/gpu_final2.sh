#!/bin/bash
# final check of the round (HEAD with the f-3 / f-4 rows): smoke(), the full GPU parity suite, the default bench
# line, the reference arm, C2 / C4 lines, an ncu capture of the identity-distance kernel, FASTA I/O on the box's cores
O=gpurun_out/final2; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?"; tail -1 $O/smoke.log
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?"; tail -4 $O/pytest.log
timeout 600 python bench.py > $O/bench_C3.json 2> $O/bench_C3.err; echo "bench exit $?"; tail -2 $O/bench_C3.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_C3_reference.json 2> $O/bench_C3_reference.err; echo "reference arm exit $?"
for w in C2 C4; do timeout 400 python bench.py --workload $w --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err; done
python - <<'PY'
import json
for w in ("C3", "C2", "C4"):
    try:
        d = json.load(open("gpurun_out/final2/bench_%s.json" % w))
        print(w, round(d["ms_per_step"], 2), round(d["e2e"]["seconds_per_call"], 4), round(d["e2e"].get("seconds_per_call_incl_python_marshalling", 0), 4),
              round(d["roofline"]["frac"], 3), d["msa_identical_to_reference"], "apair", (d.get("apair") or {}).get("kernel_seconds"), (d.get("apair") or {}).get("call_seconds"))
    except Exception as e:
        print(w, "failed", e)
PY
head -c 300 $O/bench_C3_reference.json; echo
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f -k "regex:kb_apair_tile_kernel" -c 1 -o $O/apair python tools/apair_driver.py 4096 5136 > $O/ncu_apair.log 2>&1; tail -2 $O/ncu_apair.log
python tools/apair_driver.py 10000 5136
timeout 200 python tools/bench_io.py C4 > $O/io_C4.json 2> $O/io_C4.err; echo "io exit $?"
# tunables at HEAD (each line checks the alignment hash against the reference's golden)
for v in "KB200_SMALL_ROWS_SS=32" "KB200_SMALL_ROWS_SS=128" "KB200_THIN_K=4" "KB200_SMALL_ROWS_PROF=32"; do
  env $v timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > $O/var.json 2> $O/var.err
  python - "$v" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/final2/var.json"))
    print(sys.argv[1], round(d["ms_per_step"], 2), round(d["e2e"]["seconds_per_call"], 4), d["msa_identical_to_reference"])
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
