#!/bin/bash
# final check of the round: smoke(), the full GPU parity suite, the default bench line with its CPU baseline,
# the reference arm, the other workloads, an ncu launch list of the bench command
O=gpurun_out/final; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?"; tail -1 $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?"; tail -3 $O/pytest.log
timeout 900 python bench.py > $O/bench_C3.json 2> $O/bench_C3.err; echo "bench exit $?"
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_C3_reference.json 2> $O/bench_C3_reference.err; echo "reference arm exit $?"
for w in C2 C4; do timeout 600 python bench.py --workload $w --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err; done
timeout 900 python bench.py --workload C5 --steps 1 --warmup 0 --no-cpu-baseline > $O/bench_C5.json 2> $O/bench_C5.err
for f in C3 C2 C4 C5; do echo -n "$f: "; grep -o '"ms_per_step": [0-9.]*\|"seconds_per_call": [0-9.]*\|"msa_identical_to_reference": [a-z]*\|"frac": [0-9.]*' $O/bench_$f.json | tr '\n' ' '; echo; done
head -c 400 $O/bench_C3_reference.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches.csv python bench.py --workload C3 --steps 1 --warmup 0 --no-cpu-baseline > $O/ncu_launch.log 2>&1
KB200_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/trace_C3.json 2> $O/trace_C3.err; python tools/trace_sum.py $O/trace_C3.err
