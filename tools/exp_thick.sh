#!/bin/bash
# thick vs thin strips in the few-big-boxes regime (top tree levels, long sequences) + GPU parity tests
TAG=${1:-exp}
O=gpurun_out/$TAG
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/pytest.log
tail -5 $O/pytest.log
for thin in 0 1; do
  KB200_THIN=$thin KB200_TRACE=1 timeout 200 python tools/check_long.py 2 8000 0 > $O/ss_thin$thin.out 2> $O/ss_thin$thin.err
  KB200_THIN=$thin KB200_TRACE=1 timeout 200 python tools/check_long.py 4 6000 0 > $O/pp_thin$thin.out 2> $O/pp_thin$thin.err
done
KB200_THIN=0 KB200_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/c3_thin0.json 2> $O/c3_thin0.err
KB200_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/c3_auto.json 2> $O/c3_auto.err
for f in $O/ss_thin*.err $O/pp_thin*.err; do echo $f; grep "round=[012] \|small" $f | tail -8; done
for f in $O/c3_*.err; do echo $f; grep "tree level" $f | tail -21 | awk '{s+=$11} END {print "dp sum", s}'; grep "tree level \(12\|14\|20\|21\)" $f | tail -4; done
cat $O/*.out
grep -o '"ms_per_step": [0-9.]*' $O/c3_*.json
grep -o '"value": [0-9.]*' $O/c3_*.json | head -4
