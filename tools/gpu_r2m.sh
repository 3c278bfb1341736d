#!/bin/bash
O=gpurun_out/r2m; mkdir -p $O
for off in 1 0; do
  if [ $off = 1 ]; then export KB200_THIN_SMALL_OFF=1; else unset KB200_THIN_SMALL_OFF; fi
  KB200_TRACE=1 timeout 600 python bench.py --workload C3 --steps 1 --warmup 1 --no-cpu-baseline > $O/t_$off.json 2> $O/t_$off.err
  echo "== thin-small off=$off"; grep "tree level \(12\|16\|20\|21\):" $O/t_$off.err | tail -4; grep "jobs=1 small" $O/t_$off.err | tail -2
  for w in C3 C2; do
    timeout 600 python bench.py --workload $w --steps 3 --warmup 2 --no-cpu-baseline > $O/b_${w}_$off.json 2> $O/b_${w}_$off.err
    echo -n "$w: "; grep -o '"ms_per_step": [0-9.]*\|"msa_identical_to_reference": [a-z]*' $O/b_${w}_$off.json | tr '\n' ' '; echo
  done
done
