#!/bin/bash
# Round-2 profiling session (under gpurun): launch list of the contract bench + ncu --set full captures of every kernel
# of the path.  Results in gpurun_out/ (digests are made offline with scripts/ncu_summary.py).
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --steps 45 --warmup 15 --no-cpu-baseline --no-shapes --no-train"
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv --log-file $OUT/r02_launches.csv $B > $OUT/r02_ncu_list.log 2>&1
for k in step_tma plan_kernel reset_fill reset_prep; do
  echo "== ncu full $k"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o $OUT/r02_prof_$k $B > $OUT/r02_ncu_$k.log 2>&1
done
for k in features_actor features_critic ig_plan eval_metrics own_update; do
  echo "== ncu full $k"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o $OUT/r02_prof_$k python scripts/split_loop.py > $OUT/r02_ncu_$k.log 2>&1
done
ls -la $OUT | tail -30
