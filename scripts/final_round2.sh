#!/bin/bash
# Final evidence run of round 2 (under gpurun): smoke, GPU tests, the contract bench as the driver runs it, the full
# bench line, the reference arm, the launch list and the last ncu captures.  Results in gpurun_out/.
OUT=gpurun_out
mkdir -p $OUT
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2 | tee $OUT/r02_smoke.log
echo "== tests"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee $OUT/r02_pytest_gpu.log
echo "== bench (driver arguments)"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 2>$OUT/r02_bench.err | tail -1 > $OUT/r02_bench_n1_driver_args.json
echo "== bench (default)"; timeout 900 python bench.py 2>>$OUT/r02_bench.err | tail -1 > $OUT/r02_bench_n1.json
echo "== reference arm"; timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2>>$OUT/r02_bench.err | tail -1 > $OUT/r02_bench_reference.json
B="python bench.py --steps 45 --warmup 15 --no-cpu-baseline --no-shapes --no-train"
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv --log-file $OUT/r02_launches.csv $B > $OUT/r02_ncu_list.log 2>&1
for k in step_tma reset_fill plan_kernel; do
  echo "== ncu full $k"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o $OUT/r02_prof_$k $B > $OUT/r02_ncu_$k.log 2>&1
done
for k in features_actor features_critic eval_metrics; do
  echo "== ncu full $k"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o $OUT/r02_prof_$k python scripts/split_loop.py > $OUT/r02_ncu_$k.log 2>&1
done
python scripts/kernel_times.py 2>&1 | grep tma | tee $OUT/r02_kernel_times.log
python scripts/split_times.py 2>&1 | tail -1 | tee $OUT/r02_split_times.log
tail -c 300 $OUT/r02_bench.err
ls -la $OUT | tail -12
