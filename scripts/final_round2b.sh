#!/bin/bash
# Re-run of the parts of final_round2.sh that depend on the map kernel (after the ld.global.cg change).
OUT=gpurun_out
mkdir -p $OUT
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1 | tee $OUT/r02_smoke.log
echo "== tests"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -2 | tee $OUT/r02_pytest_gpu.log
echo "== bench (driver arguments)"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 2>$OUT/r02_bench.err | tail -1 > $OUT/r02_bench_n1_driver_args.json
echo "== bench (default)"; timeout 900 python bench.py 2>>$OUT/r02_bench.err | tail -1 > $OUT/r02_bench_n1.json
B="python bench.py --steps 45 --warmup 15 --no-cpu-baseline --no-shapes --no-train"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv --log-file $OUT/r02_launches.csv $B > $OUT/r02_ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_tma -s 6 -c 1 -f -o $OUT/r02_prof_step_tma $B > $OUT/r02_ncu_step_tma.log 2>&1
python scripts/kernel_times.py 2>&1 | grep tma | tee $OUT/r02_kernel_times.log
tail -c 200 $OUT/r02_bench.err
