#!/bin/bash
# One GPU-box session: parity tests, variant timings, contract bench, ncu launch list + full captures.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_round.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
echo "== tests"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_$TAG.log
echo "== quick bench"; timeout 400 python scripts/quick_bench.py 2>&1 | tail -12 | tee $OUT/quick_$TAG.log
echo "== bench.py"; timeout 600 python bench.py --steps 150 --warmup 15 2>$OUT/bench_$TAG.err | tail -1 | tee $OUT/bench_$TAG.json
echo "== bench.py direct"; IPP_STEP_VARIANT=direct timeout 600 python bench.py --steps 150 --warmup 15 --no-cpu-baseline 2>>$OUT/bench_$TAG.err | tail -1 | tee $OUT/bench_direct_$TAG.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv --log-file $OUT/launches_$TAG.csv \
  python bench.py --steps 45 --warmup 15 --no-cpu-baseline > $OUT/ncu_list_$TAG.log 2>&1
echo "== ncu full (tma)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_tma -s 20 -c 2 -f -o $OUT/prof_tma_$TAG \
  python bench.py --steps 45 --warmup 15 --no-cpu-baseline > $OUT/ncu_tma_$TAG.log 2>&1
echo "== ncu full (direct)"
IPP_STEP_VARIANT=direct timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_direct -s 20 -c 2 -f -o $OUT/prof_direct_$TAG \
  python bench.py --steps 45 --warmup 15 --no-cpu-baseline > $OUT/ncu_direct_$TAG.log 2>&1
ls -la $OUT | tail -20
