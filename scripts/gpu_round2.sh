#!/bin/bash
# Lean GPU-box session: parity tests, contract bench, ncu launch list + full captures of the top kernels.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_round2.sh <tag> [sections]
# sections: any of  t(ests) q(uick) b(ench) l(aunches) n(cu tma) d(ncu direct) p(ncu plan) r(eference arm) c(oma train)
TAG=${1:-r1}
SEC=${2:-tblnp}
OUT=gpurun_out
mkdir -p $OUT
if [[ $SEC == *s* ]]; then echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke_$TAG.log; fi
if [[ $SEC == *t* ]]; then echo "== tests"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_$TAG.log; fi
if [[ $SEC == *q* ]]; then echo "== quick"; timeout 400 python scripts/quick_bench.py 2>&1 | tail -12 | tee $OUT/quick_$TAG.log; fi
if [[ $SEC == *b* ]]; then echo "== bench.py"; timeout 600 python bench.py 2>$OUT/bench_$TAG.err | tail -1 | tee $OUT/bench_$TAG.json; fi
if [[ $SEC == *r* ]]; then echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 10 --warmup 3 2>>$OUT/bench_$TAG.err | tail -1 | tee $OUT/bench_ref_$TAG.json; fi
if [[ $SEC == *c* ]]; then echo "== coma train"; timeout 600 python scripts/train_bench.py --envs 8192 --iters 3 2>$OUT/train_$TAG.err | tail -1 | tee $OUT/train_$TAG.json; fi
if [[ $SEC == *l* ]]; then echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv --log-file $OUT/launches_$TAG.csv \
  python bench.py --steps 45 --warmup 15 --no-cpu-baseline --no-shapes --no-train > $OUT/ncu_list_$TAG.log 2>&1; fi
if [[ $SEC == *n* ]]; then echo "== ncu full (tma)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_tma -s 20 -c 1 -f -o $OUT/prof_tma_$TAG \
  python bench.py --steps 45 --warmup 15 --no-cpu-baseline --no-shapes --no-train > $OUT/ncu_tma_$TAG.log 2>&1; fi
if [[ $SEC == *d* ]]; then echo "== ncu full (direct)"
IPP_STEP_VARIANT=direct timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_direct -s 20 -c 1 -f -o $OUT/prof_direct_$TAG \
  python bench.py --steps 45 --warmup 15 --no-cpu-baseline --no-shapes --no-train > $OUT/ncu_direct_$TAG.log 2>&1; fi
if [[ $SEC == *p* ]]; then echo "== ncu full (plan)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:plan_kernel -s 20 -c 1 -f -o $OUT/prof_plan_$TAG \
  python bench.py --steps 45 --warmup 15 --no-cpu-baseline --no-shapes --no-train > $OUT/ncu_plan_$TAG.log 2>&1; fi
ls -la $OUT | tail -20
