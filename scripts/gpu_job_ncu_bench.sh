#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_|rvq_|dia_" -c 1400 --csv \
  --log-file gpurun_out/r01_launches_bench_dac.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu2.log 2>&1
grep -c '^"' gpurun_out/r01_launches_bench_dac.csv; cut -c1-200 gpurun_out/bench_under_ncu2.log | tail -2
