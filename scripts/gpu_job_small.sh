#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 300 python bench.py --workload dac44k_b1x10s --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_b1.json 2> gpurun_out/bench_b1.err
timeout 400 python bench.py --workload dia_dac_decode_b256x20s --no-cpu-baseline > gpurun_out/bench_dia.json 2> gpurun_out/bench_dia.err
cut -c1-400 gpurun_out/bench_b1.json; cut -c1-400 gpurun_out/bench_dia.json; tail -2 gpurun_out/bench_dia.err
