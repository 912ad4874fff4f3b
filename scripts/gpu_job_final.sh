#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_final.log
timeout 200 python scripts/layer_profile.py 16 30 > gpurun_out/layers16.txt 2>&1
timeout 600 python bench.py > gpurun_out/bench_full6.json 2> gpurun_out/bench_full6.err
cat gpurun_out/pytest_final.log; head -1 gpurun_out/layers16.txt; cut -c1-200 gpurun_out/bench_full6.json
