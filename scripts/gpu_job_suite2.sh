#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_all2.log
timeout 600 python bench.py > gpurun_out/bench_full4.json 2> gpurun_out/bench_full4.err
cat gpurun_out/pytest_all2.log; cut -c1-300 gpurun_out/bench_full4.json
