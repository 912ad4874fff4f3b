#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_all3.log
timeout 300 python bench.py --workload encodec24k_b64x10s > gpurun_out/bench_encodec2.json 2> gpurun_out/bench_encodec2.err
timeout 300 python bench.py --workload snac24k_b32x10s > gpurun_out/bench_snac2.json 2> gpurun_out/bench_snac2.err
tail -5 gpurun_out/pytest_all3.log; cut -c1-200 gpurun_out/bench_encodec2.json; cut -c1-200 gpurun_out/bench_snac2.json
