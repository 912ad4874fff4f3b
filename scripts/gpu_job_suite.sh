#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" > gpurun_out/pytest_all.log; echo "rc=${PIPESTATUS[0]}" >> gpurun_out/pytest_all.log
timeout 300 python scripts/decoder_precision_exp.py 3 6 mixed bf16x3 f16x2 f16 > gpurun_out/decprec_default.log 2>&1
timeout 200 python scripts/layer_profile.py 16 30 > gpurun_out/layers10.txt 2>&1
timeout 600 python bench.py > gpurun_out/bench_full3.json 2> gpurun_out/bench_full3.err
tail -15 gpurun_out/pytest_all.log; cat gpurun_out/decprec_default.log; head -1 gpurun_out/layers10.txt; cut -c1-900 gpurun_out/bench_full3.json
