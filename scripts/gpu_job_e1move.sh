#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 120 python scripts/decoder_precision_exp.py 3 6 mixed bf16x3 > gpurun_out/decprec_e1.log 2>&1; echo "rc=$?"
tail -2 gpurun_out/decprec_e1.log
timeout 120 python scripts/layer_profile.py 16 30 > gpurun_out/layers15.txt 2>&1; echo "rc=$?"
head -1 gpurun_out/layers15.txt; grep "ru_fused" gpurun_out/layers15.txt
timeout 400 python -m pytest tests/test_dac_gpu.py -x -q 2>&1 | tail -3
