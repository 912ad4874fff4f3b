#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dac_gpu.py -x -q 2>&1 | tail -12
timeout 120 python scripts/layer_profile.py 16 30 > gpurun_out/layers17.txt 2>&1; echo "rc=$?"
head -1 gpurun_out/layers17.txt; grep "ru_fused" gpurun_out/layers17.txt | head -4
