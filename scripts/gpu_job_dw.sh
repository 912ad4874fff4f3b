#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_snac_gpu.py -x -q 2>&1 | tail -12
timeout 120 python scripts/time_codec.py snac 32 10 prof=1 2>&1 | head -8
timeout 120 python scripts/time_codec.py snac 32 10 prof=1 fuse_dw=0 2>&1 | head -3
timeout 120 python scripts/layer_profile.py 16 30 > gpurun_out/layers18.txt 2>&1; head -1 gpurun_out/layers18.txt
