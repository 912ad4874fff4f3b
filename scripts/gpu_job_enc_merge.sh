#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_encodec_gpu.py tests/test_snac_gpu.py tests/test_dac_gpu.py -x -q 2>&1 | tail -6
timeout 200 python scripts/time_codec.py encodec 64 10 prof=1 2>&1 | head -5
