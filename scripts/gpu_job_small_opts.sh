#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python scripts/time_codec.py encodec 64 10 prof=1 2>&1 | head -4
timeout 200 python scripts/time_codec.py snac 32 10 prof=1 2>&1 | head -5
timeout 120 python scripts/layer_profile.py 16 30 2>&1 | head -5
