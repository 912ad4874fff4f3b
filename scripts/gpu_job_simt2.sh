#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 200 python scripts/time_codec.py encodec 64 10 prof=2 2>&1 | head -24
