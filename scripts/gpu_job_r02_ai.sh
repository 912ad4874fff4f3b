#!/bin/bash
# round 2, job AI: SNAC tests + timing (FFMA2 in the depthwise prologue)
mkdir -p gpurun_out/r02ai
timeout 900 python -m pytest tests/test_snac_gpu.py -x -q -m gpu 2>&1 | tail -2
for i in 1 2; do timeout 300 python scripts/time_codec.py snac 32 10 prof=0 2>&1 | tail -1; done
timeout 300 python scripts/time_codec.py snac 32 10 > gpurun_out/r02ai/time_snac.txt 2>&1; head -4 gpurun_out/r02ai/time_snac.txt
