#!/bin/bash
# round 2, call L: Encodec with the encoder on tensor cores (3xTF32, short chains) -- tests + timing against the fp32 encoder
mkdir -p gpurun_out/r02l
timeout 900 python -m pytest tests/test_encodec_gpu.py -x -q -s > gpurun_out/r02l/pytest_encodec.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|encodec24k" gpurun_out/r02l/pytest_encodec.log | tail -5
timeout 300 python scripts/time_codec.py encodec 64 10 > gpurun_out/r02l/encodec_tc.txt 2>&1; head -14 gpurun_out/r02l/encodec_tc.txt
timeout 300 python scripts/time_codec.py encodec 64 10 encoder_precision=fp32 > gpurun_out/r02l/encodec_fp32.txt 2>&1; head -8 gpurun_out/r02l/encodec_fp32.txt
timeout 300 python scripts/time_codec.py snac 32 10 > gpurun_out/r02l/snac.txt 2>&1; head -12 gpurun_out/r02l/snac.txt
