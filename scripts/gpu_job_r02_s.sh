#!/bin/bash
# round 2, job S: Encodec 48 kHz preset -- tests incl. the .ecdc container, timing + per-kernel profile, 24 kHz regression
mkdir -p gpurun_out/r02s
timeout 900 python -m pytest tests/test_encodec48_gpu.py -x -q -s -m gpu > gpurun_out/r02s/pytest_encodec48.log 2>&1
echo "encodec48 rc=$?"; tail -25 gpurun_out/r02s/pytest_encodec48.log
timeout 600 python -m pytest tests/test_encodec_gpu.py -x -q -m gpu > gpurun_out/r02s/pytest_encodec24.log 2>&1
echo "encodec24 rc=$?"; tail -3 gpurun_out/r02s/pytest_encodec24.log
timeout 600 python scripts/time_codec.py encodec48 32 10 > gpurun_out/r02s/time_encodec48_b32x10s.txt 2>&1
head -45 gpurun_out/r02s/time_encodec48_b32x10s.txt
timeout 600 python scripts/time_codec.py encodec 64 10 prof=0 2>&1 | tail -2
