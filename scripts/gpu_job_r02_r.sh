#!/bin/bash
# round 2, job R: Encodec 48 kHz preset -- first GPU run of the new tests + the 24 kHz suite (regression)
mkdir -p gpurun_out/r02r
timeout 900 python -m pytest tests/test_encodec48_gpu.py -x -q -s -m gpu > gpurun_out/r02r/pytest_encodec48.log 2>&1
echo "encodec48 rc=$?"; tail -30 gpurun_out/r02r/pytest_encodec48.log
timeout 900 python -m pytest tests/test_encodec_gpu.py -x -q -m gpu > gpurun_out/r02r/pytest_encodec24.log 2>&1
echo "encodec24 rc=$?"; tail -5 gpurun_out/r02r/pytest_encodec24.log
