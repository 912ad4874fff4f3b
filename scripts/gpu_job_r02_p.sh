#!/bin/bash
mkdir -p gpurun_out/r02p
timeout 900 python -m pytest tests/test_snac_gpu.py -x -q > gpurun_out/r02p/pytest_snac.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02p/pytest_snac.log
