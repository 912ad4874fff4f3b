#!/bin/bash
# round 2, call H: 64-column epilogue stages in conv_h16
mkdir -p gpurun_out/r02h
timeout 300 python scripts/layer_profile.py 8 30 > gpurun_out/r02h/layers.txt 2>&1; head -1 gpurun_out/r02h/layers.txt
grep -E "^decoder" gpurun_out/r02h/layers.txt | awk '{printf "%-42s %8s %8s %8s\n",$1,$3,$5,$6}'
timeout 600 python -m pytest tests/test_dac_gpu.py -x -q > gpurun_out/r02h/pytest_dac.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02h/pytest_dac.log
