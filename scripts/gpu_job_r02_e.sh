#!/bin/bash
# round 2, call E: fp16-operand kernel after removing the release.cluster membars; full DAC GPU tests; second parity set
mkdir -p gpurun_out/r02e
NC_H16_PAIR=1 timeout 300 python scripts/layer_profile.py 8 30 > gpurun_out/r02e/layers_h16_pair.txt 2>&1; head -1 gpurun_out/r02e/layers_h16_pair.txt
NC_H16_PAIR=0 timeout 300 python scripts/layer_profile.py 8 30 > gpurun_out/r02e/layers_h16_single.txt 2>&1; head -1 gpurun_out/r02e/layers_h16_single.txt
for k in 1 4 8 16 29; do
  NC_KNOCK=$k timeout 300 python scripts/layer_profile.py 8 30 > gpurun_out/r02e/layers_h16_knock$k.txt 2>&1
  head -1 gpurun_out/r02e/layers_h16_knock$k.txt
done
timeout 900 python -m pytest tests/test_dac_gpu.py -x -q -s > gpurun_out/r02e/pytest_dac.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error|flips|near-tie|h16 vs|e2e|config" gpurun_out/r02e/pytest_dac.log | tail -30
V='{"set2_default": {}}'
timeout 600 python scripts/parity_exp_gpu.py 12 10 40 "$V" 2>&1 | tail -2
