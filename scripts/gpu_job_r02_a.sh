#!/bin/bash
# round 2, call A: regression tests after the host-side refactors, code-parity experiment, knock-out timing
mkdir -p gpurun_out/r02a
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02a/pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02a/pytest.log
timeout 900 python scripts/parity_exp_gpu.py 10 10 11 > gpurun_out/r02a/parity_exp.log 2>&1; echo "parity rc=$?"
cat gpurun_out/r02a/parity_exp.log | tail -12
for k in 0 1 2 4 8 16 32 3 7 15; do
  NC_KNOCK=$k timeout 300 python scripts/layer_profile.py 8 30 > gpurun_out/r02a/layers_knock$k.txt 2>&1
  head -1 gpurun_out/r02a/layers_knock$k.txt
done
