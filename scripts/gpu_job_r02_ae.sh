#!/bin/bash
# round 2, job AE: one-box A/B of two library builds on the DAC forward (burst per-layer) + DAC parity tests on the new build
mkdir -p gpurun_out/r02ae
for rep in 1 2 3; do
  for v in old new; do
    cp scratch/lib_$v.so neuralcodecs_b200/libneuralcodecs_cuda.so
    f=gpurun_out/r02ae/layers_${v}$rep.txt
    timeout 300 python scripts/layer_profile.py 8 30 > $f 2>&1
    echo "$v$rep DAC $(head -1 $f | sed 's/.*total//') | fused $(grep ru_fused $f | awk '{s+=$3} END {print s}') umma $(grep 'conv_umma_bf16x3' $f | awk '{s+=$3} END {print s}') h16 $(grep 'conv_h16' $f | awk '{s+=$3} END {print s}')"
  done
done
cp scratch/lib_new.so neuralcodecs_b200/libneuralcodecs_cuda.so
timeout 900 python -m pytest tests/test_dac_gpu.py -x -q -m gpu 2>&1 | tail -2
