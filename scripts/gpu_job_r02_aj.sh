#!/bin/bash
# round 2, job AJ: one-box A/B of packed fp32x2 math (FFMA2 / FMUL2) in the Snake prologues / epilogues and SNAC's depthwise taps
mkdir -p gpurun_out/r02aj
for rep in 1 2; do
  for v in old new; do
    cp scratch/lib_$v.so neuralcodecs_b200/libneuralcodecs_cuda.so
    f=gpurun_out/r02aj/layers_${v}$rep.txt
    timeout 300 python scripts/layer_profile.py 8 30 > $f 2>&1
    echo "$v$rep DAC $(head -1 $f | sed 's/.*total//') | fused $(grep ru_fused $f | awk '{s+=$3} END {print s}') umma $(grep 'conv_umma_bf16x3' $f | awk '{s+=$3} END {print s}') h16 $(grep 'conv_h16' $f | awk '{s+=$3} END {print s}')"
    timeout 300 python scripts/time_codec.py snac 32 10 prof=0 2>&1 | tail -1
  done
done
cp scratch/lib_new.so neuralcodecs_b200/libneuralcodecs_cuda.so
timeout 1200 python -m pytest tests/test_dac_gpu.py tests/test_snac_gpu.py -x -q -m gpu 2>&1 | tail -2
