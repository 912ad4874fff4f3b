#!/bin/bash
# round 2, job AB: one-box A/B of two builds of the library (scratch/lib_old.so = previous commit, scratch/lib_new.so = working tree)
mkdir -p gpurun_out/r02ab
for rep in 1 2; do
  for v in old new; do
    cp scratch/lib_$v.so neuralcodecs_b200/libneuralcodecs_cuda.so
    timeout 300 python scripts/layer_profile.py 8 30 > gpurun_out/r02ab/layers_${v}$rep.txt 2>&1
    f=gpurun_out/r02ab/layers_${v}$rep.txt
    echo "$v$rep DAC $(head -1 $f | sed 's/.*total//') | fused $(grep ru_fused $f | awk '{s+=$3} END {print s}') umma $(grep 'conv_umma_bf16x3' $f | awk '{s+=$3} END {print s}') h16 $(grep 'conv_h16' $f | awk '{s+=$3} END {print s}')"
    timeout 300 python scripts/time_codec.py encodec 64 10 prof=0 2>&1 | tail -1
    timeout 300 python scripts/time_codec.py snac 32 10 prof=0 2>&1 | tail -1
  done
done
cp scratch/lib_new.so neuralcodecs_b200/libneuralcodecs_cuda.so
