#!/bin/bash
# round 2, job T: dissecting the skeleton of the fused residual-unit kernel: NC_KNOCK_RU bits 2 transform, 8 MMAs, 64 drain math,
# 128 weight stream, 256 epilogue math, 512 store, 1024 residual load
mkdir -p gpurun_out/r02t
for k in 0 128 74 202 458 970 1994 1920; do
  NC_KNOCK_RU=$k timeout 300 python scripts/layer_profile.py 8 30 > gpurun_out/r02t/layers_ru_knock$k.txt 2>&1
  echo "== knock_ru $k: $(head -1 gpurun_out/r02t/layers_ru_knock$k.txt | sed 's/.*total//')"
  grep -E "encoder.block.0.res_unit1|encoder.block.1.res_unit1|decoder.block.3.res_unit1" gpurun_out/r02t/layers_ru_knock$k.txt | awk '{printf "%s %s %s | ", $1, $3, $5} END {print ""}'
done
