#!/bin/bash
# round 2, call K: Encodec encoder on tensor cores with short chains; knock-outs of the fused residual-unit kernel
mkdir -p gpurun_out/r02k
timeout 900 python scripts/encodec_chain_exp.py 8 10 > gpurun_out/r02k/encodec_chain_exp.txt 2>&1; echo "encodec rc=$?"; tail -7 gpurun_out/r02k/encodec_chain_exp.txt
for k in 0 2 8 64 10 74; do
  NC_KNOCK_RU=$k timeout 300 python scripts/layer_profile.py 8 30 > gpurun_out/r02k/layers_ru_knock$k.txt 2>&1
  echo "== knock_ru $k: $(head -1 gpurun_out/r02k/layers_ru_knock$k.txt | sed 's/.*total//')"
  grep -E "encoder.block.0.res_unit1|encoder.block.1.res_unit1|decoder.block.3.res_unit1" gpurun_out/r02k/layers_ru_knock$k.txt | awk '{printf "%s %s %s | ", $1, $3, $5} END {print ""}'
done
