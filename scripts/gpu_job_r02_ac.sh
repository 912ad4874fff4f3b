#!/bin/bash
# round 2, job AC: one-box A/B of the two-team epilogue (NC_EPI_TEAMS=0/1), alternating
mkdir -p gpurun_out/r02ac
for rep in 1 2; do
  for t in 0 1; do
    export NC_EPI_TEAMS=$t
    f=gpurun_out/r02ac/layers_t${t}_$rep.txt
    timeout 300 python scripts/layer_profile.py 8 30 > $f 2>&1
    echo "teams=$t rep=$rep DAC $(head -1 $f | sed 's/.*total//') | fused $(grep ru_fused $f | awk '{s+=$3} END {print s}') umma $(grep 'conv_umma_bf16x3' $f | awk '{s+=$3} END {print s}') h16 $(grep 'conv_h16' $f | awk '{s+=$3} END {print s}')"
    timeout 300 python scripts/time_codec.py encodec 64 10 prof=0 2>&1 | tail -1
    timeout 300 python scripts/time_codec.py snac 32 10 prof=0 2>&1 | tail -1
    timeout 300 python scripts/time_codec.py encodec48 32 10 prof=0 2>&1 | tail -1
  done
done
