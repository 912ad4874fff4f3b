#!/bin/bash
# round 2, job AF: one-box A/B of the fused unit's two-team acc1 drain (NC_RU_E1_TEAMS=0/1) + DAC parity tests
mkdir -p gpurun_out/r02af
for rep in 1 2 3; do
  for t in 0 1; do
    f=gpurun_out/r02af/layers_t${t}_$rep.txt
    NC_RU_E1_TEAMS=$t timeout 300 python scripts/layer_profile.py 8 30 > $f 2>&1
    echo "e1_teams=$t rep=$rep DAC $(head -1 $f | sed 's/.*total//') | fused $(grep ru_fused $f | awk '{s+=$3} END {print s}') enc0 $(grep -E 'encoder.block.0.res' $f | awk '{s+=$3} END {print s}') enc1 $(grep -E 'encoder.block.1.res' $f | awk '{s+=$3} END {print s}') dec3 $(grep -E 'decoder.block.3.res' $f | awk '{s+=$3} END {print s}')"
  done
done
timeout 900 python -m pytest tests/test_dac_gpu.py -x -q -m gpu 2>&1 | tail -2
