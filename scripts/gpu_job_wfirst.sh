#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for w in 6 8 12; do
NC_W_FIRST=$w timeout 200 python scripts/layer_profile.py 16 30 > gpurun_out/layers_wfirst$w.txt 2>&1
echo "== W_FIRST=$w"; head -1 gpurun_out/layers_wfirst$w.txt; grep "decoder.block.[012].res_unit1.conv1\|block.1.conv_t1\|encoder.block.[23].res_unit1.conv1" gpurun_out/layers_wfirst$w.txt
done
