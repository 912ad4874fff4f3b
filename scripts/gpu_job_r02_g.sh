#!/bin/bash
# round 2, call G: epilogue ring with early stage release; knock-outs on the epilogue streams
mkdir -p gpurun_out/r02g
for k in 0 32 64 96 16; do
  NC_KNOCK=$k timeout 300 python scripts/layer_profile.py 8 30 > gpurun_out/r02g/layers_knock$k.txt 2>&1
  head -1 gpurun_out/r02g/layers_knock$k.txt
  grep -E "decoder.block.1.res_unit1.conv2|decoder.block.2.res_unit1.conv2|decoder.block.0.res_unit1.conv2|decoder.block.1.conv_t1" gpurun_out/r02g/layers_knock$k.txt | awk '{printf "%s %s %s | ", $1, $3, $6} END {print ""}'
done
timeout 600 python -m pytest tests/test_dac_gpu.py -x -q > gpurun_out/r02g/pytest_dac.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02g/pytest_dac.log
