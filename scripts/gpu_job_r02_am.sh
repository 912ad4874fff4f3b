#!/bin/bash
# round 2, job AM: second MMA-issuing warp in conv_umma_kernel (NC_DUAL_ISSUE=0/1): Encodec / SNAC / DAC tests, then a one-box A/B
mkdir -p gpurun_out/r02am
timeout 600 python -m pytest tests/test_encodec_gpu.py tests/test_encodec48_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 600 python -m pytest tests/test_snac_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 900 python -m pytest tests/test_dac_gpu.py -x -q -m gpu 2>&1 | tail -2
for rep in 1 2; do
  for t in 0 1; do
    export NC_DUAL_ISSUE=$t
    echo "dual=$t rep=$rep"
    timeout 300 python scripts/time_codec.py encodec 64 10 prof=0 2>&1 | tail -1
    timeout 300 python scripts/time_codec.py encodec48 32 10 prof=0 2>&1 | tail -1
    timeout 300 python scripts/time_codec.py snac 32 10 prof=0 2>&1 | tail -1
    f=gpurun_out/r02am/layers_d${t}_$rep.txt
    timeout 300 python scripts/layer_profile.py 8 30 > $f 2>&1
    echo "DAC $(head -1 $f | sed 's/.*total//') | umma $(grep 'conv_umma_bf16x3' $f | awk '{s+=$3} END {print s}')"
  done
done
