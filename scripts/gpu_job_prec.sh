#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_encodec_gpu.py -x -q > gpurun_out/pytest_ecdc.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_ecdc.log
timeout 600 python scripts/decoder_precision_exp.py 6 6 > gpurun_out/decprec1.log 2>&1
for m in f16x2 f16; do
  timeout 200 python scripts/layer_profile.py 16 30 bf16x3 $m > gpurun_out/layers_dec_$m.txt 2>&1
done
tail -3 gpurun_out/pytest_ecdc.log; cat gpurun_out/decprec1.log; head -1 gpurun_out/layers_dec_*.txt
