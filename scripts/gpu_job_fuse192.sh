#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 300 python scripts/decoder_precision_exp.py 3 6 mixed > gpurun_out/decprec_fuse192.log 2>&1
timeout 200 python scripts/layer_profile.py 16 30 > gpurun_out/layers13.txt 2>&1
tail -1 gpurun_out/decprec_fuse192.log; head -1 gpurun_out/layers13.txt; grep "decoder.block.2" gpurun_out/layers13.txt
timeout 600 python -m pytest tests/test_dac_gpu.py -x -q 2>&1 | tail -3
