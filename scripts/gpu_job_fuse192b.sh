#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
NC_DEC_K1=f16 timeout 300 python scripts/decoder_precision_exp.py 3 6 mixed > gpurun_out/decprec_fuse192b.log 2>&1
NC_DEC_K1=f16 timeout 200 python scripts/layer_profile.py 16 30 > gpurun_out/layers14.txt 2>&1
tail -1 gpurun_out/decprec_fuse192b.log; head -1 gpurun_out/layers14.txt; grep "decoder.block.[012]" gpurun_out/layers14.txt
