#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 300 python scripts/decoder_precision_exp.py 3 6 mixed f16x2 > gpurun_out/decprec_q.log 2>&1
timeout 200 python scripts/layer_profile.py 16 30 > gpurun_out/layers11.txt 2>&1
cat gpurun_out/decprec_q.log | tail -3; head -1 gpurun_out/layers11.txt; grep "decoder.block.[012].res_unit1.conv1\|conv_t1" gpurun_out/layers11.txt
