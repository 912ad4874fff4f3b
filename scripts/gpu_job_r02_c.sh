#!/bin/bash
# round 2, call C: first run of the fp16-operand decoder kernel (single-CTA, then CTA pairs), chain policy re-check
mkdir -p gpurun_out/r02c
NC_H16_PAIR=0 timeout 180 python scripts/h16_check.py mid 2 1 > gpurun_out/r02c/h16_mid_single.log 2>&1; echo "mid single rc=$?"; tail -6 gpurun_out/r02c/h16_mid_single.log
NC_H16_PAIR=1 timeout 180 python scripts/h16_check.py mid 2 1 > gpurun_out/r02c/h16_mid_pair.log 2>&1; echo "mid pair rc=$?"; tail -6 gpurun_out/r02c/h16_mid_pair.log
NC_H16_PAIR=0 timeout 300 python scripts/h16_check.py full 2 2 > gpurun_out/r02c/h16_full_single.log 2>&1; echo "full single rc=$?"; tail -6 gpurun_out/r02c/h16_full_single.log
NC_H16_PAIR=1 timeout 300 python scripts/h16_check.py full 2 2 > gpurun_out/r02c/h16_full_pair.log 2>&1; echo "full pair rc=$?"; tail -6 gpurun_out/r02c/h16_full_pair.log
V='{"p1": {}, "p1_precise": {"fast_sin": "0"}, "p2": {"encoder_short_chains": "2"}, "p0_precise": {"encoder_short_chains": "0", "fast_sin": "0"}}'
timeout 600 python scripts/parity_exp_gpu.py 10 10 11 "$V" 2>&1 | tail -5
NC_H16_PAIR=0 timeout 300 python scripts/layer_profile.py 8 30 > gpurun_out/r02c/layers_h16_single.txt 2>&1; head -1 gpurun_out/r02c/layers_h16_single.txt
NC_H16_PAIR=1 timeout 300 python scripts/layer_profile.py 8 30 > gpurun_out/r02c/layers_h16_pair.txt 2>&1; head -1 gpurun_out/r02c/layers_h16_pair.txt
timeout 300 python scripts/layer_profile.py 8 30 bf16x3 mixed decoder_h16=0 > gpurun_out/r02c/layers_f32act.txt 2>&1; head -1 gpurun_out/r02c/layers_f32act.txt
timeout 600 python -m pytest tests/test_dac_gpu.py -x -q > gpurun_out/r02c/pytest_dac.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02c/pytest_dac.log
