#!/bin/bash
# round 2, call B: short accumulation chains (acc_split) -- correctness, code parity, cost
mkdir -p gpurun_out/r02b
timeout 600 python -m pytest tests/test_dac_gpu.py -x -q > gpurun_out/r02b/pytest_dac.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02b/pytest_dac.log
V='{"sc0": {"encoder_short_chains": "0"}, "sc1": {}, "sc2": {"encoder_short_chains": "2"}, "sc3": {"encoder_short_chains": "3"}}'
timeout 600 python scripts/parity_exp_gpu.py 10 10 11 "$V" 2>&1 | tail -5
V='{"sc1_f24": {}, "sc3_f24": {"encoder_short_chains": "3"}}'
NC_FOLD_STEPS=24 timeout 600 python scripts/parity_exp_gpu.py 10 10 11 "$V" 2>&1 | tail -3
V='{"sc1_f96": {}, "sc3_f96": {"encoder_short_chains": "3"}}'
NC_FOLD_STEPS=96 timeout 600 python scripts/parity_exp_gpu.py 10 10 11 "$V" 2>&1 | tail -3
V='{"sc1_f1000": {}, "sc3_f1000": {"encoder_short_chains": "3"}}'
NC_FOLD_STEPS=100000 timeout 600 python scripts/parity_exp_gpu.py 10 10 11 "$V" 2>&1 | tail -3
timeout 300 python scripts/layer_profile.py 8 30 > gpurun_out/r02b/layers_sc1.txt 2>&1; head -1 gpurun_out/r02b/layers_sc1.txt
timeout 300 python scripts/layer_profile.py 8 30 bf16x3 mixed encoder_short_chains=0 > gpurun_out/r02b/layers_sc0.txt 2>&1; head -1 gpurun_out/r02b/layers_sc0.txt
timeout 300 python scripts/layer_profile.py 8 30 bf16x3 mixed encoder_short_chains=3 > gpurun_out/r02b/layers_sc3.txt 2>&1; head -1 gpurun_out/r02b/layers_sc3.txt
