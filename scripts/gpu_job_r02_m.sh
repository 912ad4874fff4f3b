#!/bin/bash
# round 2, call M: Encodec -- selective short chains, two-slice interleaved LSTM
mkdir -p gpurun_out/r02m
timeout 900 python -m pytest tests/test_encodec_gpu.py -x -q -s > gpurun_out/r02m/pytest_encodec.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|encodec24k|Error|error" gpurun_out/r02m/pytest_encodec.log | tail -5
timeout 300 python scripts/time_codec.py encodec 64 10 > gpurun_out/r02m/encodec_tc.txt 2>&1; head -8 gpurun_out/r02m/encodec_tc.txt
NC_LSTM2=0 timeout 300 python scripts/time_codec.py encodec 64 10 > gpurun_out/r02m/encodec_tc_lstm1.txt 2>&1; head -4 gpurun_out/r02m/encodec_tc_lstm1.txt
timeout 900 python scripts/encodec_chain_exp.py 8 10 > gpurun_out/r02m/encodec_chain_exp.txt 2>&1; tail -7 gpurun_out/r02m/encodec_chain_exp.txt
