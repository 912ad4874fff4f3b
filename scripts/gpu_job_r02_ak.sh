#!/bin/bash
# round 2, job AK: Encodec with the LSTM slicing inside run_lstm (convs on the whole micro-batch): tests + timings
mkdir -p gpurun_out/r02ak
timeout 900 python -m pytest tests/test_encodec_gpu.py tests/test_encodec48_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python scripts/time_codec.py encodec48 32 10 > gpurun_out/r02ak/time_encodec48.txt 2>&1; head -8 gpurun_out/r02ak/time_encodec48.txt
timeout 300 python scripts/time_codec.py encodec 64 10 prof=0 2>&1 | tail -1
timeout 300 python scripts/time_codec.py encodec 256 10 prof=0 2>&1 | tail -1
