#!/bin/bash
# round 2, job AG: Encodec VQ stage with row-major codebook staging -- Encodec tests (codes are compared with the oracle) + timing
mkdir -p gpurun_out/r02ag
timeout 900 python -m pytest tests/test_encodec_gpu.py tests/test_encodec48_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python scripts/time_codec.py encodec 64 10 > gpurun_out/r02ag/time_encodec.txt 2>&1; head -8 gpurun_out/r02ag/time_encodec.txt
