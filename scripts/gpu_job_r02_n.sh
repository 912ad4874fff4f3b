#!/bin/bash
# round 2, call N: transform row mapping without bank conflicts (DAC), single-thread fence in the LSTM barrier (Encodec)
mkdir -p gpurun_out/r02n
timeout 300 python scripts/layer_profile.py 8 30 > gpurun_out/r02n/layers.txt 2>&1; head -1 gpurun_out/r02n/layers.txt
grep -E "encoder.block.[01].res_unit1|decoder.block.3.res_unit1|encoder.block.2.res_unit1.conv1" gpurun_out/r02n/layers.txt | awk '{printf "%s %s %s | ", $1, $3, $5} END {print ""}'
timeout 300 python scripts/time_codec.py encodec 64 10 > gpurun_out/r02n/encodec.txt 2>&1; head -4 gpurun_out/r02n/encodec.txt
timeout 300 python scripts/time_codec.py snac 32 10 > gpurun_out/r02n/snac.txt 2>&1; head -3 gpurun_out/r02n/snac.txt
timeout 900 python -m pytest tests/test_encodec_gpu.py tests/test_snac_gpu.py -x -q > gpurun_out/r02n/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02n/pytest.log
