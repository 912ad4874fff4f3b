#!/bin/bash
# round 2, job AA: store-warp epilogue -- full GPU suite + timings of the three codecs
mkdir -p gpurun_out/r02aa
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02aa/pytest.log; cat gpurun_out/r02aa/pytest.log
timeout 300 python scripts/time_codec.py encodec 64 10 prof=0 2>&1 | tail -1
timeout 300 python scripts/time_codec.py snac 32 10 prof=0 2>&1 | tail -1
timeout 300 python scripts/time_codec.py encodec48 32 10 prof=0 2>&1 | tail -1
timeout 300 python scripts/layer_profile.py 8 30 > gpurun_out/r02aa/layers_dac.txt 2>&1; head -1 gpurun_out/r02aa/layers_dac.txt
