#!/bin/bash
# round 2, job Y: per-layer tables of SNAC 24 kHz (32 x 10 s) and Encodec 24 kHz (64 x 10 s)
mkdir -p gpurun_out/r02y
timeout 600 python scripts/time_codec.py snac 32 10 prof=2 > gpurun_out/r02y/snac_layers.txt 2>&1; head -45 gpurun_out/r02y/snac_layers.txt
timeout 600 python scripts/time_codec.py encodec 64 10 prof=2 > gpurun_out/r02y/encodec_layers.txt 2>&1; head -45 gpurun_out/r02y/encodec_layers.txt
