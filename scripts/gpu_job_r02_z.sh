#!/bin/bash
# round 2, job Z: role timelines of conv_umma_kernel on the narrow layers of Encodec 24 kHz (k3 conv C=32: bf16x3 decoder / 3xtf32 encoder,
# 1x1 C=32) and on a SNAC depthwise unit (C=64)
mkdir -p gpurun_out/r02z
t() {  # name codec batch match skip
  NC_TRACE_UMMA=gpurun_out/r02z/trace_$1.txt NC_TRACE_UMMA_MATCH=$4 NC_TRACE_UMMA_SKIP=$5 timeout 300 python scripts/time_codec.py $2 $3 10 prof=0 > gpurun_out/r02z/run_$1.txt 2>&1
  tail -1 gpurun_out/r02z/run_$1.txt
  python scripts/ru_trace_analyze.py gpurun_out/r02z/trace_$1.txt 30 32 > gpurun_out/r02z/analysis_$1.txt 2>&1
  echo "== $1"; head -1 gpurun_out/r02z/analysis_$1.txt; sed -n '/steady-state period/,$p' gpurun_out/r02z/analysis_$1.txt | grep -v "^$"
}
t enc_k3_dec_bf16x3 encodec 64 32,32,3 3
t enc_k3_enc_3xtf32 encodec 64 32,32,3 2
t enc_1x1_c32 encodec 64 32,32,1 3
t snac_dw_c64 snac 32 64,64,1 3
