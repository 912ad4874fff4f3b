#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
run() { # name, dec mode, env...
  name=$1; mode=$2; shift 2
  env "$@" timeout 300 python scripts/decoder_precision_exp.py 3 6 $mode > gpurun_out/decprec_$name.log 2>&1
  env "$@" timeout 200 python scripts/layer_profile.py 16 30 bf16x3 $mode > gpurun_out/layers_dec_$name.txt 2>&1
  echo "== $name"; tail -1 gpurun_out/decprec_$name.log; head -1 gpurun_out/layers_dec_$name.txt
}
run f16x2_k1x3 f16x2 A=1
run f16x2_widef16 f16x2 NC_DEC_WIDE=f16
run bf16x3_widef16x2 bf16x3 NC_DEC_WIDE=f16x2
run bf16x3_widef16 bf16x3 NC_DEC_WIDE=f16
