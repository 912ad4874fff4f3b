#!/bin/bash
# round 2, job U: fused residual-unit kernel -- shared-memory split sweep (A / H / E ring depths; the weight ring gets the rest)
mkdir -p gpurun_out/r02u
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 300 python scripts/layer_profile.py 8 30 > gpurun_out/r02u/layers_$name.txt 2>&1
  echo "== $name ($*): $(head -1 gpurun_out/r02u/layers_$name.txt | sed 's/.*total//')"
  grep -E "encoder.block.0.res_unit1|encoder.block.0.res_unit3|encoder.block.1.res_unit1|decoder.block.3.res_unit1" gpurun_out/r02u/layers_$name.txt | awk '{printf "%s %s | ", $1, $3} END {print ""}'
}
run default NC_X=0
run old NC_RU_AS=4 NC_RU_HS=3 NC_RU_ES=3
run e3 NC_RU_ES=3
run h3 NC_RU_HS=3
run a2 NC_RU_AS=2
run a3 NC_RU_AS=3
run a2e3 NC_RU_AS=2 NC_RU_ES=3
run a3e3 NC_RU_AS=3 NC_RU_ES=3
run a4 NC_RU_AS=4
timeout 600 python -m pytest tests/test_dac_gpu.py -x -q -m gpu -k "parity or oracle or fused" 2>&1 | tail -3
