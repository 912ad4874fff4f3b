#!/bin/bash
# round 2, job V: role timeline of conv_ru_fused_kernel (CTA 0, clock64 at hand-offs) at C = 64 / 128 / 96
mkdir -p gpurun_out/r02v
for c in 64 128 96; do
  NC_TRACE_RU=gpurun_out/r02v/trace_c$c.txt NC_TRACE_RU_C=$c timeout 300 python scripts/layer_profile.py 8 30 > gpurun_out/r02v/layers_c$c.txt 2>&1
  echo "== C=$c: $(head -1 gpurun_out/r02v/layers_c$c.txt | sed 's/.*total//')"
  python scripts/ru_trace_analyze.py gpurun_out/r02v/trace_c$c.txt 30 34 > gpurun_out/r02v/analysis_c$c.txt 2>&1
  tail -32 gpurun_out/r02v/analysis_c$c.txt
done
