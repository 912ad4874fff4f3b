#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for c in snac encodec; do
B=32; [ $c = encodec ] && B=64
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active --clock-control none \
  -k regex:"conv_|vq|lstm|dwconv|attn|layernorm|reflect|randn|trim|decode_codes" --launch-skip 0 -c 400 --csv --log-file gpurun_out/r01_launches_dram_${c}.csv \
  python scripts/time_codec.py $c $B 10 prof=0 > gpurun_out/ncu_${c}.log 2>&1
grep -c '^"' gpurun_out/r01_launches_dram_${c}.csv
done
