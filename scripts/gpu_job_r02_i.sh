#!/bin/bash
# round 2, call I: sustained bench line + ncu tensor-pipe list of one whole forward
mkdir -p gpurun_out/r02i
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02i/bench.json 2> gpurun_out/r02i/bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r02i/bench.json
timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none -k regex:"conv_|rvq_|f32_to" -s 53 -c 53 --csv --log-file gpurun_out/r02i/launches_fwd_b4x30s.csv \
  python scripts/one_forward.py 4 30 reps=2 > gpurun_out/r02i/ncu_fwd.log 2>&1; echo "ncu rc=$?"; tail -1 gpurun_out/r02i/ncu_fwd.log
