#!/bin/bash
# round 2, job AD: ncu launch list of one whole forward at the final code (refreshes profiles/r02_launches_fwd_b4x30s.csv, r02_tensor_pipe.json)
mkdir -p gpurun_out/r02ad
timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none -k regex:"conv_|rvq_|f32_to" -s 53 -c 53 --csv --log-file gpurun_out/r02ad/launches_fwd_b4x30s.csv \
  python scripts/one_forward.py 4 30 reps=2 > gpurun_out/r02ad/ncu_fwd.log 2>&1; echo "ncu rc=$?"; tail -1 gpurun_out/r02ad/ncu_fwd.log
python scripts/ncu_tensor_pipe.py gpurun_out/r02ad/launches_fwd_b4x30s.csv gpurun_out/r02ad/tensor_pipe.json "one DAC 44.1 kHz forward, 4 clips x 30 s (53 launches), final code of round 2" > gpurun_out/r02ad/summary.txt
tail -12 gpurun_out/r02ad/summary.txt
