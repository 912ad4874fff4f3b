#!/bin/bash
# round 2, call F: deeper epilogue ring; full GPU suite; ncu tensor-pipe list of one forward; bench line
mkdir -p gpurun_out/r02f
timeout 300 python scripts/layer_profile.py 8 30 > gpurun_out/r02f/layers.txt 2>&1; head -1 gpurun_out/r02f/layers.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02f/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02f/pytest.log
timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none -k regex:"conv_|rvq_|f32_to" -s 79 -c 79 --csv --log-file gpurun_out/r02f/launches_fwd_b4x30s.csv \
  python scripts/one_forward.py 4 30 reps=2 > gpurun_out/r02f/ncu_fwd.log 2>&1; echo "ncu rc=$?"; tail -1 gpurun_out/r02f/ncu_fwd.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02f/bench.json 2> gpurun_out/r02f/bench.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r02f/bench.json
