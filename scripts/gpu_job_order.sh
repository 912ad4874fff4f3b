#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dac_gpu.py tests/test_snac_gpu.py tests/test_encodec_gpu.py -x -q 2>&1 | tail -3 > gpurun_out/pytest_order.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv -k regex:"conv_|rvq" \
  --log-file gpurun_out/launches_dram_order.csv python scripts/one_forward.py 16 30 > gpurun_out/ncu_order.log 2>&1
timeout 200 python scripts/layer_profile.py 16 30 > gpurun_out/layers12.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_full5.json 2> gpurun_out/bench_full5.err
cat gpurun_out/pytest_order.log; head -1 gpurun_out/layers12.txt; cut -c1-200 gpurun_out/bench_full5.json
