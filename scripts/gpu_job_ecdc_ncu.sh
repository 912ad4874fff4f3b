#!/bin/bash
# one GPU call: ecdc tests, launch list + DRAM traffic of one DAC forward, launch list of the bench command, full-set summary
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_encodec_gpu.py -x -q > gpurun_out/pytest_ecdc.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_ecdc.log
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/r01_launches_dram_dac_b16x30s.csv python scripts/one_forward.py 16 30 > gpurun_out/ncu_d1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r01_launches_bench_dac.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none -o gpurun_out/prof_r1d python scripts/one_forward.py 4 30 > gpurun_out/ncu_d2.log 2>&1
ncu -i gpurun_out/prof_r1d.ncu-rep --page raw --csv > gpurun_out/r01_ncu_full_dac_b4x30s_raw.csv 2>/dev/null
rm -f gpurun_out/prof_r1d.ncu-rep
ls -la gpurun_out
tail -3 gpurun_out/pytest_ecdc.log
