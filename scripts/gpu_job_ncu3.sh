#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_ru_fused -s 0 -c 1 -f -o gpurun_out/prof_ruf python scripts/one_forward.py 4 30 > gpurun_out/ncu_ruf.log 2>&1
ncu -i gpurun_out/prof_ruf.ncu-rep --page raw --csv > gpurun_out/ruf_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_ruf.ncu-rep --page source --csv > gpurun_out/ruf_src.csv 2>/dev/null
rm -f gpurun_out/prof_ruf.ncu-rep
ls -la gpurun_out/ruf_*
