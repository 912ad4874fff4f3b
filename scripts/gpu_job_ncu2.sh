#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for s in 26 33; do
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_umma_kernel -s $s -c 1 -f -o gpurun_out/prof_k7_$s python scripts/one_forward.py 4 30 > gpurun_out/ncu_k7_$s.log 2>&1
ncu -i gpurun_out/prof_k7_$s.ncu-rep --page raw --csv > gpurun_out/k7_${s}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_k7_$s.ncu-rep --page source --csv > gpurun_out/k7_${s}_src.csv 2>/dev/null
ncu -i gpurun_out/prof_k7_$s.ncu-rep --page details > gpurun_out/k7_${s}_details.txt 2>/dev/null
rm -f gpurun_out/prof_k7_$s.ncu-rep
done
ls -la gpurun_out | tail -8
