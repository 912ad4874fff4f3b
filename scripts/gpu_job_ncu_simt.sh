#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_simt_k3 -s 0 -c 1 -f -o gpurun_out/prof_simt_k3 python scripts/time_codec.py encodec 64 10 prof=0 > gpurun_out/ncu_simt1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv_simt_kernel<32" -s 0 -c 1 -f -o gpurun_out/prof_simt_32 python scripts/time_codec.py encodec 64 10 prof=0 > gpurun_out/ncu_simt2.log 2>&1
for n in simt_k3 simt_32; do
ncu -i gpurun_out/prof_$n.ncu-rep --page source --csv > gpurun_out/${n}_src.csv 2>/dev/null
ncu -i gpurun_out/prof_$n.ncu-rep --page details > gpurun_out/${n}_details.txt 2>/dev/null
rm -f gpurun_out/prof_$n.ncu-rep
done
ls -la gpurun_out/simt_*
