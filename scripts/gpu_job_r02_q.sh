#!/bin/bash
# round 2, call Q: ncu source-level profile of the fused residual-unit kernel (C = 64, 128) and the lo-accumulator conv kernel (C = 256 k7)
mkdir -p gpurun_out/r02q
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_ru_fused|conv_umma" --launch-skip 26 --launch-count 10 -o gpurun_out/r02q/ncu_enc python scripts/one_forward.py 4 10 reps=2 > gpurun_out/r02q/ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r02q/ncu.log
ls -la gpurun_out/r02q
