#!/bin/bash
# round 2, call D: where does the fp16-operand kernel spend its time (knock-outs + ncu), block RVQ kernel check
mkdir -p gpurun_out/r02d
NC_RVQ_BLOCK=0 timeout 300 python scripts/rvq_check.py gpurun_out/r02d/rvq_warp.npz 2>&1 | tail -1
NC_RVQ_BLOCK=1 timeout 300 python scripts/rvq_check.py gpurun_out/r02d/rvq_block.npz 2>&1 | tail -1
python - <<'P'
import numpy as np
a=np.load("gpurun_out/r02d/rvq_warp.npz"); b=np.load("gpurun_out/r02d/rvq_block.npz")
print("rvq block vs warp: codes equal", np.array_equal(a["codes"], b["codes"]), "z equal", np.array_equal(a["z"], b["z"]), "latents equal", np.array_equal(a["lat"], b["lat"]))
P
for k in 0 1 4 8 16 5 13 29; do
  NC_KNOCK=$k timeout 300 python scripts/layer_profile.py 8 30 > gpurun_out/r02d/layers_h16_knock$k.txt 2>&1
  head -1 gpurun_out/r02d/layers_h16_knock$k.txt
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_h16 --launch-skip 19 --launch-count 11 -o gpurun_out/r02d/ncu_h16 python scripts/one_decode.py 4 10 > gpurun_out/r02d/ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r02d/ncu.log
ls -la gpurun_out/r02d/
