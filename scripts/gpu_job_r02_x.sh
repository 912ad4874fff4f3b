#!/bin/bash
# round 2, job X: sustained A/B of the fused-unit smem split on ONE box (old 4/3/3 vs new default), alternating
mkdir -p gpurun_out/r02x
for i in 1 2; do
  NC_RU_AS=4 NC_RU_HS=3 NC_RU_ES=3 timeout 600 python bench.py --steps 3 --warmup 3 --no-other-configs > gpurun_out/r02x/bench_old$i.json 2>/dev/null
  timeout 600 python bench.py --steps 3 --warmup 3 --no-other-configs > gpurun_out/r02x/bench_new$i.json 2>/dev/null
done
python - <<'PY'
import json
for n in ("old1", "new1", "old2", "new2"):
    d = json.load(open(f"gpurun_out/r02x/bench_{n}.json"))
    print(n, round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["clocks"]["sm_mhz"], d["clocks"]["power_w_max"])
PY
