#!/bin/bash
# round 2, job AN: sustained (power-capped) A/B on ONE box of the tree at commit 64cf155 (before this round's last third: fused-unit smem split,
# coefficient caches, two-team epilogue, fast ELU, ...) against the final tree: bench.py --steps 3 --warmup 3 --no-other-configs, alternating
mkdir -p gpurun_out/r02an
R=$PWD
for i in 1 2; do
  (cd scratch/old_tree && timeout 600 python bench.py --steps 3 --warmup 3 --no-other-configs > $R/gpurun_out/r02an/bench_old$i.json 2>/dev/null)
  timeout 600 python bench.py --steps 3 --warmup 3 --no-other-configs > gpurun_out/r02an/bench_new$i.json 2>/dev/null
done
python - <<'PY'
import json
for n in ("old1", "new1", "old2", "new2"):
    d = json.load(open(f"gpurun_out/r02an/bench_{n}.json"))
    print(n, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "sm_mhz", d["clocks"]["sm_mhz"], "power_w_max", d["clocks"]["power_w_max"])
PY
