#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_e2e.log
timeout 300 python bench.py --workload snac24k_b32x10s > gpurun_out/bench_snac3.json 2> gpurun_out/bench_snac3.err
timeout 300 python bench.py --workload encodec24k_b64x10s > gpurun_out/bench_encodec3.json 2> gpurun_out/bench_encodec3.err
timeout 300 python bench.py --workload dac44k_b1x10s --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_b1b.json 2> gpurun_out/bench_b1b.err
cat gpurun_out/pytest_e2e.log
for f in snac3 encodec3 b1b; do python - <<PY
import json
d=json.load(open("gpurun_out/bench_$f.json")); print("$f", round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"])
PY
done
