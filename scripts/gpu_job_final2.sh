#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_final2.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final2.log 2>&1
timeout 300 python bench.py --workload snac24k_b32x10s > gpurun_out/bench_snac4.json 2> gpurun_out/bench_snac4.err
timeout 300 python bench.py --workload encodec24k_b64x10s > gpurun_out/bench_encodec4.json 2> gpurun_out/bench_encodec4.err
timeout 600 python bench.py > gpurun_out/bench_full7.json 2> gpurun_out/bench_full7.err
cat gpurun_out/pytest_final2.log; tail -1 gpurun_out/smoke_final2.log
for f in snac4 encodec4 full7; do python - <<PY
import json
d=json.load(open("gpurun_out/bench_$f.json")); r=d["roofline"]; print("$f", round(d["value"]), round(d["e2e"]["value"]), r["bound"], round(r["frac"],3), r["kernel"][:30])
PY
done
