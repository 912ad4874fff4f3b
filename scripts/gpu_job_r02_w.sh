#!/bin/bash
# round 2, job W: full GPU suite, smoke, default bench (with other_configs incl. Encodec 48 kHz) at the current code
mkdir -p gpurun_out/r02w
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02w/pytest.log; cat gpurun_out/r02w/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02w/smoke.log 2>&1; tail -2 gpurun_out/r02w/smoke.log
timeout 900 python bench.py > gpurun_out/r02w/bench.json 2> gpurun_out/r02w/bench.err; tail -3 gpurun_out/r02w/bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02w/bench.json"))
r = d["roofline"]
print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 1), "clocks", d["clocks"])
print("roofline", r["bound"], round(r["frac"], 3), r.get("kernel", "")[:40])
for k, v in d.get("other_configs", {}).items():
    print(" ", k, round(v["value"], 1) if isinstance(v, dict) else v)
print("parity", d.get("parity"))
PY
