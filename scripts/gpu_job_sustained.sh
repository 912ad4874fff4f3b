#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/sus_$name.json 2> gpurun_out/sus_$name.err; python - <<PY
import json
d=json.load(open("gpurun_out/sus_$name.json"))
print("$name", round(d["value"],1), d["clocks"]["sm_mhz"], {k:round(v["ms"]) for k,v in d["kernels"].items()})
PY
}
run unfused NC_RU_FUSE_WIDE_MAX_C=128
run fusedwide A=1
run fusedwide_k1f16 NC_DEC_K1=f16
run unfused_k1f16 NC_RU_FUSE_WIDE_MAX_C=128 NC_DEC_K1=f16
