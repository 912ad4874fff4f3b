#!/bin/bash
# round 2, call O: full GPU suite + bench line after the e2e copy/compute overlap
mkdir -p gpurun_out/r02o
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02o/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02o/pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02o/bench.json 2> gpurun_out/r02o/bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r02o/bench.json; tail -3 gpurun_out/r02o/bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r02o/bench.json').read())
print('value',d['value'],'e2e',d['e2e']['value'],'clocks',d['clocks'])
print(d['other_configs'])
"
