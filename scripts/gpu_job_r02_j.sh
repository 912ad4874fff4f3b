#!/bin/bash
# round 2, call J: full GPU suite after the Encodec short-input branch
mkdir -p gpurun_out/r02j
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02j/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02j/pytest.log
