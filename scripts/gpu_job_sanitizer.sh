#!/bin/bash
# compute-sanitizer (memcheck, then racecheck on shared memory) over small GPU tests of each codec
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
SEL="test_mid_tcgen05_path or test_tiny_fp32_cuda_core_path_is_tight or test_dia_handoff or test_tiny_tensor_core_path or test_conv_stacks_without_lstm_tensor_core or test_ecdc_compress or test_input_conditioning"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 --log-file gpurun_out/sanitizer_memcheck.log \
  python -m pytest tests -m gpu -x -q -k "$SEL" > gpurun_out/sanitizer_memcheck_pytest.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck_pytest.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 --log-file gpurun_out/sanitizer_racecheck.log \
  python -m pytest tests -m gpu -x -q -k "test_tiny_fp32_cuda_core_path_is_tight or test_tiny_tensor_core_path or test_ecdc_compress" > gpurun_out/sanitizer_racecheck_pytest.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck_pytest.log
tail -4 gpurun_out/sanitizer_memcheck_pytest.log; tail -5 gpurun_out/sanitizer_memcheck.log; tail -4 gpurun_out/sanitizer_racecheck_pytest.log; tail -5 gpurun_out/sanitizer_racecheck.log
