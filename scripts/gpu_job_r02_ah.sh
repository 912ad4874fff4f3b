#!/bin/bash
# round 2, job AH: compute-sanitizer over small GPU tests of every codec at the final code (memcheck; racecheck on shared memory):
# covers the two-team epilogue, the coefficient caches, conv_h16 (mid config), the fused unit, the Encodec 48 kHz kernels
mkdir -p gpurun_out/r02ah
SEL="test_mid_tcgen05_path or test_tiny_fp32_cuda_core_path_is_tight or test_dia_handoff or test_tiny_tensor_core_path or test_conv_stacks_without_lstm_tensor_core or test_ecdc_compress or test_truncated_penultimate_segment or test_one_unnormalised_frame or test_local_attention_other_window_sizes or test_decoder_fp16_operand_path"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 --log-file gpurun_out/r02ah/memcheck.log \
  python -m pytest tests -m gpu -x -q -k "$SEL" > gpurun_out/r02ah/memcheck_pytest.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02ah/memcheck_pytest.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 3 --log-file gpurun_out/r02ah/racecheck.log \
  python -m pytest tests -m gpu -x -q -k "test_tiny_fp32_cuda_core_path_is_tight or test_tiny_tensor_core_path or test_ecdc_compress or test_one_unnormalised_frame" > gpurun_out/r02ah/racecheck_pytest.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r02ah/racecheck_pytest.log
tail -3 gpurun_out/r02ah/memcheck_pytest.log; tail -3 gpurun_out/r02ah/memcheck.log; tail -3 gpurun_out/r02ah/racecheck_pytest.log; tail -3 gpurun_out/r02ah/racecheck.log
