#!/bin/bash
# compute-sanitizer over a small slice of the GPU tests (new kernels of the round): memcheck, then racecheck.
set -u
mkdir -p gpurun_out
K='test_pair_lists_that_overflow or test_both_filter_kernels_and_their_score_tables or test_tensor_core_filter_with_ties or test_two_pass_approx_with_massive_ties'
timeout -k 10 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py -x -q -k "$K" > gpurun_out/sanitize_memcheck.txt 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/sanitize_memcheck.txt | tail -8
timeout -k 10 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py -x -q -k "test_pair_lists_that_overflow or test_two_pass_approx_with_massive_ties" > gpurun_out/sanitize_racecheck.txt 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard|Error" gpurun_out/sanitize_racecheck.txt | tail -8
