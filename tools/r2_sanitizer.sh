#!/bin/bash
# 1-GPU box: compute-sanitizer memcheck + racecheck over a small slice of the GPU suite (VERDICT hygiene item)
set -u
mkdir -p gpurun_out
SEL="test_cells_and_sort_order_bit_exact or test_deposit_fp32 or test_gather or test_short_range_forces or test_incremental_sort_equals_full_sort or test_step_matches_manual_sequence or test_poisson_with_reference_table"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -k "$SEL" > gpurun_out/sanitizer_memcheck_r02.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|leaked" gpurun_out/sanitizer_memcheck_r02.log | tail -5
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -k "test_deposit_fp32 or test_gather or test_incremental_sort_equals_full_sort or test_poisson_with_reference_table" > gpurun_out/sanitizer_racecheck_r02.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitizer_racecheck_r02.log | tail -5
