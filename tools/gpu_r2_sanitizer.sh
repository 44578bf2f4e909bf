#!/bin/bash
# compute-sanitizer over the parity tests that exercise every new kernel path (memcheck), and racecheck over the tensor-core
# and row-tile ones (the kernels that synchronise through shared memory, mbarriers and warp-level hand-offs)
mkdir -p gpurun_out
SEL="tc_raw_scores or tc_plan_matches or both_operand_kinds or aligned or baseline_c1 or binary_metrics_on_the_tensor or binary16_rows_give or merge_shards or tc_l2_rows_without or tc_l2_side_value or packed_index or flat_haversine or several_device or two_round or tile_scan or large_k_seeded or helper_warp"
timeout 3000 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r2_memcheck.log python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_tile_scan.py -m gpu -q -x -k "$SEL" > gpurun_out/r2_memcheck_pytest.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/r2_memcheck_pytest.log; tail -4 gpurun_out/r2_memcheck.log
SEL2="tc_raw_scores or both_operand_kinds or tc_l2_side_value or binary_metrics_of_an_f32 or two_round or tile_scan_equals or tile_scan_with_a_row_filter or tile_scan_large_k or large_k_seeded"
# (racecheck does not follow the mbarrier-ordered queue between the scanner and helper warps of the list-mode pair kernel, and the
# instrumented scan of a partition's first tiles runs into the kernel's 2 s barrier time-out: that kernel shape is covered by
# memcheck above and by byte-equality with LYNSE_B200_TC_HELPER=0, which is what racecheck runs)
LYNSE_B200_TC_HELPER=0 timeout 3000 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/r2_racecheck.log python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_tile_scan.py -m gpu -q -x -k "$SEL2" > gpurun_out/r2_racecheck_pytest.log 2>&1
echo "racecheck rc=$?"; tail -3 gpurun_out/r2_racecheck_pytest.log; tail -4 gpurun_out/r2_racecheck.log
