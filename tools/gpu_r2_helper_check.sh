#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r2_memcheck_helper.log python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "helper_warp or baseline_c1" > gpurun_out/r2_memcheck_helper_pytest.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/r2_memcheck_helper_pytest.log; tail -4 gpurun_out/r2_memcheck_helper.log
