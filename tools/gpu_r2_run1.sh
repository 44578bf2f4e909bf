#!/bin/bash
# round-2 GPU call 1: parity suite, MMA-rate probe for both operand kinds, first bench lines of every workload
mkdir -p gpurun_out
export LYNSE_B200_TC_TRACE=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -15 gpurun_out/r2a_pytest.log
timeout 300 python tools/mma_probe.py > gpurun_out/r2_mma_issue_probe.txt 2>&1
for w in c2 c3 c1 c4; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_$w.json 2> gpurun_out/r2a_bench_$w.err
  echo "== $w rc=$?"; tail -c 1500 gpurun_out/r2a_bench_$w.json; tail -5 gpurun_out/r2a_bench_$w.err
done
LYNSE_B200_TC_OPERAND=bf16 timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_c2_bf16.json 2> gpurun_out/r2a_bench_c2_bf16.err
echo "== c2 bf16 rc=$?"; tail -c 1200 gpurun_out/r2a_bench_c2_bf16.json
LYNSE_B200_TC_OPERAND=bf16 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tc_" > gpurun_out/r2a_pytest_bf16.log 2>&1; tail -5 gpurun_out/r2a_pytest_bf16.log
