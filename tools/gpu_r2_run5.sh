#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/golden/ref_tests > gpurun_out/r2f_pytest.log 2>&1; tail -15 gpurun_out/r2f_pytest.log
timeout 900 python -m pytest tests/golden/ref_tests -m gpu -q > gpurun_out/r2f_pytest_ref.log 2>&1; tail -40 gpurun_out/r2f_pytest_ref.log
bash tools/gpu_exp.sh r2f c4 3 default:
python bench.py --workload c2 --rows 1250000 --steps 20 --warmup 5 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(\"c2 shard 1.25M rows: ms/step %.3f e2e %.3f kernel %.3f\" % (d[\"ms_per_step\"], d[\"e2e\"][\"ms_per_step\"], d[\"roofline\"][\"kernel_ms\"]))"
