#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_exp.sh r2l c3 10 default:
bash tools/gpu_exp.sh r2l c4 3 default:
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2l_launches_c3.csv python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline --no-api-e2e --verify-queries 4 > /dev/null 2>&1
