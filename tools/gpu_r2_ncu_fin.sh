#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:finalize_kernel -s 3 -c 1 -f -o gpurun_out/r2_fin_c3 python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline --no-api-e2e --verify-queries 4 > gpurun_out/r2_ncu_fin_c3.log 2>&1; echo "rc=$?"
ncu -i gpurun_out/r2_fin_c3.ncu-rep --page raw --csv > gpurun_out/r2_fin_c3_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_fin_c3.ncu-rep --page source --csv > gpurun_out/r2_fin_c3_source.csv 2>/dev/null
rm -f gpurun_out/r2_fin_c3.ncu-rep
