#!/bin/bash
# ncu --set full + source page of the coarse kernel on a 1.25M-row shard (what each of 8 GPUs runs on C2)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:coarse_pair_kernel -s 2 -c 1 -f -o gpurun_out/r2_shard python bench.py --workload c2 --rows 1250000 --steps 1 --warmup 1 --no-cpu-baseline --no-api-e2e --verify-queries 2 > gpurun_out/r2_ncu_shard.log 2>&1; echo "rc=$?"
ncu -i gpurun_out/r2_shard.ncu-rep --page raw --csv > gpurun_out/r2_shard_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_shard.ncu-rep --page source --csv > gpurun_out/r2_shard_source.csv 2>/dev/null
rm -f gpurun_out/r2_shard.ncu-rep
ls -la gpurun_out/r2_shard*
