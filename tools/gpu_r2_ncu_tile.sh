#!/bin/bash
mkdir -p gpurun_out
M=${1:-l1}; D=${2:-256}
rows=$((1024000000 / D / 4 * 4))
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_tile_kernel -s 2 -c 1 -f -o gpurun_out/r2_tile_$M python tools/metric_sweep.py $rows $D 64 $M > gpurun_out/r2_ncu_tile_$M.log 2>&1; echo "rc=$?"
ncu -i gpurun_out/r2_tile_$M.ncu-rep --page raw --csv > gpurun_out/r2_tile_${M}_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_tile_$M.ncu-rep --page source --csv > gpurun_out/r2_tile_${M}_source.csv 2>/dev/null
ncu -i gpurun_out/r2_tile_$M.ncu-rep --page details > gpurun_out/r2_tile_${M}_details.txt 2>/dev/null
rm -f gpurun_out/r2_tile_$M.ncu-rep
tail -3 gpurun_out/r2_ncu_tile_$M.log
