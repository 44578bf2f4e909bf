#!/bin/bash
mkdir -p gpurun_out
for v in "A=1" "LYNSE_B200_FIN_R=128" "LYNSE_B200_FIN_R=256" "LYNSE_B200_FIN_THREADS=512" "LYNSE_B200_FIN_THREADS=1024" "LYNSE_B200_FIN_THREADS=128"; do
env $v timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:finalize -s 2 -c 2 --csv --log-file gpurun_out/r2m_fin.csv python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline --no-api-e2e --verify-queries 4 > /dev/null 2>&1
echo "$v: $(grep finalize gpurun_out/r2m_fin.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"
done
