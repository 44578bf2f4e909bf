#!/bin/bash
# where does the short-shard tile time go: epilogue that does not scan (debug bit 2) / does not read (bit 1)
mkdir -p gpurun_out
for spec in base: noscan:LYNSE_B200_TC_DEBUG=4 noread:LYNSE_B200_TC_DEBUG=2 nopoll:LYNSE_B200_TC_DEBUG=128; do
  name=${spec%%:*}; envs=${spec#*:}
  env $envs LYNSE_B200_TC_TRACE=1 LYNSE_B200_TC_PROF=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-api-e2e --verify-queries 0 --workload c2 --rows 1250000 > gpurun_out/dbg_$name.json 2> gpurun_out/dbg_$name.err
  echo "== $name rc=$?"; grep -o '"kernel_ms": [0-9.]*' gpurun_out/dbg_$name.json | tail -1
  grep -E "per tile|scan \(" gpurun_out/dbg_$name.err | tail -2
done
