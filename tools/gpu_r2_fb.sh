#!/bin/bash
# fallback queries and step time by shard size and by the depth of the second-best exchange (LYNSE_B200_TC_PBEST_DEPTH)
mkdir -p gpurun_out
for spec in "d4:" "d3:LYNSE_B200_TC_PBEST_DEPTH=3" "d2:LYNSE_B200_TC_PBEST_DEPTH=2" "nopb:LYNSE_B200_TC_PBEST=0"; do
  name=${spec%%:*}; envs=${spec#*:}
  for rows in 1250000 2500000 5000000 10000000; do
    env $envs timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-api-e2e --workload c2 --rows $rows > gpurun_out/fb_${name}_$rows.json 2> gpurun_out/fb_${name}_$rows.err
    python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/fb_${name}_$rows.json').read().strip().splitlines()[-1])
    print('$name rows $rows: QPS %.0f ms/step %.3f kernel %.3f fb %d parts %s ids %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['fallback_queries'], d['partitions'], (d['verified'] or {}).get('ids_exact_vs_exact_plan')))
except Exception as e:
    print('$name $rows no line', e)
PY
  done
done
