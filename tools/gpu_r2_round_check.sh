#!/bin/bash
# one box: GPU tests, then the bench lines of the shard / C1 / C2 (cycle counters off)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for spec in "shard:--workload c2 --rows 1250000" "c1:--workload c1" "c2:--workload c2"; do
  name=${spec%%:*}; args=${spec#*:}
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $args > gpurun_out/rc_$name.json 2> gpurun_out/rc_$name.err
  echo "== $name rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/rc_$name.json').read().strip().splitlines()[-1])
    r=d['roofline']; v=d['verified'] or {}; a=d.get('api_e2e') or {}
    print('  QPS %.0f e2e %.0f api %s ms/step %.3f kernel %.3f frac %.3f fb %d ids_exact %s clocks %s' % (d['value'], d['e2e']['value'], a.get('value'), d['ms_per_step'], r['kernel_ms'], r['frac'], d['fallback_queries'], v.get('ids_exact_vs_exact_plan'), d['clocks']['sm_mhz']))
except Exception as e:
    print('  no line:', e)
PY
done
