#!/bin/bash
# usage: tools/gpu_exp.sh <tag> <workload> <steps> name1:ENV=V,ENV2=V name2:...   -> one summary line per variant
tag=$1; wl=$2; steps=$3; shift 3
mkdir -p gpurun_out
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}; [ "$envs" = "$spec" ] && envs=""
  envs=$(echo "$envs" | tr ',' ' ')
  env $envs LYNSE_B200_TC_TRACE=1 LYNSE_B200_TC_PROF=1 timeout 600 python bench.py --workload $wl --steps $steps --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_${wl}_$name.json 2> gpurun_out/${tag}_${wl}_$name.err
  echo "== $wl $name ($envs) rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_${wl}_$name.json').read().strip().splitlines()[-1])
    r=d['roofline']; v=d['verified'] or {}
    print('  QPS %.0f e2e %.0f ms/step %.3f kernel %.3f frac %.3f fb %d ids_exact %s clocks %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], r['kernel_ms'], r['frac'], d['fallback_queries'], v.get('ids_exact_vs_exact_plan'), d['clocks']['sm_mhz']))
except Exception as e:
    print('  no line:', e)
PY
  grep -E "per tile|scan \(" gpurun_out/${tag}_${wl}_$name.err | tail -2
done
