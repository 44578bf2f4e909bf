#!/bin/bash
mkdir -p gpurun_out
for rows in 1250000 10000000; do
for v in 1 0; do
LYNSE_B200_TC_SEED_SMALLK=$v LYNSE_B200_TC_TRACE=1 LYNSE_B200_TC_PROF=1 python bench.py --workload c2 --rows $rows --steps 10 --warmup 3 --no-cpu-baseline --no-api-e2e 2> gpurun_out/r2k_$rows.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2 rows $rows seed_smallk=$v: ms/step %.3f e2e %.3f kernel %.3f fb %d' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['fallback_queries']), d['verified']['ids_exact_vs_exact_plan'])"
grep -E "per tile|scan \(" gpurun_out/r2k_$rows.err | tail -2
done
done
