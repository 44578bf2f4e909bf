#!/bin/bash
for parts in 9 12; do
for rows in 1250000 10000000; do
LYNSE_B200_TC_PARTS=$parts python bench.py --workload c2 --rows $rows --steps 5 --warmup 3 --no-cpu-baseline --no-api-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('parts=$parts c2 rows $rows: ms/step %.3f kernel %.3f fb %d ids %s P %s' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['fallback_queries'], d['verified'].get('ids_exact_vs_exact_plan'), d['partitions']))"
done
done
