#!/bin/bash
for poll in 7 3 1 0; do
for rows in 1250000 10000000; do
LYNSE_B200_TC_POLL=$poll python bench.py --workload c2 --rows $rows --steps 20 --warmup 5 --no-cpu-baseline --no-api-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('poll=$poll c2 rows $rows: ms/step %.3f kernel %.3f fb %d ids %s' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['fallback_queries'], d['verified'].get('ids_exact_vs_exact_plan')))"
done
done
timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
