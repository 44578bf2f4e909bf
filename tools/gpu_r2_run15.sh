#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2q_pytest.log 2>&1; tail -4 gpurun_out/r2q_pytest.log
for v in 1 0; do
for rows in 1250000 10000000; do
LYNSE_B200_FIN_TWO_ROUNDS=$v python bench.py --workload c2 --rows $rows --steps 20 --warmup 5 --no-cpu-baseline --no-api-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('two_rounds=$v c2 rows $rows: ms/step %.3f e2e %.3f kernel %.3f fb %d ids %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['fallback_queries'], d['verified'].get('ids_exact_vs_exact_plan')))"
done
LYNSE_B200_FIN_TWO_ROUNDS=$v python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline --no-api-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('two_rounds=$v c3: ms/step %.3f e2e %.3f kernel %.3f fb %d ids %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['fallback_queries'], d['verified'].get('ids_exact_vs_exact_plan')))"
done
