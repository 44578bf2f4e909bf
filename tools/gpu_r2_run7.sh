#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2i_pytest.log 2>&1; tail -4 gpurun_out/r2i_pytest.log
for rows in 1250000 10000000; do
python bench.py --workload c2 --rows $rows --steps 20 --warmup 5 --no-cpu-baseline --no-api-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2 rows $rows: ms/step %.3f e2e %.3f kernel %.3f fb %d P %d' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['fallback_queries'], d['partitions']), d['verified']['ids_exact_vs_exact_plan'])"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2i_launches_c2shard.csv python bench.py --workload c2 --rows 1250000 --steps 2 --warmup 1 --no-cpu-baseline --no-api-e2e --verify-queries 4 > /dev/null 2>&1
