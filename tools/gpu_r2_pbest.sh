#!/bin/bash
# exchange between the partitions of a query: second best (1, shipped), best (2, diagnostics), off (0); long and short shards
for pb in 1 2 0; do
for rows in 1250000 10000000; do
LYNSE_B200_TC_PBEST=$pb LYNSE_B200_TC_PROF=1 LYNSE_B200_TC_TRACE=1 python bench.py --workload c2 --rows $rows --steps 10 --warmup 3 --no-cpu-baseline --no-api-e2e 2> gpurun_out/pb.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pbest=$pb c2 rows $rows: ms/step %.3f kernel %.3f fb %d ids %s' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['fallback_queries'], d['verified'].get('ids_exact_vs_exact_plan')))"
grep "scan (" gpurun_out/pb.err | tail -1
done
done
timeout 900 python -m pytest tests/test_gpu_tile_scan.py -m gpu -q -x 2>&1 | tail -2
