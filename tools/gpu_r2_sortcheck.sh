#!/bin/bash
timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for w in "c2 --rows 1250000" "c2" "c1" "c3"; do
python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-api-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w: ms/step %.3f e2e %.3f kernel %.3f fb %d ids %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['fallback_queries'], d['verified'].get('ids_exact_vs_exact_plan')))"
done
python tools/latency_probe.py 2>&1 | tail -4
