#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/core_peaks.py > gpurun_out/r2_core_peaks.log 2>&1; tail -25 gpurun_out/r2_core_peaks.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2g_pytest.log 2>&1; tail -6 gpurun_out/r2g_pytest.log
for w in c2 c1 c3; do
timeout 900 python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/r2g_bench_$w.json 2> gpurun_out/r2g_bench_$w.err; echo "== $w rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2g_bench_$w.json').read().strip().splitlines()[-1])
print(' value %.0f e2e %.0f api %s cpu %s' % (d['value'], d['e2e']['value'], d.get('api_e2e'), d.get('cpu_baseline')))
print(' verified', d['verified'])
PY
done
