#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2s_pytest.log 2>&1; tail -4 gpurun_out/r2s_pytest.log
python tools/core_peaks.py > gpurun_out/r2s_core_peaks.log 2>&1; echo "core_peaks rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2s_bench_c2.json 2> gpurun_out/r2s_bench_c2.err; echo "c2 rc=$?"
python bench.py --workload c1 --steps 20 --warmup 5 > gpurun_out/r2s_bench_c1.json 2> gpurun_out/r2s_bench_c1.err; echo "c1 rc=$?"
python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2s_bench_c3.json 2> gpurun_out/r2s_bench_c3.err; echo "c3 rc=$?"
python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2s_bench_c4.json 2> gpurun_out/r2s_bench_c4.err; echo "c4 rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2s_bench_c2_reference.json 2> gpurun_out/r2s_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
for w in ('c1','c2','c3','c4'):
    try:
        d=json.loads([l for l in open(f'gpurun_out/r2s_bench_{w}.json') if l.startswith('{')][-1])
        a=d.get('api_e2e') or {}
        print(w, 'QPS %.0f e2e %.0f ms %.3f kernel %.3f frac %.3f fb %d ids %s | api %.3f ms native %.3f py %.3f (%.1f%%)' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['fallback_queries'], (d['verified'] or {}).get('ids_exact_vs_exact_plan'), a.get('ms_per_step',0), a.get('device_index_search_ms',0), a.get('python_layer_ms',0), 100*a.get('python_layer_frac_of_step',0)))
    except Exception as e: print(w, 'no line', e)
PY
