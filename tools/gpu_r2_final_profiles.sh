#!/bin/bash
# Round-2 evidence: ncu --set full captures of the dominant kernels (CSV exports only) and launch lists of every workload.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_ivf_client.py -m gpu -q -x 2>&1 | tail -2
bash tools/gpu_r2_ncu.sh
rm -f gpurun_out/*.ncu-rep
bash tools/gpu_r2_launches.sh
bash tools/gpu_r2_ncu_tile.sh l1 256
python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_bench_c3.json 2>/dev/null
python bench.py --workload c1 --steps 20 --warmup 5 > gpurun_out/r2u_bench_c1.json 2>/dev/null
python - <<'PY'
import json
for w in ('c1','c3'):
    d=json.loads([l for l in open(f'gpurun_out/r2u_bench_{w}.json') if l.startswith('{')][-1]); a=d.get('api_e2e') or {}
    print(w, 'QPS %.0f api %.3f ms native %.3f py %.3f (%.1f%%)' % (d['value'], a.get('ms_per_step',0), a.get('device_index_search_ms',0), a.get('python_layer_ms',0), 100*a.get('python_layer_frac_of_step',0)))
PY
du -sh gpurun_out
