#!/bin/bash
# C2 corpus, batch-size sweep: where the search is HBM-bound (few queries) and where it is tensor-bound
mkdir -p gpurun_out
: > gpurun_out/r2_q_sweep.jsonl
for nq in 1 8 64 128 256 512 1024 4096; do
  python bench.py --workload c2 --nq $nq --steps 10 --warmup 3 --no-cpu-baseline --no-api-e2e --verify-queries 4 2>/dev/null | grep '^{' | tail -1 >> gpurun_out/r2_q_sweep.jsonl
done
python - <<'PY'
import json
print("| Q | plan | ms per batch | QPS | dominant kernel ms | bytes the kernel must read | GB/s of those bytes | of HBM peak | TOP/s (2QND) | fallbacks |")
print("|---|---|---|---|---|---|---|---|---|---|")
for l in open('gpurun_out/r2_q_sweep.jsonl'):
    d=json.loads(l); r=d['roofline']; c=d['config']
    ab=r.get('algorithmic_bytes') or 0
    gbs=ab/(r['kernel_ms']*1e-3)/1e9 if r['kernel_ms'] else 0
    tops=2*c['nq']*c['rows']*c['dim']/(r['kernel_ms']*1e-3)/1e12
    print(f"| {c['nq']} | {r['kernel'].split('::')[-1]} | {d['ms_per_step']:.3f} | {d['value']:.0f} | {r['kernel_ms']:.3f} | {ab/1e9:.2f} GB | {gbs:.0f} | {gbs/6447.5:.2f} | {tops:.0f} | {d['fallback_queries']} |")
PY
