#!/bin/bash
# Final round-2 evidence after the helper-warp kernel: ncu --set full of the C2 coarse kernel (10M rows and the 1.25M-row
# shard), launch lists, bench lines of every workload, the reference arm.
mkdir -p gpurun_out
B="--steps 1 --warmup 1 --no-cpu-baseline --no-api-e2e --verify-queries 2"
cap() { # name bench-args
  name=$1; shift
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:coarse_pair_kernel -s 2 -c 1 -f -o gpurun_out/r2f_$name python bench.py "$@" $B > gpurun_out/r2f_ncu_$name.log 2>&1; echo "$name rc=$?"
  python tools/ncu_summary.py full gpurun_out/r2f_$name.ncu-rep gpurun_out/r2f_${name}_ncu.json
  rm -f gpurun_out/r2f_$name.ncu-rep
}
cap c2 --workload c2
cap c2shard --workload c2 --rows 1250000
for spec in "c2:--rows 1250000:c2shard" "c2::c2"; do
  w=${spec%%:*}; rest=${spec#*:}; extra=${rest%%:*}; tag=${rest#*:}
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_$tag.csv python bench.py --workload $w $extra --steps 2 --warmup 1 --no-cpu-baseline --no-api-e2e --verify-queries 4 > gpurun_out/r2f_launches_$tag.log 2>&1
  echo "== launches $tag rc=$?"
done
for w in c1 c2 c3 c4; do
  timeout 900 python bench.py --workload $w > gpurun_out/r2f_bench_$w.json 2> gpurun_out/r2f_bench_$w.err; echo "bench $w rc=$?"
done
timeout 900 python bench.py --impl reference > gpurun_out/r2f_bench_c2_reference.json 2> gpurun_out/r2f_bench_ref.err; echo "reference rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python - <<'PY'
import json
for w in ('c1','c2','c3','c4'):
    try:
        d=json.loads([l for l in open(f'gpurun_out/r2f_bench_{w}.json') if l.startswith('{')][-1]); a=d.get('api_e2e') or {}; r=d['roofline']
        print(w, 'QPS %.0f e2e %.0f api %s ms %.3f kernel %.3f frac %.3f fb %d verified %s cpu %s' % (d['value'], d['e2e']['value'], a.get('value'), d['ms_per_step'], r['kernel_ms'], r['frac'], d['fallback_queries'], (d['verified'] or {}).get('ids_exact_vs_exact_plan'), (d.get('cpu_baseline') or {}).get('value')))
    except Exception as e: print(w, 'no line', e)
print(open('gpurun_out/r2f_bench_c2_reference.json').read()[-600:])
PY
