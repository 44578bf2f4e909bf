#!/bin/bash
# usage: tools/run_bench_line.sh <bench.py args...>  -> one compact line (nq, ms/step, TFLOP/s or GB/s, kernel ms, partitions, clocks)
python bench.py "$@" --no-cpu-baseline 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=d['roofline']
print(d['config']['nq'], 'ms/step %.3f' % d['ms_per_step'], 'e2e %.3f' % d['e2e']['ms_per_step'], r['bound'], '%.1f' % r['achieved'], 'kernel_ms %.3f' % r['kernel_ms'], 'P', d['partitions'], 'fb', d['fallback_queries'], d['clocks']['sm_mhz'], d['clocks']['reasons'], d['verified'])
"
