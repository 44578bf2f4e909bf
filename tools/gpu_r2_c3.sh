#!/bin/bash
# C3 experiments: where does the epilogue time go
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" LYNSE_B200_TC_TRACE=1 LYNSE_B200_TC_PROF=1 timeout 300 python bench.py --workload c3 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_c3_$name.json 2> gpurun_out/r2b_c3_$name.err; echo "== $name"; python - <<PY
import json
d=json.loads(open('gpurun_out/r2b_c3_$name.json').read().strip().splitlines()[-1])
print('ms/step %.3f kernel %.3f fb %d' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['fallback_queries']), d['verified']['ids_exact_vs_exact_plan'])
PY
grep -E "per tile|scan \(" gpurun_out/r2b_c3_$name.err | tail -2; }
run default A=1
run nohits LYNSE_B200_TC_HITS=0
run epi1 LYNSE_B200_TC_EPI=1
run epi1_nacc2 LYNSE_B200_TC_EPI=1 LYNSE_B200_TC_NACC=2
run noscan LYNSE_B200_TC_DEBUG=4
run noread LYNSE_B200_TC_DEBUG=2
run bn64 LYNSE_B200_TC_BN=64
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; tail -5 gpurun_out/r2b_pytest.log
