#!/bin/bash
mkdir -p gpurun_out
timeout 2000 python -m pytest tests -m gpu -q -x > gpurun_out/r2n_pytest.log 2>&1; tail -4 gpurun_out/r2n_pytest.log
bash tools/gpu_exp.sh r2n c3 10 default: 2>&1 | grep -E "QPS|=="
for rows in 1250000 10000000; do
python bench.py --workload c2 --rows $rows --steps 20 --warmup 5 --no-cpu-baseline --no-api-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2 rows $rows: ms/step %.3f e2e %.3f kernel %.3f fb %d' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['fallback_queries']))"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:finalize -s 2 -c 2 --csv --log-file gpurun_out/r2n_fin.csv python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline --no-api-e2e --verify-queries 4 > /dev/null 2>&1
echo "finalize c3: $(grep finalize gpurun_out/r2n_fin.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:finalize -s 2 -c 2 --csv --log-file gpurun_out/r2n_fin2.csv python bench.py --workload c2 --rows 1250000 --steps 2 --warmup 1 --no-cpu-baseline --no-api-e2e --verify-queries 4 > /dev/null 2>&1
echo "finalize c2: $(grep finalize gpurun_out/r2n_fin2.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"
