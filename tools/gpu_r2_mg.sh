#!/bin/bash
# multi-GPU validation: 2 ranks, c2 (+ c5 weak extra) and c3
mkdir -p gpurun_out
N=${1:-2}
for w in c2 c3; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $w --steps 10 --warmup 3 > gpurun_out/r2e_${w}_${N}gpu.json 2> gpurun_out/r2e_${w}_${N}gpu.err
echo "== $w x$N rc=$?"; tail -c 2500 gpurun_out/r2e_${w}_${N}gpu.json; tail -3 gpurun_out/r2e_${w}_${N}gpu.err
done
