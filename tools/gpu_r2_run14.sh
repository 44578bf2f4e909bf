#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tile_scan.py -m gpu -q -x > gpurun_out/r2p_tile_pytest.log 2>&1; tail -5 gpurun_out/r2p_tile_pytest.log
M=${1:-ip,l2,l1,chebyshev}
for d in 128 256 512; do
  rows=$((1024000000 / d / 4 * 4))
  echo "== dim $d tile"; timeout 600 python tools/metric_sweep.py $rows $d 16,64,256 $M 2>&1 | grep -v f16 | tee gpurun_out/r2p_sweep_tile_$d.log
done
