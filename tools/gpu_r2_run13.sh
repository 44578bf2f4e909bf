#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tile_scan.py -m gpu -q -x > gpurun_out/r2o_tile_pytest.log 2>&1; tail -15 gpurun_out/r2o_tile_pytest.log
M=ip,l2,cosine,l1,chebyshev,canberra,bray_curtis
for d in 128 256 512; do
  rows=$((1024000000 / d / 4 * 4))
  echo "== dim $d tile"; timeout 600 python tools/metric_sweep.py $rows $d 16,64,256 $M 2>&1 | grep -v f16 | tee gpurun_out/r2o_sweep_tile_$d.log
done
echo "== dim 256 streaming (tile off)"; LYNSE_B200_SCAN_TILE=0 timeout 600 python tools/metric_sweep.py 4000000 256 16,64 $M 2>&1 | grep -v f16 | tee gpurun_out/r2o_sweep_stream_256.log
