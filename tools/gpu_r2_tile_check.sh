#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_tile_scan.py -m gpu -q -x 2>&1 | tail -3
M=${1:-ip,l1,chebyshev,l2,cosine,bray_curtis}
for d in 256 128 512; do
  rows=$((1024000000 / d / 4 * 4))
  echo "== dim $d"; timeout 600 python tools/metric_sweep.py $rows $d 16,64,256 $M 2>&1 | grep -v f16
done
