#!/bin/bash
# row-tile scan: CTA shape (warps) x CTAs per SM, L1 and IP at 64 queries
for d in 256 128; do
rows=$((1024000000 / d / 4 * 4))
for cfg in "0 2" "4 2" "2 2" "2 3" "2 4" "8 2" "4 3"; do
set -- $cfg
out=$(LYNSE_B200_SCAN_TILE_NW=$1 LYNSE_B200_SCAN_TILE_CTAS=$2 timeout 300 python tools/metric_sweep.py $rows $d 64 l1,ip,chebyshev 2>&1 | grep "nq  64" | awk '{print $1, $7}' | tr '\n' ' ')
echo "dim $d nw $1 ctas $2: $out"
done
done
