#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/f16q.py <<'PY'
import sys
sys.path.insert(0, '.')
import numpy as np
import lynsedb_b200 as L
dt = sys.argv[1]
n, dim = 4_000_000, 768
q = np.random.default_rng(3).random((1, dim), dtype=np.float32)
with L.DeviceIndex(dim, dt) as idx:
    idx.reserve(n)
    for lo in range(0, n, 500_000):
        idx.append_synthetic(500_000, 42, lo)
    idx.set_plan("exact")
    for _ in range(3):
        idx.search(q, 10, "ip")
PY
for dt in float16 float32; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_stream_tma -s 2 -c 1 -f -o gpurun_out/r2_scan_$dt python /tmp/f16q.py $dt > gpurun_out/r2_ncu_scan_$dt.log 2>&1; echo "$dt rc=$?"
ncu -i gpurun_out/r2_scan_$dt.ncu-rep --page raw --csv > gpurun_out/r2_scan_${dt}_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_scan_$dt.ncu-rep --page source --csv > gpurun_out/r2_scan_${dt}_source.csv 2>/dev/null
rm -f gpurun_out/r2_scan_$dt.ncu-rep
done
