"""Timing probe: packed Hamming scan (C4 shape by default) through the host-buffer C ABI."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lynsedb_b200 as L
from lynsedb_b200 import synthetic

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 256
k = int(sys.argv[3]) if len(sys.argv) > 3 else 32
metric = sys.argv[4] if len(sys.argv) > 4 else "hamming"
dim = 1024
idx = L.DeviceIndex(dim, "packed", device=0)
idx.reserve(rows)
done = 0
while done < rows:
    m = min(1_000_000, rows - done)
    idx.append_synthetic(m, 42, done)
    done += m
q = synthetic.rows_packed(43, np.arange(nq), dim // 64)
idx.set_timing(True)
for it in range(3):
    t0 = time.perf_counter()
    r, d, c = idx.search(q, k, metric)
    dt = time.perf_counter() - t0
    st = idx.last_stats()
    print(f"rows {rows} nq {nq} k {k} {metric}: wall {dt*1e3:.1f} ms, kernel {st['ms_dominant']:.2f} ms, total dev {st['ms_total']:.2f} ms, "
          f"{rows*nq/ (st['ms_dominant']*1e-3)/1e12:.3f} Tpair/s, {st['algorithmic_bytes']/(st['ms_dominant']*1e-3)/1e9:.0f} GB/s")
print(r[0, :5], d[0, :5])
