"""Timing probe: IVF build + search (1M x 128, 256 lists, nprobe 10 by default) through the public classes."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lynsedb_b200 as L

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 128
nlist = int(sys.argv[3]) if len(sys.argv) > 3 else 256
nprobe = int(sys.argv[4]) if len(sys.argv) > 4 else 10
rng = np.random.default_rng(0)
data = rng.random((n, dim), dtype=np.float32)
queries = rng.random((200, dim), dtype=np.float32)
idx = L.DeviceIndex(dim)
idx.append(data)
t0 = time.perf_counter()
ivf = L.IVFIndex(idx, "ip", n_clusters=nlist)
print(f"build {n} x {dim}, {nlist} lists: {time.perf_counter() - t0:.2f} s")
for q in queries[:10]:
    ivf.search(q, 10, nprobe)
lat = []
for q in queries[10:110]:
    t0 = time.perf_counter()
    ivf.search(q, 10, nprobe)
    lat.append((time.perf_counter() - t0) * 1e6)
print(f"IVFIndex.search nprobe {nprobe}: median {np.median(lat):.0f} us, p90 {np.percentile(lat, 90):.0f} us")
t0 = time.perf_counter()
rows, d, c = ivf.search(queries, 10, nprobe)
print(f"batch of {len(queries)}: {(time.perf_counter() - t0) * 1e3:.1f} ms")
flat_rows, _, _ = idx.search(queries, 10, "ip")
recall = np.mean([len(set(rows[i].tolist()) & set(flat_rows[i].tolist())) / 10 for i in range(len(queries))])
print(f"recall@10 vs flat: {recall:.3f}")
