"""Single-query latency through the public API (the way benchmarks/flat_search_bench.py drives the reference):
100k x 128 f32, FLAT-IP, k=10, one query per call, 20 warm-ups + 200 trials; prints median / p10 / p90 in microseconds."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lynsedb_b200 as L

n, dim, k = 100_000, 128, 10
rng = np.random.default_rng(42)
data = rng.random((n, dim), dtype=np.float32)
queries = rng.random((220, dim), dtype=np.float32)
queries[0] = data[0]
with L.VectorDBClient() as client:
    coll = client.create_collection("bench", "flat", dim=dim, default_index="FLAT-IP")
    for lo in range(0, n, 10_000):
        coll.add(vectors=data[lo:lo + 10_000], batch_size=10_000)
    coll.commit()
    for mode, call in (("collection.search", lambda q: coll.search(q, k=k)),
                       ("DeviceIndex.search", lambda q: coll._store.search(q.reshape(1, -1), k, "ip"))):
        for q in queries[:20]:
            call(q)
        ts = []
        for q in queries[20:]:
            t0 = time.perf_counter()
            call(q)
            ts.append((time.perf_counter() - t0) * 1e6)
        ts = np.asarray(ts)
        print(f"{mode}: median {np.median(ts):.1f} us, p10 {np.percentile(ts, 10):.1f}, p90 {np.percentile(ts, 90):.1f} "
              f"({1e6 / np.median(ts):.0f} queries/s single-stream)")
    assert coll.search(data[0], k=1).ids[0] == 0
