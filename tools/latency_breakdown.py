"""Where a single-query search call spends its time (100k x 128, k=10): wall vs device, auto vs exact plan."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lynsedb_b200 as L

n, dim, k = 100_000, 128, 10
rng = np.random.default_rng(1)
data = rng.random((n, dim), dtype=np.float32)
qs = rng.random((300, dim), dtype=np.float32)
with L.DeviceIndex(dim) as idx:
    for lo in range(0, n, 10000):
        idx.append(data[lo:lo + 10000])
    idx.set_timing(True)
    for plan in ("auto", "exact"):
        idx.set_plan(plan)
        for metric in ("ip", "l1"):
            for q in qs[:20]:
                idx.search(q, k, metric)
            wall, dev, scan = [], [], []
            for q in qs[20:]:
                t0 = time.perf_counter()
                idx.search(q, k, metric)
                wall.append((time.perf_counter() - t0) * 1e6)
                dev.append(idx.last_stats()["ms_total"] * 1e3)
                scan.append(idx.last_stats()["ms_dominant"] * 1e3)
            st = idx.last_stats()
            print(f"plan {plan:5s} metric {metric}: wall median {np.median(wall):.1f} us, device {np.median(dev):.1f} us (dominant kernel {np.median(scan):.1f} us), kernels {st['kernels_launched']}, plan_used {st['plan_used']}")
