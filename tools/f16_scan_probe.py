"""One-query exact scans over f32 rows and over binary16 rows of the same synthetic corpus (run on a B200)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lynsedb_b200 as L  # noqa: E402

n, dim = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000, 768
q = np.random.default_rng(3).random((4, dim), dtype=np.float32)
for dt in ("float32", "float16"):
    with L.DeviceIndex(dim, dt) as idx:
        idx.reserve(n)
        for lo in range(0, n, 500_000):
            idx.append_synthetic(min(500_000, n - lo), 42, lo)
        idx.set_plan("exact")
        idx.set_timing(True)
        for metric in ("ip", "l1", "cosine"):
            for nq in (1, 4):
                for _ in range(2):
                    idx.search(q[:nq], 10, metric)
                best = 1e9
                for _ in range(5):
                    idx.search(q[:nq], 10, metric)
                    best = min(best, idx.last_stats()["ms_dominant"])
                b = idx.last_stats()["algorithmic_bytes"]
                print(f"{dt:8s} {metric:7s} nq={nq}  {best:7.3f} ms  {b / best / 1e9:6.2f} TB/s  ({b / 1e9:.2f} GB)")
