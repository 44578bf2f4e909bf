"""One small index, a few single-query searches (for an ncu launch list of the latency path)."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lynsedb_b200 as L

n, dim, k = 100_000, 128, 10
plan = sys.argv[1] if len(sys.argv) > 1 else "exact"
metric = sys.argv[2] if len(sys.argv) > 2 else "ip"
rng = np.random.default_rng(1)
data = rng.random((n, dim), dtype=np.float32)
qs = rng.random((8, dim), dtype=np.float32)
with L.DeviceIndex(dim) as idx:
    idx.append(data)
    idx.set_plan(plan)
    for q in qs:
        idx.search(q, k, metric)
