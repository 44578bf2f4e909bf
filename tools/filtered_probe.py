"""Timing probe: a batch search under a row filter (allow-bitset), tensor-core plan against the exact scan."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lynsedb_b200 as L
from lynsedb_b200 import synthetic

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 768
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
density = float(sys.argv[4]) if len(sys.argv) > 4 else 0.5
idx = L.DeviceIndex(dim, device=0)
idx.reserve(rows)
done = 0
while done < rows:
    m = min(1_000_000, rows - done)
    idx.append_synthetic(m, 42, done)
    done += m
q = synthetic.rows_f32(43, np.arange(nq), dim)
rng = np.random.default_rng(0)
allow = L.make_allow_bits(rows, np.flatnonzero(rng.random(rows) < density))
res = {}
for plan in ("auto", "exact"):
    idx.set_plan(plan)
    idx.search(q, 10, "ip", allow)
    t0 = time.perf_counter()
    res[plan] = idx.search(q, 10, "ip", allow)
    dt = time.perf_counter() - t0
    st = idx.last_stats()
    print(f"{rows} x {dim}, {nq} queries, filter density {density}: plan {plan} (used {st['plan_used']}, fallbacks {st['n_fallback']}): {dt * 1e3:.1f} ms")
print("same rows:", np.array_equal(res["auto"][0], res["exact"][0]))
