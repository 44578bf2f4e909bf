"""Timing probe: exact scan of every f32 metric at a few batch sizes (device time of the scan kernel, GB/s of rows read)."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lynsedb_b200 as L
from lynsedb_b200 import synthetic

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 256
batches = [int(x) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["1", "8", "64"])]
metrics = sys.argv[4].split(",") if len(sys.argv) > 4 else ["ip", "l2", "cosine", "l1", "chebyshev", "canberra", "bray_curtis", "correlation",
                                                          "hellinger", "wasserstein", "jensen_shannon", "hamming"]
idx = L.DeviceIndex(dim, device=0)
idx.reserve(rows)
done = 0
while done < rows:
    m = min(1_000_000, rows - done)
    idx.append_synthetic(m, 42, done)
    done += m
idx.set_plan("exact")
idx.set_timing(True)
gb = rows * dim * 4 / 1e9
print(f"corpus {rows} x {dim} f32 = {gb:.2f} GB")
for metric in metrics:
    idx.prepare(metric)
    for nq in batches:
        q = synthetic.rows_f32(43, np.arange(nq), dim)
        best = 1e30
        for it in range(3):
            idx.search(q, 10, metric)
            best = min(best, idx.last_stats()["ms_dominant"])
        for f16 in ((False, True) if metric not in ("hamming",) and nq == 1 else (False,)):
            if f16:
                best = 1e30
                for it in range(3):
                    idx.search(q, 10, metric, f16_rows=True)
                    best = min(best, idx.last_stats()["ms_dominant"])
            print(f"{metric:15s} nq {nq:3d} {'f16-order' if f16 else 'flat     '}: scan {best:8.3f} ms = {gb / best:7.2f} TB/s-equivalent" if False else
                  f"{metric:15s} nq {nq:3d} {'f16-order' if f16 else 'flat     '}: scan {best:8.3f} ms  {gb / (best * 1e-3) / 1e3:6.2f} TB/s of rows")
