"""Diagnostic: tcgen05.mma cycles per instruction for several (N, accumulators, A source) shapes (run on a B200)."""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from lynsedb_b200 import _native as N  # noqa: E402

lib = N.lib()
iters = 4096
print("a_src  N  n_acc grid  cyc/MMA(total) cyc/MMA(issue)  ideal")
for grid in (1, 148):
    for a_in_tmem in (1, 0):
        for n in (64, 128, 256):
            for n_acc in (1, 2, 4, 8):
                if (384 if a_in_tmem else 0) + n * n_acc > 512:
                    continue
                t, i = C.c_uint64(0), C.c_uint64(0)
                st = lib.lb_debug_mma_rate(n, n_acc, iters, a_in_tmem, grid, C.byref(t), C.byref(i))
                if st != 0:
                    print("fail", N.last_error())
                    continue
                print(f"{'tmem' if a_in_tmem else 'smem'} {n:4d} {n_acc:5d} {grid:4d}  {t.value / iters:10.1f} {i.value / iters:12.1f}  {n / 2:6.0f}")
