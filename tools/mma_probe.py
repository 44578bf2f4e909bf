"""Diagnostic: tcgen05.mma cycles per instruction for several (N, accumulators, A source) shapes (run on a B200)."""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from lynsedb_b200 import _native as N  # noqa: E402

lib = N.probe_lib()   # liblynse_b200_probe.so (include/lynse_b200_probe.h)
iters = 4096
print("kind  a_src  N  n_acc grid  cyc/MMA(total) cyc/MMA(issue)  ideal   (kind::f16: M=128 x N x K=16; kind::i8: M=128 x N x K=32)")
for grid in (1, 148):
    for i8 in (0, 1):
        for a_in_tmem in (1, 0):
            for n in (64, 128, 256):
                for n_acc in (1, 2):
                    t, i = C.c_uint64(0), C.c_uint64(0)
                    st = lib.lb_debug_mma_rate(n, n_acc, iters, a_in_tmem, i8, grid, C.byref(t), C.byref(i))
                    if st != 0:
                        continue  # shape not instantiated
                    print(f"{'i8 ' if i8 else 'f16'}  {'tmem' if a_in_tmem else 'smem'} {n:4d} {n_acc:5d} {grid:4d}  {t.value / iters:10.1f} {i.value / iters:12.1f}  {n / 2:6.0f}")
