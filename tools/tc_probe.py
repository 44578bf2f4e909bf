"""Diagnostic: pins the tcgen05 operand layouts of the coarse kernel with structured inputs (run on a B200)."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from lynsedb_b200 import _native as N  # noqa: E402


def scores(q, c):
    q = np.ascontiguousarray(q, np.float32)
    c = np.ascontiguousarray(c, np.float32)
    out = np.full((q.shape[0], c.shape[0]), np.nan, np.float32)
    st = N.lib().lb_debug_tc_scores(N.fptr(q), q.shape[0], N.fptr(c), c.shape[0], q.shape[1], N.fptr(out))
    if st != 0:
        print("  lb_debug_tc_scores failed:", N.last_error())
        return None
    return out


def main():
    print(N.device_info(0))
    for dim in (64, 128, 768):
        nq, n = 128, 128
        print(f"== dim {dim}: all ones (expect {dim} everywhere)")
        out = scores(np.ones((nq, dim)), np.ones((n, dim)))
        if out is not None:
            print("  min/max:", out.min(), out.max(), " first row:", out[0, :8])
        print(f"== dim {dim}: A one-hot at k = m % dim, B[n][k] = k+1 (expect out[m][*] = m % dim + 1)")
        a = np.zeros((nq, dim), np.float32)
        a[np.arange(nq), np.arange(nq) % dim] = 1.0
        b = np.tile(np.arange(1, dim + 1, dtype=np.float32), (n, 1))
        out = scores(a, b)
        if out is not None:
            want = (np.arange(nq) % dim + 1).astype(np.float32)
            bad = np.nonzero(out[:, 0] != want)[0]
            print("  mismatching lanes:", len(bad), " sample got:", out[:16, 0], " column spread:", np.ptp(out, axis=1).max())
        print(f"== dim {dim}: A all ones, B one-hot at k = n % dim scaled by n+1 (expect out[*][n] = n+1)")
        b = np.zeros((n, dim), np.float32)
        b[np.arange(n), np.arange(n) % dim] = np.arange(1, n + 1)
        out = scores(np.ones((nq, dim)), b)
        if out is not None:
            print("  got row0:", out[0, :16], " ok:", np.array_equal(out[0], np.arange(1, n + 1, dtype=np.float32)))
    rng = np.random.default_rng(0)
    q = rng.random((200, 200), dtype=np.float32) - 0.5
    c = rng.random((1000, 200), dtype=np.float32) - 0.5
    out = scores(q, c)
    if out is not None:
        print("== random 200x1000x200: max abs err vs f32 matmul:", np.abs(out - q @ c.T).max())


if __name__ == "__main__":
    main()
