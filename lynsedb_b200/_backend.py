"""Operator surface of ``lynse._backend`` for the distance + top-k path, served by the B200 library.

Same names, arguments and error behaviour as the reference module
(python/lynse/_backend.py:251-276 ``compute_distance`` / ``top_k_search``) and as
the pyo3 class ``lynse._core.FlatIndex`` (src/python/mod.rs:1942-2047), so the
reference's own operator tests (tests/standard_tests/test_backend.py) read the
same against this package.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Tuple

import numpy as np

from . import _native as N
from . import metrics as M
from .index import DeviceIndex


def rust_available() -> bool:  # name kept for drop-in callers; here: "is the native library usable"
    try:
        return N.device_count() > 0
    except Exception:
        return False


def compute_distance(a: np.ndarray, b: np.ndarray, metric: str = "ip") -> float:
    """Distance between two vectors (py_compute_distance, src/python/mod.rs:2161-2185)."""
    m = M.from_str(metric)
    if m is None:
        raise ValueError(f"Unknown metric: {metric}")
    a = np.ascontiguousarray(a, dtype=np.float32).ravel()
    b = np.ascontiguousarray(b, dtype=np.float32).ravel()
    if a.size != b.size:
        raise ValueError("Vector dimensions must match")
    if not M.accepts_dimension(m, a.size):
        raise ValueError("haversine requires two values in [longitude_degrees, latitude_degrees] order")
    out = C.c_float(0.0)
    N.check(N.lib().lb_compute_distance(N.fptr(a), N.fptr(b), a.size, m, C.byref(out)))
    return float(out.value)


def top_k_search(query: np.ndarray, candidates: np.ndarray, metric: str = "ip", k: int = 10
                 ) -> Tuple[np.ndarray, np.ndarray]:
    """(ids u32, distances f32), best first (py_top_k_search, src/python/mod.rs:2189-2223)."""
    m = M.from_str(metric)
    if m is None:
        raise ValueError(f"Unknown metric: {metric}")
    q = np.ascontiguousarray(query, dtype=np.float32).ravel()
    c = np.ascontiguousarray(candidates, dtype=np.float32)
    if c.ndim != 2:
        raise ValueError("candidates must be a 2-D array")
    n, dim = c.shape
    if q.size != dim:
        raise ValueError("Query dimension must match candidate dimension")
    if not M.accepts_dimension(m, dim):
        raise ValueError("haversine requires two values in [longitude_degrees, latitude_degrees] order")
    k = int(k)
    if k < 0:
        raise OverflowError("can't convert negative int to unsigned")
    kk = min(k, n)
    ids = np.empty(max(kk, 1), dtype=np.uint32)
    dists = np.empty(max(kk, 1), dtype=np.float32)
    cnt = C.c_uint32(0)
    if kk > 0:
        N.check(N.lib().lb_top_k_search(N.fptr(q), N.fptr(c), n, dim, kk, m, N.u32ptr(ids), N.fptr(dists), C.byref(cnt)))
    return ids[: cnt.value].copy(), dists[: cnt.value].copy()


class FlatIndex:
    """``lynse._core.FlatIndex(path, dim)``: raw row-major f32 rows, brute-force search.

    The rows live in HBM.  ``path`` is honoured the way ``FlatMmap::open`` does for reads: an
    existing raw little-endian f32 file is loaded; ``write`` replaces the contents (and rewrites
    the file so a reopen sees them).  Search returns raw u32 row indices.
    """

    def __init__(self, path: str, dim: int, device: int = 0):
        if int(dim) <= 0:
            raise ValueError("dimension must be positive")
        self._path = os.fspath(path) if path is not None else None
        self._dim = int(dim)
        self._device = device
        self._index = DeviceIndex(self._dim, "float32", device)
        if self._path and os.path.exists(self._path) and os.path.getsize(self._path) > 0:
            size = os.path.getsize(self._path)
            if size % (4 * self._dim) != 0:
                raise IOError(f"{self._path}: size {size} is not a multiple of the row width {4 * self._dim}")
            self._index.append(np.fromfile(self._path, dtype="<f4").reshape(-1, self._dim))

    def __len__(self) -> int:
        return len(self._index)

    @property
    def dim(self) -> int:
        return self._dim

    def write(self, data: np.ndarray) -> None:
        data = np.asarray(data)
        if data.ndim != 2 or data.shape[1] != self._dim:
            raise ValueError(f"data must have shape (n, {self._dim})")
        if not data.flags["C_CONTIGUOUS"]:
            raise ValueError("numpy array must be contiguous (C-order)")
        data = np.ascontiguousarray(data, dtype=np.float32)
        self._index.close()
        self._index = DeviceIndex(self._dim, "float32", self._device)
        self._index.append(data)
        if self._path:
            try:
                data.astype("<f4", copy=False).tofile(self._path)
            except OSError as e:
                raise IOError(str(e)) from e

    def search(self, query: np.ndarray, k: int = 10, metric: str = "ip") -> Tuple[np.ndarray, np.ndarray]:
        m = M.from_str(metric)
        if m is None:
            raise ValueError(f"Unknown metric: {metric}")
        rows, dists, counts = self._index.search(np.asarray(query, dtype=np.float32).reshape(1, -1), k, m)
        c = int(counts[0])
        return rows[0, :c].copy(), dists[0, :c].copy()

    def batch_search(self, queries: np.ndarray, k: int = 10, metric: str = "ip") -> List[Tuple[np.ndarray, np.ndarray]]:
        m = M.from_str(metric)
        if m is None:
            raise ValueError(f"Unknown metric: {metric}")
        queries = np.asarray(queries)
        if queries.ndim != 2 or not queries.flags["C_CONTIGUOUS"]:
            raise ValueError("numpy array must be contiguous (C-order)")
        rows, dists, counts = self._index.search(queries, k, m)
        return [(rows[i, : int(counts[i])].copy(), dists[i, : int(counts[i])].copy()) for i in range(rows.shape[0])]
