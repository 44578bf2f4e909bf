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
from .ivf import IVFIndex


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

    The rows live in HBM.  ``path`` is honoured the way ``FlatMmap::open`` / ``write`` do
    (src/storage/flat_mmap.rs:187-221, :305-345): an existing raw little-endian f32 file is loaded,
    ``write`` APPENDS rows to the index and to the file, so a reopen sees them.  Search returns raw
    u32 row indices.
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
        if data.shape[0] == 0:
            return
        if self._path:
            try:
                with open(self._path, "ab") as f:
                    f.write(data.astype("<f4", copy=False).tobytes())
            except OSError as e:
                raise IOError(str(e)) from e
        self._index.append(data)

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


def _ivf_meta_path(path: str) -> str:
    """``data_path.with_extension("ivf_meta.bin")`` (src/storage/ivf_flat_mmap.rs:138)."""
    head, tail = os.path.split(path)
    stem = tail.rsplit(".", 1)[0] if "." in tail.lstrip(".") else tail
    return os.path.join(head, stem + ".ivf_meta.bin")


class IvfFlatIndex:
    """``lynse._core.IvfFlatIndex`` (src/python/mod.rs:2049-2156) = ``IvfFlatMmap`` (src/storage/ivf_flat_mmap.rs).

    ``build`` clusters the rows with the shared L2 k-means on the GPU (``kmeans::train_l2``), stores the rows
    partition-contiguous in ``path`` (raw little-endian f32) and centroids, partition offsets and original row
    positions in ``<stem>.ivf_meta.bin`` (``save_metadata``, ivf_flat_mmap.rs:450-483).  ``search`` scores every row of
    the ``nprobe`` partitions nearest to the query under the requested metric and returns ORIGINAL row positions.
    """

    def __init__(self, index: DeviceIndex, ivf: IVFIndex, dim: int, n_partitions: int):
        self._index, self._ivf, self._dim, self._n_partitions = index, ivf, int(dim), int(n_partitions)

    # -- construction -------------------------------------------------------------------------------------------
    @staticmethod
    def build(path: str, data: np.ndarray, dim: int, n_partitions: int = 256, n_iters: int = 20, metric: str = "ip",
              device: int = 0) -> "IvfFlatIndex":
        if M.from_str(metric) is None:
            raise ValueError(f"Unknown metric: {metric}")
        data = np.asarray(data)
        if data.ndim != 2 or data.shape[1] != int(dim):
            got = data.shape[1] if data.ndim == 2 else data.shape
            raise ValueError(f"data dimension mismatch: expected {dim}, got {got}")
        if not data.flags["C_CONTIGUOUS"]:
            raise ValueError("numpy array must be contiguous (C-order)")
        dim, n_partitions, n = int(dim), int(n_partitions), data.shape[0]
        if dim == 0:
            raise IOError("IVF dimension must be greater than zero")
        if n_partitions == 0:
            raise IOError("IVF partition count must be greater than zero")
        if n < n_partitions:
            raise IOError("IVF requires at least as many vectors as partitions")
        if n > 0xFFFFFFFF:
            raise IOError("IVF vector count exceeds the u32 ID capacity")
        data = np.ascontiguousarray(data, dtype=np.float32)
        index = DeviceIndex(dim, "float32", device)
        index.append(data)
        # the partitioning is always the L2 Voronoi cells; the metric only matters at search time (ivf_flat_mmap.rs:96-101)
        ivf = IVFIndex(index, M.L2, n_clusters=n_partitions, max_iter=int(n_iters))
        self = IvfFlatIndex(index, ivf, dim, ivf.n_centroids)
        if path:
            self._save(os.fspath(path), data)
        return self

    def _save(self, path: str, data: np.ndarray) -> None:
        assign = self._ivf.assignments()
        order = np.argsort(assign, kind="stable").astype(np.uint32)           # partition-contiguous, row order inside
        sizes = np.bincount(assign, minlength=self._n_partitions).astype(np.uint64)
        offsets = np.zeros(self._n_partitions + 1, dtype="<u8")
        offsets[1:] = np.cumsum(sizes)
        try:
            with open(path, "wb") as f:
                f.write(data[order].astype("<f4", copy=False).tobytes())
            with open(_ivf_meta_path(path), "wb") as f:
                f.write(np.array([self._dim, data.shape[0], self._n_partitions], dtype="<u8").tobytes())
                f.write(self._ivf.centroids().astype("<f4", copy=False).tobytes())
                f.write(offsets.tobytes())
                f.write(order.astype("<u4", copy=False).tobytes())
        except OSError as e:
            raise IOError(str(e)) from e

    @staticmethod
    def open(path: str, dim: int, device: int = 0) -> "IvfFlatIndex":
        """Load an index written by ``build`` (or by the reference: same two files, ``load_metadata`` :485-530)."""
        path, dim = os.fspath(path), int(dim)
        try:
            with open(_ivf_meta_path(path), "rb") as f:
                head = np.frombuffer(f.read(24), dtype="<u8")
                if head.size != 3:
                    raise IOError("truncated IVF metadata header")
                dim_loaded, n, n_partitions = (int(x) for x in head)
                if dim != dim_loaded:
                    raise IOError(f"IVF dimension mismatch: index has {dim_loaded}, requested {dim}")
                centroids = np.frombuffer(f.read(4 * n_partitions * dim), dtype="<f4").reshape(n_partitions, dim)
                offsets = np.frombuffer(f.read(8 * (n_partitions + 1)), dtype="<u8")
                original = np.frombuffer(f.read(4 * n), dtype="<u4")
                if offsets.size != n_partitions + 1 or original.size != n:
                    raise IOError("truncated IVF metadata")
            reordered = np.fromfile(path, dtype="<f4")
        except OSError as e:
            raise IOError(str(e)) from e
        if reordered.size != n * dim:
            raise IOError(f"{path}: expected {n} rows of {dim} values")
        # back to original row positions so results need no translation
        rows = np.empty((n, dim), dtype=np.float32)
        rows[original] = reordered.reshape(n, dim)
        assign = np.empty(n, dtype=np.uint32)
        assign[original] = np.repeat(np.arange(n_partitions, dtype=np.uint32), np.diff(offsets).astype(np.int64))
        index = DeviceIndex(dim, "float32", device)
        index.append(rows)
        ivf = IVFIndex(index, M.L2, centroids=np.ascontiguousarray(centroids, dtype=np.float32), assignments=assign)
        return IvfFlatIndex(index, ivf, dim, n_partitions)

    # -- accessors ----------------------------------------------------------------------------------------------
    def __len__(self) -> int:
        return len(self._index)

    @property
    def dim(self) -> int:
        return self._dim

    @property
    def n_partitions(self) -> int:
        return self._n_partitions

    def suggested_nprobe(self) -> int:
        """~8 % of the partitions, at least 10 (ivf_flat_mmap.rs:213-217)."""
        return min(max(self._n_partitions // 13, 10), self._n_partitions)

    # -- search -------------------------------------------------------------------------------------------------
    def search(self, query: np.ndarray, k: int = 10, nprobe: int = 10, metric: str = "ip") -> Tuple[np.ndarray, np.ndarray]:
        m = M.from_str(metric)
        if m is None:
            raise ValueError(f"Unknown metric: {metric}")
        q = np.ascontiguousarray(query, dtype=np.float32).ravel()
        if q.size != self._dim:
            raise ValueError(f"query dimension mismatch: expected {self._dim}, got {q.size}")
        rows, dists, counts = self._ivf.flat_search(q.reshape(1, -1), int(k), int(nprobe), m)
        c = int(counts[0])
        return rows[0, :c].copy(), dists[0, :c].copy()

    def close(self) -> None:
        self._ivf.close()
        self._index.close()
