"""IVFIndex — inverted-file index over the rows of one DeviceIndex.

Host-side mirror of the reference's ``IVFIndex`` (src/index/ivf.rs:131-348; no quantizer): k-means centroids and
inverted lists are built on the GPU (``lb_ivf_train``: src/index/kmeans.rs restated as CUDA kernels), a search
ranks the centroids with the routing metric, gathers the ``nprobe`` nearest lists and scores every candidate with
the exact per-pair kernels.  ``nprobe >= n_centroids`` is an exact search.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _native as N
from . import metrics as M
from .index import DeviceIndex

DEFAULT_N_CLUSTERS = 256   # IndexBuildOptions default (src/index/mod.rs)
DEFAULT_NPROBE = 32        # build-time default the search falls back to when nprobe == 0 (src/index/ivf.rs:192-196)
KMEANS_MAX_ITER = 20       # IVFIndex::build (src/index/ivf.rs:163-170)


class IVFIndex:
    def __init__(self, index: DeviceIndex, metric, n_clusters: int = DEFAULT_N_CLUSTERS, nprobe: int = DEFAULT_NPROBE,
                 centroids: Optional[np.ndarray] = None, assignments: Optional[np.ndarray] = None,
                 max_iter: int = KMEANS_MAX_ITER):
        self._index = index
        self._metric = M.require(metric)
        self._nprobe = int(nprobe)
        self._h = C.c_void_p()
        if centroids is not None:
            cent = np.ascontiguousarray(centroids, dtype=np.float32)
            assign = np.ascontiguousarray(assignments, dtype=np.uint32)
            if cent.ndim != 2 or cent.shape[1] != index.dim or assign.shape != (len(index),):
                raise ValueError("centroids must be [n_centroids, dim] and assignments [len(index)]")
            N.check(N.lib().lb_ivf_create(index._h, self._metric, N.fptr(cent), cent.shape[0], N.u32ptr(assign), C.byref(self._h)))
        else:
            N.check(N.lib().lb_ivf_train(index._h, self._metric, int(n_clusters), int(max_iter), C.byref(self._h)))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            N.lib().lb_ivf_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def n_centroids(self) -> int:
        nc, n = C.c_uint32(0), C.c_uint64(0)
        N.check(N.lib().lb_ivf_info(self._h, C.byref(nc), C.byref(n)))
        return nc.value

    @property
    def n_rows(self) -> int:
        nc, n = C.c_uint32(0), C.c_uint64(0)
        N.check(N.lib().lb_ivf_info(self._h, C.byref(nc), C.byref(n)))
        return n.value

    def centroids(self) -> np.ndarray:
        out = np.empty((self.n_centroids, self._index.dim), dtype=np.float32)
        N.check(N.lib().lb_ivf_centroids(self._h, N.fptr(out)))
        return out

    def assignments(self) -> np.ndarray:
        out = np.empty(self.n_rows, dtype=np.uint32)
        N.check(N.lib().lb_ivf_assignments(self._h, N.u32ptr(out)))
        return out

    def search(self, queries: np.ndarray, k: int, nprobe: int = 0, allow_bits: Optional[np.ndarray] = None
               ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """``(rows[nq,k] u32, dists[nq,k] f32, counts[nq] u32)``; ``nprobe == 0`` uses the build-time default."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        q = q.reshape(1, -1) if q.ndim == 1 else q
        if q.shape[1] != self._index.dim:
            raise ValueError(f"Dimension mismatch: expected {self._index.dim}, got {q.shape[1]}")
        nq, k = q.shape[0], int(k)
        np_eff = max(int(nprobe) if int(nprobe) > 0 else self._nprobe, 1)
        rows = np.empty((nq, max(k, 1)), dtype=np.uint32)[:, :k]
        dists = np.empty((nq, max(k, 1)), dtype=np.float32)[:, :k]
        rows, dists = np.ascontiguousarray(rows), np.ascontiguousarray(dists)
        counts = np.zeros(max(nq, 1), dtype=np.uint32)[:nq]
        ab, aw = None, 0
        if allow_bits is not None:
            allow = np.ascontiguousarray(allow_bits, dtype=np.uint64)
            ab, aw = N.u64ptr(allow), allow.size
        N.check(N.lib().lb_ivf_search(self._h, N.fptr(q), nq, k, np_eff, ab, aw, N.u32ptr(rows), N.fptr(dists), N.u32ptr(counts)))
        return rows, dists, counts

    def flat_search(self, queries: np.ndarray, k: int, nprobe: int, metric) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """The standalone IVF_FLAT search rule (``IvfFlatMmap::search``, src/storage/ivf_flat_mmap.rs:225-300):
        partitions chosen under ``metric`` (routing-dimension shortlist for inner product), no corpus fallback."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        q = q.reshape(1, -1) if q.ndim == 1 else q
        if q.shape[1] != self._index.dim:
            raise ValueError(f"query dimension mismatch: expected {self._index.dim}, got {q.shape[1]}")
        nq, k = q.shape[0], int(k)
        rows = np.ascontiguousarray(np.empty((nq, max(k, 1)), dtype=np.uint32)[:, :k])
        dists = np.ascontiguousarray(np.empty((nq, max(k, 1)), dtype=np.float32)[:, :k])
        counts = np.zeros(max(nq, 1), dtype=np.uint32)[:nq]
        N.check(N.lib().lb_ivf_flat_search(self._h, N.fptr(q), nq, k, max(int(nprobe), 0), M.require(metric), N.u32ptr(rows),
                                           N.fptr(dists), N.u32ptr(counts)))
        return rows, dists, counts
