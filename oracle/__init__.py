"""ctypes wrapper over oracle/liblynse_oracle.so — the CPU restatement of the
reference's distance + top-k path.

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(lynsedb_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "liblynse_oracle.so"

METRICS = {
    "ip": 0, "l2": 1, "cosine": 2, "hamming": 3, "jaccard": 4, "l1": 5, "haversine": 6,
    "correlation": 7, "hellinger": 8, "wasserstein": 9, "dice": 10, "tanimoto": 11,
    "jensen_shannon": 12, "chebyshev": 13, "canberra": 14, "bray_curtis": 15,
}


def build(force: bool = False) -> Path:
    """Compile the oracle with the committed Makefile (g++ only; no reference sources)."""
    src = _HERE / "lynse_oracle.cpp"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            build()
        L = C.CDLL(str(_LIB_PATH))
        f32p, u32p, u64p = C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
        L.lo_compute_distance.restype = C.c_float
        L.lo_compute_distance.argtypes = [f32p, f32p, C.c_uint64, C.c_int]
        L.lo_inner_product_batch8_order.restype = C.c_float
        L.lo_inner_product_batch8_order.argtypes = [f32p, f32p, C.c_uint64]
        L.lo_probability_row_stats.restype = None
        L.lo_probability_row_stats.argtypes = [f32p, C.c_uint64, C.c_uint64, f32p]
        L.lo_jensen_shannon_precomputed.restype = C.c_float
        L.lo_jensen_shannon_precomputed.argtypes = [f32p, f32p, C.c_uint64, C.c_float, C.c_float, C.c_float]
        L.lo_jensen_shannon_precomputed_divergence.restype = C.c_float
        L.lo_jensen_shannon_precomputed_divergence.argtypes = [f32p, f32p, C.c_uint64, C.c_float, C.c_float, C.c_float]
        L.lo_pack_binary.restype = None
        L.lo_pack_binary.argtypes = [f32p, C.c_uint64, C.c_uint64, C.c_float, u64p]
        L.lo_packed_distance.restype = C.c_float
        L.lo_packed_distance.argtypes = [u64p, u64p, C.c_uint64, C.c_int]
        L.lo_packed_search.restype = C.c_uint32
        L.lo_packed_search.argtypes = [u64p, u64p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_int, u32p, f32p]
        L.lo_top_k_search.restype = C.c_uint32
        L.lo_top_k_search.argtypes = [f32p, f32p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_int, u32p, f32p]
        L.lo_flat_search.restype = C.c_uint32
        L.lo_flat_search.argtypes = [f32p, C.c_uint64, C.c_uint64, f32p, C.c_uint32, C.c_int, C.c_int, u32p, f32p]
        L.lo_store_batch_search.restype = None
        L.lo_store_batch_search.argtypes = [f32p, u64p, C.c_uint64, C.c_uint64, f32p, C.c_uint64, C.c_uint32,
                                            C.c_int, C.c_int, u64p, f32p, u32p]
        L.lo_packed_batch_search.restype = None
        L.lo_packed_batch_search.argtypes = [u64p, C.c_uint64, C.c_uint64, u64p, C.c_uint64, C.c_uint32, C.c_int,
                                             C.c_int, u64p, f32p, u32p]
        L.lo_kmeans_train.restype = C.c_uint32
        L.lo_kmeans_train.argtypes = [f32p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, f32p, u32p]
        L.lo_compute_distance_f16.restype = C.c_float
        L.lo_compute_distance_f16.argtypes = [f32p, f32p, C.c_uint64, C.c_int]
        L.lo_store_search_f16.restype = C.c_uint32
        L.lo_store_search_f16.argtypes = [f32p, u64p, C.c_uint64, C.c_uint64, f32p, C.c_uint32, C.c_int, C.c_int, u64p, f32p]
        L.lo_ivf_flat_search.restype = C.c_uint32
        L.lo_ivf_flat_search.argtypes = [f32p, C.c_uint64, C.c_uint64, f32p, C.c_uint32, u32p, f32p, C.c_uint32, C.c_uint32,
                                         C.c_int, u32p, f32p, u32p, u32p]
        L.lo_ivf_flat_routing_dims.restype = C.c_uint32
        L.lo_ivf_flat_routing_dims.argtypes = [f32p, C.c_uint64, C.c_uint32, u32p]
        L.lo_ivf_search.restype = C.c_uint32
        L.lo_ivf_search.argtypes = [f32p, C.c_uint64, C.c_uint64, f32p, C.c_uint32, u32p, f32p, C.c_uint32, C.c_uint32,
                                    C.c_int, u64p, u32p, f32p]
        L.lo_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def metric_id(metric) -> int:
    if isinstance(metric, int):
        return metric
    return METRICS[metric]


def host_threads() -> int:
    return os.cpu_count() or 1


def compute_distance(a, b, metric) -> float:
    a, b = _f32(a).ravel(), _f32(b).ravel()
    assert a.shape == b.shape
    return float(lib().lo_compute_distance(_p(a, C.c_float), _p(b, C.c_float), a.size, metric_id(metric)))


def inner_product_batch8_order(q, row) -> float:
    q, row = _f32(q).ravel(), _f32(row).ravel()
    return float(lib().lo_inner_product_batch8_order(_p(q, C.c_float), _p(row, C.c_float), q.size))


def probability_row_stats(rows):
    rows = _f32(rows)
    rows = rows.reshape(1, -1) if rows.ndim == 1 else rows
    out = np.empty((rows.shape[0], 2), dtype=np.float32)
    lib().lo_probability_row_stats(_p(rows, C.c_float), rows.shape[0], rows.shape[1], _p(out, C.c_float))
    return out


def jensen_shannon_precomputed(nq, cand, query_entropy, inv_mass, entropy, divergence=False) -> float:
    nq, cand = _f32(nq).ravel(), _f32(cand).ravel()
    fn = lib().lo_jensen_shannon_precomputed_divergence if divergence else lib().lo_jensen_shannon_precomputed
    return float(fn(_p(nq, C.c_float), _p(cand, C.c_float), nq.size, query_entropy, inv_mass, entropy))


def pack_binary(rows, threshold: float = 0.5):
    rows = _f32(rows)
    rows = rows.reshape(1, -1) if rows.ndim == 1 else rows
    words = (rows.shape[1] + 63) // 64
    out = np.zeros((rows.shape[0], words), dtype=np.uint64)
    lib().lo_pack_binary(_p(rows, C.c_float), rows.shape[0], rows.shape[1], threshold, _p(out, C.c_uint64))
    return out


def packed_distance(a, b, metric) -> float:
    a = np.ascontiguousarray(a, dtype=np.uint64).ravel()
    b = np.ascontiguousarray(b, dtype=np.uint64).ravel()
    return float(lib().lo_packed_distance(_p(a, C.c_uint64), _p(b, C.c_uint64), a.size, metric_id(metric)))


def packed_search(query_words, data_words, k, metric, n_threads=None):
    q = np.ascontiguousarray(query_words, dtype=np.uint64).ravel()
    d = np.ascontiguousarray(data_words, dtype=np.uint64)
    n, words = d.shape
    ids = np.empty(max(k, 1), dtype=np.uint32)
    dists = np.empty(max(k, 1), dtype=np.float32)
    cnt = lib().lo_packed_search(_p(q, C.c_uint64), _p(d, C.c_uint64), words, n, k, metric_id(metric),
                                 n_threads or host_threads(), _p(ids, C.c_uint32), _p(dists, C.c_float))
    return ids[:cnt].copy(), dists[:cnt].copy()


def top_k_search(query, candidates, metric="ip", k=10, n_threads=None):
    q, c = _f32(query).ravel(), _f32(candidates)
    n, dim = c.shape if c.ndim == 2 else (0, q.size)
    ids = np.empty(max(min(k, n), 1), dtype=np.uint32)
    dists = np.empty(max(min(k, n), 1), dtype=np.float32)
    cnt = lib().lo_top_k_search(_p(q, C.c_float), _p(c, C.c_float), n, dim, k, metric_id(metric),
                                n_threads or host_threads(), _p(ids, C.c_uint32), _p(dists, C.c_float))
    return ids[:cnt].copy(), dists[:cnt].copy()


def flat_search(corpus, query, k, metric, n_threads=None):
    """FlatMmap::search on one segment -> (u32 rows, f32 dists)."""
    c, q = _f32(corpus), _f32(query).ravel()
    n, dim = c.shape
    ids = np.empty(max(min(k, n), 1), dtype=np.uint32)
    dists = np.empty(max(min(k, n), 1), dtype=np.float32)
    cnt = lib().lo_flat_search(_p(c, C.c_float), n, dim, _p(q, C.c_float), k, metric_id(metric),
                               n_threads or host_threads(), _p(ids, C.c_uint32), _p(dists, C.c_float))
    return ids[:cnt].copy(), dists[:cnt].copy()


def store_batch_search(corpus, queries, k, metric, segment_rows=None, n_threads=None):
    """VectorStore::search per query (sequential over queries, segments merged by (score,row)).

    Returns (ids[nq,k] u64, dists[nq,k] f32, counts[nq] u32); entries past counts[q] are undefined.
    """
    c, q = _f32(corpus), _f32(queries)
    q = q.reshape(1, -1) if q.ndim == 1 else q
    n, dim = c.shape
    segs = np.asarray([n] if segment_rows is None else segment_rows, dtype=np.uint64)
    assert int(segs.sum()) == n
    nq = q.shape[0]
    ids = np.zeros((nq, max(k, 1)), dtype=np.uint64)
    dists = np.zeros((nq, max(k, 1)), dtype=np.float32)
    counts = np.zeros(nq, dtype=np.uint32)
    lib().lo_store_batch_search(_p(c, C.c_float), _p(segs, C.c_uint64), segs.size, dim, _p(q, C.c_float), nq, k,
                                metric_id(metric), n_threads or host_threads(), _p(ids, C.c_uint64),
                                _p(dists, C.c_float), _p(counts, C.c_uint32))
    return ids, dists, counts


def packed_batch_search(data_words, query_words, k, metric, n_threads=None):
    d = np.ascontiguousarray(data_words, dtype=np.uint64)
    q = np.ascontiguousarray(query_words, dtype=np.uint64)
    q = q.reshape(1, -1) if q.ndim == 1 else q
    n, words = d.shape
    nq = q.shape[0]
    ids = np.zeros((nq, max(k, 1)), dtype=np.uint64)
    dists = np.zeros((nq, max(k, 1)), dtype=np.float32)
    counts = np.zeros(nq, dtype=np.uint32)
    lib().lo_packed_batch_search(_p(d, C.c_uint64), words, n, _p(q, C.c_uint64), nq, k, metric_id(metric),
                                 n_threads or host_threads(), _p(ids, C.c_uint64), _p(dists, C.c_float),
                                 _p(counts, C.c_uint32))
    return ids, dists, counts


def kmeans_train(data, n_clusters, metric, max_iter=20):
    """kmeans::train_for_metric (src/index/kmeans.rs:74-139) -> (centroids[nc,dim] f32, assignments[n] u32)."""
    d = _f32(data)
    n, dim = d.shape
    cent = np.zeros((max(n_clusters, 1), dim), dtype=np.float32)
    assign = np.zeros(max(n, 1), dtype=np.uint32)
    nc = lib().lo_kmeans_train(_p(d, C.c_float), n, dim, n_clusters, max_iter, metric_id(metric), _p(cent, C.c_float),
                               _p(assign, C.c_uint32))
    return cent[:nc].copy(), assign[:n].copy()


def ivf_search(data, centroids, assignments, query, k, nprobe, metric, allow_bits=None):
    """IVFIndex::search (src/index/ivf.rs:181-348, no quantizer) for one query -> (u32 rows, f32 dists)."""
    d, c, q = _f32(data), _f32(centroids), _f32(query).ravel()
    a = np.ascontiguousarray(assignments, dtype=np.uint32)
    n, dim = d.shape
    ids = np.empty(max(k, 1), dtype=np.uint32)
    dists = np.empty(max(k, 1), dtype=np.float32)
    ab = None
    if allow_bits is not None:
        allow = np.ascontiguousarray(allow_bits, dtype=np.uint64)
        ab = _p(allow, C.c_uint64)
    cnt = lib().lo_ivf_search(_p(d, C.c_float), n, dim, _p(c, C.c_float), c.shape[0], _p(a, C.c_uint32), _p(q, C.c_float),
                              k, nprobe, metric_id(metric), ab, _p(ids, C.c_uint32), _p(dists, C.c_float))
    return ids[:cnt].copy(), dists[:cnt].copy()


def ivf_flat_search(data, centroids, assignments, query, k, nprobe, metric, return_probes=False):
    """IvfFlatMmap::search (src/storage/ivf_flat_mmap.rs:225-300) for one query -> (u32 original ids, f32 dists)."""
    d, c, q = _f32(data), _f32(centroids), _f32(query).ravel()
    a = np.ascontiguousarray(assignments, dtype=np.uint32)
    n, dim = d.shape
    ids = np.empty(max(k, 1), dtype=np.uint32)
    dists = np.empty(max(k, 1), dtype=np.float32)
    probes = np.zeros(max(c.shape[0], 1), dtype=np.uint32)
    n_probes = np.zeros(1, dtype=np.uint32)
    cnt = lib().lo_ivf_flat_search(_p(d, C.c_float), n, dim, _p(c, C.c_float), c.shape[0], _p(a, C.c_uint32), _p(q, C.c_float),
                                   k, nprobe, metric_id(metric), _p(ids, C.c_uint32), _p(dists, C.c_float),
                                   _p(probes, C.c_uint32), _p(n_probes, C.c_uint32))
    if return_probes:
        return ids[:cnt].copy(), dists[:cnt].copy(), probes[: int(n_probes[0])].copy()
    return ids[:cnt].copy(), dists[:cnt].copy()


def ivf_flat_routing_dims(centroids):
    """select_routing_dims (src/storage/ivf_flat_mmap.rs:312-345)."""
    c = _f32(centroids)
    out = np.zeros(16, dtype=np.uint32)
    cnt = lib().lo_ivf_flat_routing_dims(_p(c, C.c_float), c.shape[1], c.shape[0], _p(out, C.c_uint32))
    return out[:cnt].copy()


def compute_distance_f16(query, row, metric):
    """compute_distance_f16 (src/distance/mod.rs:217-237): f32 query x binary16 row in the scalar kernels' order.
    ``row`` is rounded to binary16 first (what the F16 storage dtype holds)."""
    q = _f32(query).ravel()
    r = np.ascontiguousarray(np.asarray(row, dtype=np.float32).astype(np.float16).astype(np.float32)).ravel()
    return float(lib().lo_compute_distance_f16(_p(q, C.c_float), _p(r, C.c_float), q.size, metric_id(metric)))


def store_search_f16(data, query, k, metric, segments=None, n_threads=None):
    """VectorStore::search of a float16 collection for one query -> (u64 rows, f32 dists).  ``data`` is rounded to
    binary16 first; ``segments`` are the segment row counts (default: one segment)."""
    d = np.ascontiguousarray(_f32(data).astype(np.float16).astype(np.float32))
    q = _f32(query).ravel()
    n, dim = d.shape
    seg = np.ascontiguousarray(segments if segments is not None else [n], dtype=np.uint64)
    ids = np.empty(max(k, 1), dtype=np.uint64)
    dists = np.empty(max(k, 1), dtype=np.float32)
    cnt = lib().lo_store_search_f16(_p(d, C.c_float), _p(seg, C.c_uint64), seg.size, dim, _p(q, C.c_float), k, metric_id(metric),
                                    n_threads or host_threads(), _p(ids, C.c_uint64), _p(dists, C.c_float))
    return ids[:cnt].copy(), dists[:cnt].copy()
