"""DeviceIndex — one HBM-resident corpus shard behind the C ABI.

The host-side mirror of the reference's per-collection vector store for the
search path: ``VectorStore`` + ``FlatMmap`` (reference src/storage/vector_store.rs:972-1039,
src/storage/flat_mmap.rs:824-923).  Rows live in HBM; segment bookkeeping, side
structures (bf16 shadow, packed bits, Jensen-Shannon row stats) and the scan /
top-k kernels are native (lynsedb_b200/csrc).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _native as N
from . import metrics as M


class DeviceIndex:
    """Rows of one GPU.  ``dtype`` is ``"float32"`` or ``"packed"`` (pre-packed one-bit rows, u64 words)."""

    def __init__(self, dim: int, dtype: str = "float32", device: int = 0):
        if dtype not in ("float32", "packed"):
            raise ValueError(f"unsupported dtype: {dtype}")
        self._dtype = dtype
        self._dim = int(dim)
        self._device = int(device)
        self._h = C.c_void_p()
        N.check(N.lib().lb_index_create(C.byref(self._h), self._dim, N.LB_F32 if dtype == "float32" else N.LB_PACKED_U64,
                                        self._device))

    # -- lifetime -------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            N.lib().lb_index_destroy(self._h)
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

    # -- properties -------------------------------------------------------------
    @property
    def dim(self) -> int:
        return self._dim

    @property
    def device(self) -> int:
        return self._device

    @property
    def dtype(self) -> str:
        return self._dtype

    @property
    def n_words(self) -> int:
        return (self._dim + 63) // 64

    def __len__(self) -> int:
        return int(N.lib().lb_index_len(self._h))

    def segments(self):
        n = C.c_int(0)
        N.check(N.lib().lb_index_segments(self._h, None, 0, C.byref(n)))
        out = np.zeros(max(n.value, 1), dtype=np.uint64)
        N.check(N.lib().lb_index_segments(self._h, N.u64ptr(out), n.value, C.byref(n)))
        return [int(x) for x in out[: n.value]]

    # -- ingest -------------------------------------------------------------------
    def reserve(self, n_rows: int) -> None:
        N.check(N.lib().lb_index_reserve(self._h, int(n_rows)))

    def set_segment_target(self, n_bytes: int) -> None:
        N.check(N.lib().lb_index_set_segment_target(self._h, int(n_bytes)))

    def append(self, rows: np.ndarray) -> None:
        """One append == one ``VectorStore::append`` call (never split across segments)."""
        if self._dtype == "packed":
            words = np.ascontiguousarray(rows, dtype=np.uint64)
            words = words.reshape(1, -1) if words.ndim == 1 else words
            if words.shape[1] != self.n_words:
                raise ValueError(f"packed rows must have {self.n_words} words, got {words.shape[1]}")
            N.check(N.lib().lb_index_append_packed(self._h, N.u64ptr(words), words.shape[0]))
            return
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        rows = rows.reshape(1, -1) if rows.ndim == 1 else rows
        if rows.ndim != 2 or rows.shape[1] != self._dim:
            raise ValueError(f"Dimension mismatch: expected {self._dim}, got {rows.shape[-1]}")
        N.check(N.lib().lb_index_append_f32(self._h, N.fptr(rows), rows.shape[0]))

    def append_synthetic(self, n: int, seed: int, row_offset: int = 0) -> None:
        """Rows generated on the device; reproducible on the host with ``lynsedb_b200.synthetic``."""
        N.check(N.lib().lb_index_append_synthetic(self._h, int(n), int(seed), int(row_offset)))

    def read_rows(self, first: int, n: int) -> np.ndarray:
        out = np.empty((n, self._dim), dtype=np.float32)
        N.check(N.lib().lb_index_read_rows_f32(self._h, int(first), int(n), N.fptr(out)))
        return out

    # -- search ---------------------------------------------------------------------
    def prepare(self, metric) -> None:
        N.check(N.lib().lb_index_prepare(self._h, M.require(metric)))

    def set_plan(self, plan: str) -> None:
        N.check(N.lib().lb_index_set_plan(self._h, {"auto": N.LB_PLAN_AUTO, "exact": N.LB_PLAN_EXACT}[plan]))

    def set_timing(self, enabled: bool) -> None:
        N.check(N.lib().lb_index_set_timing(self._h, 1 if enabled else 0))

    def last_stats(self) -> dict:
        st = N.SearchStats()
        N.check(N.lib().lb_index_last_stats(self._h, C.byref(st)))
        return st.as_dict()

    def search(self, queries: np.ndarray, k: int, metric, allow_bits: Optional[np.ndarray] = None, pairwise: bool = False,
               f16_rows: bool = False) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """Batched top-k.  Returns ``(rows[nq,k] u32, dists[nq,k] f32, counts[nq] u32)``; entries past
        ``counts[q]`` hold row 0xFFFFFFFF.  Order per query: best score first, ties by ascending row
        (reference src/storage/vector_store.rs:953-970).  ``pairwise=True`` scores every row with the reference's
        single-pair kernels (``compute_distance_f32``; it only differs from the scan order for IP, in the last ulp).
        ``f16_rows=True`` is the search of a float16 collection's single-query and filtered paths: the rows hold binary16
        values and every pair goes through the reference's scalar ``compute_distance_f16`` kernels."""
        m = M.require(metric)
        k = int(k)
        if k < 0:
            raise ValueError("k must be non-negative")
        if self._dtype == "packed":
            q = np.ascontiguousarray(queries, dtype=np.uint64)
            q = q.reshape(1, -1) if q.ndim == 1 else q
            if q.shape[1] != self.n_words:
                raise ValueError(f"Dimension mismatch: expected {self.n_words} words, got {q.shape[1]}")
        else:
            q = np.ascontiguousarray(queries, dtype=np.float32)
            q = q.reshape(1, -1) if q.ndim == 1 else q
            if q.ndim != 2 or q.shape[1] != self._dim:
                raise ValueError(f"Dimension mismatch: expected {self._dim}, got {q.shape[-1]}")
        nq = q.shape[0]
        rows = np.empty((nq, max(k, 1)), dtype=np.uint32)[:, :k]
        dists = np.empty((nq, max(k, 1)), dtype=np.float32)[:, :k]
        rows = np.ascontiguousarray(rows)
        dists = np.ascontiguousarray(dists)
        counts = np.zeros(max(nq, 1), dtype=np.uint32)[:nq]
        if self._dtype == "packed":
            N.check(N.lib().lb_index_search_packed(self._h, m, N.u64ptr(q), nq, k, N.u32ptr(rows), N.fptr(dists),
                                                   N.u32ptr(counts)))
        else:
            ab, aw = None, 0
            if allow_bits is not None:
                allow = np.ascontiguousarray(allow_bits, dtype=np.uint64)
                ab, aw = N.u64ptr(allow), allow.size
            if pairwise and f16_rows:
                raise ValueError("pairwise and f16_rows are different scoring rules; pick one")
            fn = (N.lib().lb_index_search_pairwise if pairwise else
                  N.lib().lb_index_search_f16_rows if f16_rows else N.lib().lb_index_search)
            N.check(fn(self._h, m, N.fptr(q), nq, k, ab, aw, N.u32ptr(rows), N.fptr(dists), N.u32ptr(counts)))
        return rows, dists, counts

    def search_device(self, metric, d_queries: int, nq: int, k: int, d_rows: int, d_dists: int, d_counts: int) -> None:
        """Queries and results are device pointers (bench: inputs already resident in HBM)."""
        N.check(N.lib().lb_index_search_device(self._h, M.require(metric), C.c_void_p(d_queries), int(nq), int(k),
                                               C.c_void_p(d_rows), C.c_void_p(d_dists), C.c_void_p(d_counts)))


def make_allow_bits(n_rows: int, allowed_rows) -> np.ndarray:
    """Row filter in the reference's BitSet layout (src/storage/bitset.rs): bit r of word r//64, LSB first."""
    words = np.zeros((n_rows + 63) // 64, dtype=np.uint64)
    rows = np.asarray(allowed_rows, dtype=np.uint64)
    np.bitwise_or.at(words, (rows >> np.uint64(6)).astype(np.int64), np.uint64(1) << (rows & np.uint64(63)))
    return words
