"""DeviceIndex — one HBM-resident corpus shard behind the C ABI.

The host-side mirror of the reference's per-collection vector store for the
search path: ``VectorStore`` + ``FlatMmap`` (reference src/storage/vector_store.rs:972-1039,
src/storage/flat_mmap.rs:824-923).  Rows live in HBM; segment bookkeeping, side
structures (bf16 shadow, packed bits, Jensen-Shannon row stats) and the scan /
top-k kernels are native (lynsedb_b200/csrc).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _native as N
from . import metrics as M


class DeviceIndex:
    """Rows of one GPU.  ``dtype`` is ``"float32"``, ``"float16"`` (rows kept as IEEE binary16 in HBM — half the bytes of
    every scan — and decoded exactly on load; appends and queries stay float32) or ``"packed"`` (pre-packed one-bit rows,
    u64 words)."""

    _DTYPES = {"float32": N.LB_F32, "packed": N.LB_PACKED_U64, "float16": N.LB_F16}

    def __init__(self, dim: int, dtype: str = "float32", device: int = 0):
        if dtype not in self._DTYPES:
            raise ValueError(f"unsupported dtype: {dtype}")
        self._dtype = dtype
        self._dim = int(dim)
        self._device = int(device)
        self._h = C.c_void_p()
        N.check(N.lib().lb_index_create(C.byref(self._h), self._dim, self._DTYPES[dtype], self._device))
        import os

        target = os.environ.get("LYNSE_SEGMENT_TARGET_BYTES", "").strip()     # src/storage/vector_store.rs:225-229
        if target.isdigit() and int(target) > 0:
            N.check(N.lib().lb_index_set_segment_target(self._h, int(target)))

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

    def new_segment(self) -> None:
        """The next append opens a new segment even if it would fit the last one."""
        N.check(N.lib().lb_index_new_segment(self._h))

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


class ShardedDeviceIndex:
    """One vector store spread over several GPUs of this process: every SEGMENT of the reference's accounting
    (``VectorStore::append``, src/storage/vector_store.rs:379-445: an append joins the last segment while it fits 256 MiB,
    and is never split) lives whole on one device, segments go round-robin, and a search fans out to the devices in
    parallel and merges the per-device top-k blocks on the host by (score, global row) —
    ``VectorStore::merge_results`` (vector_store.rs:953-970; the cluster's shard merge, src/cluster.rs:327-393).
    Rows keep their global numbering (the order they were appended in); ``search`` returns global u32 rows, so the
    object is interchangeable with a ``DeviceIndex`` of the same rows.  Devices: ``LYNSE_B200_DEVICES=0,1,..`` or all."""

    SEGMENT_TARGET_BYTES = 256 << 20

    def __init__(self, dim: int, dtype: str = "float32", devices: Optional[Sequence[int]] = None):
        if devices is None or len(devices) == 0:
            raise ValueError("ShardedDeviceIndex needs at least one device")
        self._dim, self._dtype = int(dim), dtype
        self._devices = [int(d) for d in devices]
        self._shards = [DeviceIndex(dim, dtype, d) for d in self._devices]
        self._seg_target = self.SEGMENT_TARGET_BYTES
        self._segments: list = []            # reference accounting: rows per segment, oldest first
        self._seg_shard: list = []           # segment -> shard
        self._n = 0
        # per shard: blocks (local_start, global_start, rows) in local order, as arrays for searchsorted
        self._blocks = [([], [], []) for _ in self._shards]
        self._maps = [None] * len(self._shards)
        self._pool = None

    # -- lifetime ---------------------------------------------------------------------------------------------------
    def close(self) -> None:
        for s in self._shards:
            s.close()
        if self._pool is not None:
            self._pool.shutdown(wait=False)
            self._pool = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __len__(self) -> int:
        return self._n

    @property
    def dim(self) -> int:
        return self._dim

    @property
    def dtype(self) -> str:
        return self._dtype

    @property
    def devices(self):
        return list(self._devices)

    @property
    def n_words(self) -> int:
        return (self._dim + 63) // 64

    def shard_rows(self):
        return [len(s) for s in self._shards]

    def segments(self):
        return list(self._segments)

    def set_segment_target(self, n_bytes: int) -> None:
        self._seg_target = int(n_bytes)
        for s in self._shards:
            s.set_segment_target(n_bytes)

    def reserve(self, n_rows: int) -> None:
        per = (int(n_rows) + len(self._shards) - 1) // len(self._shards)
        for s in self._shards:
            s.reserve(per)

    # -- ingest -------------------------------------------------------------------------------------------------------
    def _row_bytes(self) -> int:
        return self.n_words * 8 if self._dtype == "packed" else self._dim * (2 if self._dtype == "float16" else 4)

    def append(self, rows: np.ndarray) -> None:
        rows = np.ascontiguousarray(rows, dtype=np.uint64 if self._dtype == "packed" else np.float32)
        rows = rows.reshape(1, -1) if rows.ndim == 1 else rows
        n = rows.shape[0]
        if n == 0:
            return
        rb = self._row_bytes()
        target = max(self._seg_target, rb)
        if self._segments and self._segments[-1] * rb + n * rb <= target:
            shard = self._seg_shard[-1]
            self._segments[-1] += n
        else:
            shard = len(self._segments) % len(self._shards)
            self._segments.append(n)
            self._seg_shard.append(shard)
            self._shards[shard].new_segment()
        idx = self._shards[shard]
        ls, gs, ln = self._blocks[shard]
        local = len(idx)
        idx.append(rows)
        if ls and ls[-1] + ln[-1] == local and gs[-1] + ln[-1] == self._n:
            ln[-1] += n
        else:
            ls.append(local)
            gs.append(self._n)
            ln.append(n)
        self._maps[shard] = None
        self._n += n

    def _shard_map(self, shard: int):
        if self._maps[shard] is None:
            ls, gs, ln = self._blocks[shard]
            self._maps[shard] = (np.asarray(ls, np.int64), np.asarray(gs, np.int64), np.asarray(ln, np.int64))
        return self._maps[shard]

    def _to_global(self, shard: int, local_rows: np.ndarray) -> np.ndarray:
        ls, gs, _ = self._shard_map(shard)
        lr = local_rows.astype(np.int64)
        b = np.searchsorted(ls, lr, side="right") - 1
        return (gs[b] + (lr - ls[b])).astype(np.uint32)

    def _split_allow(self, allow_bits: np.ndarray):
        """Global allow-bitset -> one local bitset per shard."""
        allow = np.ascontiguousarray(allow_bits, dtype=np.uint64)
        bits = np.unpackbits(allow.view(np.uint8), bitorder="little")[: self._n].astype(bool)
        out = []
        for shard, idx in enumerate(self._shards):
            ls, gs, ln = self._shard_map(shard)
            local = np.zeros(len(idx), dtype=bool)
            for a, g, m in zip(ls.tolist(), gs.tolist(), ln.tolist()):
                local[a:a + m] = bits[g:g + m]
            out.append(np.packbits(local, bitorder="little"))
        res = []
        for o, idx in zip(out, self._shards):
            words = np.zeros((len(idx) + 63) // 64 * 8, dtype=np.uint8)
            words[: o.size] = o
            res.append(words.view(np.uint64))
        return res

    def read_rows(self, first: int, n: int) -> np.ndarray:
        out = np.empty((n, self._dim), dtype=np.float32)
        for shard, idx in enumerate(self._shards):
            ls, gs, ln = self._shard_map(shard)
            for a, g, m in zip(ls.tolist(), gs.tolist(), ln.tolist()):
                lo, hi = max(g, first), min(g + m, first + n)
                if lo < hi:
                    out[lo - first:hi - first] = idx.read_rows(a + (lo - g), hi - lo)
        return out

    # -- search -------------------------------------------------------------------------------------------------------
    def prepare(self, metric) -> None:
        list(self._executor().map(lambda s: s.prepare(metric) if len(s) else None, self._shards))

    def set_plan(self, plan: str) -> None:
        for s in self._shards:
            s.set_plan(plan)

    def last_stats(self) -> dict:
        stats = [s.last_stats() for s in self._shards if len(s)]
        if not stats:
            return DeviceIndex.last_stats(self._shards[0])
        out = dict(stats[0])
        out["n_fallback"] = sum(s["n_fallback"] for s in stats)
        out["kernels_launched"] = sum(s["kernels_launched"] for s in stats)
        out["ms_dominant"] = max(s["ms_dominant"] for s in stats)
        out["ms_total"] = max(s["ms_total"] for s in stats)
        out["algorithmic_bytes"] = sum(s["algorithmic_bytes"] for s in stats)
        out["algorithmic_flops"] = sum(s["algorithmic_flops"] for s in stats)
        return out

    def _executor(self):
        if self._pool is None:
            from concurrent.futures import ThreadPoolExecutor

            self._pool = ThreadPoolExecutor(max_workers=len(self._shards), thread_name_prefix="lynse-shard")
        return self._pool

    def search(self, queries: np.ndarray, k: int, metric, allow_bits: Optional[np.ndarray] = None, pairwise: bool = False,
               f16_rows: bool = False) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """Same contract as ``DeviceIndex.search`` over the union of the shards (rows are global)."""
        k = int(k)
        q = np.ascontiguousarray(queries, dtype=np.uint64 if self._dtype == "packed" else np.float32)
        q = q.reshape(1, -1) if q.ndim == 1 else q
        nq = q.shape[0]
        active = [i for i, s in enumerate(self._shards) if len(s)]
        allows = self._split_allow(allow_bits) if allow_bits is not None else [None] * len(self._shards)

        def one(i):
            s = self._shards[i]
            r, d, c = s.search(q, min(k, len(s)), metric, allows[i], pairwise=pairwise, f16_rows=f16_rows)
            return i, r, d, c

        # ctypes releases the GIL for the duration of the native call: the devices run side by side
        parts = list(self._executor().map(one, active)) if len(active) > 1 else [one(i) for i in active]
        return merge_device_blocks([(self._to_global(i, r), d, c) for i, r, d, c in parts], nq, k, M.is_ascending(M.require(metric)))


def merge_device_blocks(parts, nq: int, k: int, ascending: bool):
    """Per-device ``(global rows[nq, k_i] u32, dists, counts)`` -> ``(rows[nq, k], dists, counts)`` by (score, global row):
    ``VectorStore::merge_results`` (src/storage/vector_store.rs:953-970), vectorised over the queries."""
    out_r = np.full((nq, k), N.ROW_NONE, dtype=np.uint32)
    out_d = np.full((nq, k), np.nan, dtype=np.float32)
    out_c = np.zeros(nq, dtype=np.uint32)
    if not parts or k == 0 or nq == 0:
        return out_r, out_d, out_c
    rows = np.concatenate([p[0] for p in parts], axis=1).astype(np.int64)
    dists = np.concatenate([p[1] for p in parts], axis=1)
    valid = np.concatenate([np.arange(p[0].shape[1])[None, :] < p[2][:, None].astype(np.int64) for p in parts], axis=1)
    key = np.where(ascending, dists, -dists).astype(np.float64) + 0.0       # -0.0 -> +0.0: equal scores tie on the row
    key = np.where(valid, key, np.inf)
    rows_k = np.where(valid, rows, np.iinfo(np.int64).max)
    o1 = np.argsort(rows_k, axis=1, kind="stable")
    key1 = np.take_along_axis(key, o1, axis=1)
    o2 = np.argsort(key1, axis=1, kind="stable")[:, :k]
    order = np.take_along_axis(o1, o2, axis=1)
    ok = np.take_along_axis(valid, order, axis=1)
    m = order.shape[1]
    out_r[:, :m] = np.where(ok, np.take_along_axis(rows, order, axis=1), N.ROW_NONE).astype(np.uint32)
    out_d[:, :m] = np.where(ok, np.take_along_axis(dists, order, axis=1), np.nan)
    out_c[:] = ok.sum(axis=1)
    return out_r, out_d, out_c


def visible_devices() -> list:
    """``LYNSE_B200_DEVICES`` (comma-separated device ordinals) or every CUDA device of the process."""
    import os

    env = os.environ.get("LYNSE_B200_DEVICES", "").strip()
    if env:
        return [int(x) for x in env.split(",") if x.strip() != ""]
    return list(range(N.device_count()))


def make_allow_bits(n_rows: int, allowed_rows) -> np.ndarray:
    """Row filter in the reference's BitSet layout (src/storage/bitset.rs): bit r of word r//64, LSB first."""
    words = np.zeros((n_rows + 63) // 64, dtype=np.uint64)
    rows = np.asarray(allowed_rows, dtype=np.uint64)
    np.bitwise_or.at(words, (rows >> np.uint64(6)).astype(np.int64), np.uint64(1) << (rows & np.uint64(63)))
    return words
