"""VectorDBClient / Database / Collection — LynseDB's Python object model for the search path, on the B200 library.

Keeps the names, arguments and return types of the reference for what lies on the path
(python/lynse/__init__.py:12-323 ``VectorDBClient``; python/lynse/api/local_client.py:35-276 ``LocalClient``,
:278-1420 ``LocalCollection``: ``add`` / ``commit`` / ``build_index`` / ``search`` / ``batch_search`` / ``delete`` /
``restore`` / ``shape`` / ``index_mode``), and restates the engine's result post-processing around the scan
(src/engine.rs:4718-4833): ``search_k = k + |tombstones|``, the scan of the un-flushed pending rows
(``pending_search`` :3310-3360), ``merge_row_results`` (:3362-3419), row -> external id, ``filter_tombstoned_limit``
(:3286-3308).  Rows live in HBM (``DeviceIndex``); the database is in-memory — persistence, WAL, metadata SQL,
snapshots, HTTP and the other index families are outside this package.
"""
from __future__ import annotations

import threading
from typing import Any, Callable, Dict, Iterable, List, Optional, Sequence, Union

import numpy as np

from . import metrics as M
from .index import DeviceIndex, ShardedDeviceIndex, make_allow_bits, visible_devices
from .ivf import DEFAULT_N_CLUSTERS, DEFAULT_NPROBE, IVFIndex
from .result_view import ResultView

PENDING_FLUSH_ROWS = 10_000          # src/engine.rs:93-94
PENDING_FLUSH_BYTES = 32 << 20
MAX_DATABASES = 64                   # python/lynse/__init__.py:128
TENSOR_PLAN_MAX_K = 256              # largest k the tensor-core plan of the library serves
NATIVE_MAX_K = 2048                  # largest k of any native search (lb_index_search)
KNOWN_BUILD_KEYS = frozenset({"n_clusters", "n_centroids", "m", "ef_construction", "ef_search", "max_level", "r", "l",
                              "alpha", "max_degree", "nprobe", "replica_count"})  # python/lynse/_index_build.py:49-66
_DOMAIN_FREE = (M.IP, M.L2, M.COSINE, M.HAMMING, M.JACCARD)


def _index_family(mode: str) -> str:
    upper = mode.upper()
    for fam in ("DISKANN", "HNSW", "SPANN", "IVF", "FLAT"):
        if upper.startswith(fam):
            return fam
    return "FLAT"


class _IdMap:
    """row <-> external id (``row_to_user_id``, src/engine.rs:3071-3073).  While the ids are exactly 0, 1, 2, ... in row
    order (what ``add`` assigns by default) nothing is stored: a 10M-row collection costs no Python objects.  The first
    id that breaks the pattern materialises a list and a dict."""

    def __init__(self):
        self.n = 0
        self.ids: Optional[list] = None       # None = identity
        self.rows: Optional[dict] = None
        self.max_int = -1
        self.all_int = True
        self._np: Optional[np.ndarray] = None

    def __len__(self) -> int:
        return self.n

    def _materialise(self) -> None:
        if self.ids is None:
            self.ids = list(range(self.n))
            self.rows = {i: i for i in range(self.n)}

    def extend(self, ext: list) -> None:
        m = len(ext)
        if self.ids is None and m and ext[0] == self.n and ext[-1] == self.n + m - 1 and \
                all(type(e) is int for e in ext) and ext == list(range(self.n, self.n + m)):
            self.n += m
            self.max_int = self.n - 1
            return
        self._materialise()
        for e in ext:
            self.rows[e] = self.n
            self.ids.append(e)
            self.n += 1
            if isinstance(e, (int, np.integer)) and not isinstance(e, bool):
                if e > self.max_int:
                    self.max_int = int(e)
            else:
                self.all_int = False
        self._np = None

    def extend_identity(self, m: int) -> None:
        """m more rows whose ids continue the 0, 1, 2, ... pattern."""
        if self.ids is None:
            self.n += m
            self.max_int = self.n - 1
        else:
            self.extend(list(range(self.n, self.n + m)))

    def replace(self, ids: list) -> None:
        self.__init__()
        self.extend(list(ids))

    def __contains__(self, ident) -> bool:
        if self.ids is None:
            return isinstance(ident, (int, np.integer)) and not isinstance(ident, bool) and 0 <= ident < self.n
        return ident in self.rows

    def row_of(self, ident) -> int:
        return int(ident) if self.ids is None else self.rows[ident]

    def id_of(self, row: int):
        return int(row) if self.ids is None else self.ids[int(row)]

    def array(self) -> Optional[np.ndarray]:
        """row -> id as int64 (``None`` when some id is not an integer)."""
        if not self.all_int:
            return None
        if self._np is None or len(self._np) != self.n:
            self._np = np.arange(self.n, dtype=np.int64) if self.ids is None else np.asarray(self.ids, dtype=np.int64)
        return self._np

    @property
    def identity(self) -> bool:
        return self.ids is None


class Collection:
    """``LocalCollection`` for the search path.  ``where`` accepts ``None``, a callable over the row's field dict, or a
    dict of field == value conditions (the reference's SQL ``where`` strings need its metadata engine, out of scope)."""

    def __init__(self, name: str, dim: Optional[int] = None, *, dtypes: str = "float32", default_index: Optional[str] = "FLAT-IP",
                 description: Optional[str] = None, device: int = 0, devices: Optional[Sequence[int]] = None):
        if dtypes not in ("float32", "float16"):
            raise ValueError(f"unsupported dtypes: {dtypes!r}")
        # float16 collections: vectors are rounded to IEEE binary16 at write time exactly as the reference encodes them
        # (src/storage/dtype.rs:60-67, round-to-nearest-even) and kept DECODED (f32) in HBM — decoding is exact, so the
        # arithmetic below sees the reference's values.  Unfiltered FLAT batch_search takes the reference's F16 batch
        # path, which decodes the store and runs the f32 kernels (src/engine.rs:5440-5474); search() and every filtered
        # search take FlatMmap::search / search_filtered on F16 storage, i.e. the scalar f32-query x f16-row kernels
        # (src/distance/simd.rs:805-1092) — `lb_index_search_f16_rows`.  The half-width HBM layout is not built.
        self._dtypes = dtypes
        self.name = name
        self.description = description
        self._dim = int(dim) if dim else None
        self._device = device
        # several devices: the store's segments are spread over them (ShardedDeviceIndex) and every FLAT search fans out
        self._devices = [int(d) for d in devices] if devices else [int(device)]
        self._default_index = default_index
        self._index_mode: Optional[str] = None
        self._metric = M.IP
        self._lock = threading.RLock()
        self._store: Optional[DeviceIndex] = None
        self._ivf: Optional[IVFIndex] = None
        self._ivf_params: Dict[str, int] = {}
        self._pending: List[np.ndarray] = []
        self._pending_rows = 0
        self._pending_index: Optional[DeviceIndex] = None   # the un-flushed rows in HBM, rebuilt when they change
        self._ids = _IdMap()                     # row <-> external id (engine.rs:3071-3073)
        self._fields: Dict[int, dict] = {}
        self._tombstones: set = set()
        self._dead_rows: Optional[np.ndarray] = None   # sorted rows of the tombstoned ids (rebuilt when the set changes)
        self._dead_masks: Dict[Any, np.ndarray] = {}    # cached allow-bitsets with the tombstoned rows cleared
        self.COMMIT_FLAG = True

    # ------------------------------------------------------------------ basics
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self) -> None:
        with self._lock:
            if self._ivf is not None:
                self._ivf.close()
                self._ivf = None
            self._drop_pending_index()
            if self._store is not None:
                self._store.close()
                self._store = None

    def exists(self) -> bool:
        return True

    @property
    def shape(self):
        return (len(self._ids), self._dim or 0)

    @property
    def index_mode(self) -> Optional[str]:
        return self._index_mode

    def vector_dtype(self) -> str:
        return self._dtypes

    def max_id(self) -> int:
        return self._ids.max_int

    def is_id_exists(self, id) -> bool:
        return id in self._ids and id not in self._tombstones

    def list_deleted_ids(self) -> list:
        return sorted(self._tombstones, key=lambda x: (str(type(x)), x))

    def stats(self) -> dict:
        return {"name": self.name, "rows": len(self._ids), "dim": self._dim, "index_mode": self._index_mode,
                "pending_rows": self._pending_rows, "deleted": len(self._tombstones),
                "segments": self._store.segments() if self._store is not None else []}

    # ------------------------------------------------------------------ ingest
    def add(self, ids=None, *, vectors=None, documents=None, embed_func=None, fields=None, batch_size: int = 1000,
            wire_dtype: str = "float32"):
        del wire_dtype
        if documents is not None or embed_func is not None:
            raise NotImplementedError("document embedding is outside this package's scope; pass vectors")
        if not isinstance(batch_size, int) or batch_size <= 0:
            raise ValueError("batch_size must be a positive integer")
        if vectors is None:
            raise ValueError("add() requires vectors or documents")
        vec = np.asarray(vectors, dtype=np.float32)
        if vec.ndim == 1:
            vec = vec.reshape(1, -1)
        elif vec.ndim != 2:
            raise ValueError("vectors must be a 1D vector or a 2D matrix")
        vec = np.ascontiguousarray(vec, dtype=np.float32)
        if self._dtypes == "float16":
            vec = vec.astype(np.float16).astype(np.float32)
        n = vec.shape[0]
        if n == 0:
            raise ValueError("vectors cannot be empty")
        with self._lock:
            if self._dim is None:
                self._dim = int(vec.shape[1])
            if vec.shape[1] != self._dim:
                raise ValueError(f"Dimension mismatch: expected {self._dim}, got {vec.shape[1]}")
            single = False
            if ids is None:
                start = self.max_id() + 1
                ext = list(range(start, start + n))
            else:
                single = not isinstance(ids, (list, tuple, np.ndarray))
                ext = [ids] if single else [i.item() if isinstance(i, np.generic) else i for i in ids]
                if len(ext) != n:
                    raise ValueError(f"ids length ({len(ext)}) must match vectors row count ({n})")
                if len(set(ext)) != len(ext):
                    raise ValueError("duplicate ids in one add() call")
                for e in ext:
                    if e in self._ids:
                        raise ValueError(f"id {e!r} already exists; use upsert")
            field_list = None
            if fields is not None:
                field_list = [fields] if isinstance(fields, dict) else list(fields)
                if len(field_list) != n:
                    raise ValueError(f"fields length ({len(field_list)}) must match vectors row count ({n})")
            base = len(self._ids)
            self._ids.extend(ext)
            if field_list is not None:
                for j in range(n):
                    if field_list[j] is not None:
                        self._fields[base + j] = dict(field_list[j])
            for s in range(0, n, batch_size):   # one add_items call per batch, as the reference client does
                chunk = vec[s:s + batch_size]
                self._pending.append(chunk)
                self._pending_rows += chunk.shape[0]
                self._drop_pending_index()
                if self._pending_rows >= PENDING_FLUSH_ROWS or self._pending_rows * self._dim * 4 >= PENDING_FLUSH_BYTES:
                    self._flush_pending()
            if self._ivf is not None:
                self._ivf.close()
                self._ivf = None        # lists are rebuilt over the new rows on the next search
            self.COMMIT_FLAG = False
        self._maybe_build_default_index()
        if ids is None:
            return ext[0] if n == 1 else ext
        return ext[0] if single else ext

    def _ensure_store(self):
        if self._store is None:
            if self._dim is None:
                raise ValueError("collection dimension is not known yet")
            # float16 collections keep their rows as binary16 in HBM (half the bytes of every scan); the values were rounded
            # at write time, so narrowing them is exact
            row_dtype = "float16" if self._dtypes == "float16" else "float32"
            if len(self._devices) > 1:
                self._store = ShardedDeviceIndex(self._dim, row_dtype, self._devices)
            else:
                self._store = DeviceIndex(self._dim, row_dtype, self._devices[0])
            if self._dtypes == "float16":
                # segment boundaries never reach a float16 collection's results: search() scores rows one by one, and the
                # batch path scans `read_all_f32()` — the whole store as ONE array (src/engine.rs:5448-5453), so the
                # inner-product small-segment rule is evaluated on the total row count.  One segment says the same.
                self._store.set_segment_target(1 << 62)
        return self._store

    def _drop_pending_index(self) -> None:
        if self._pending_index is not None:
            self._pending_index.close()
            self._pending_index = None

    def _pending_store(self) -> DeviceIndex:
        """The un-flushed rows as a device index: one upload per change of the buffer, not one per query."""
        if self._pending_index is None:
            block = self._pending[0] if len(self._pending) == 1 else np.concatenate(self._pending, axis=0)
            self._pending_index = DeviceIndex(self._dim, "float32", self._device)
            self._pending_index.append(block)
        return self._pending_index

    def _flush_pending(self) -> None:
        """One flush == one ``VectorStore::append`` (never split across segments)."""
        if not self._pending:
            return
        block = self._pending[0] if len(self._pending) == 1 else np.concatenate(self._pending, axis=0)
        self._ensure_store().append(block)
        self._pending = []
        self._pending_rows = 0
        self._drop_pending_index()
        self._dead_masks.clear()
        if self._ivf is not None:
            # the inverted lists cover the rows of the store at build time: rebuilt over the grown store on the next search
            self._ivf.close()
            self._ivf = None

    def commit(self) -> None:
        with self._lock:
            self._flush_pending()
            self.COMMIT_FLAG = True

    def attach_store(self, store) -> None:
        """Adopt an already filled ``DeviceIndex`` / ``ShardedDeviceIndex`` (rows generated on the device, bench.py): ids
        are the row numbers.  The collection must be empty; it owns the store from here on."""
        with self._lock:
            if len(self._ids) or self._store is not None:
                raise ValueError("attach_store needs an empty collection")
            if self._dim is not None and store.dim != self._dim:
                raise ValueError(f"Dimension mismatch: expected {self._dim}, got {store.dim}")
            self._dim = store.dim
            self._store = store
            self._ids.extend_identity(len(store))
        self._maybe_build_default_index()

    def insert_session(self) -> "InsertSession":
        """``with coll.insert_session() as s: s.add(...)`` — adds are applied and committed when the block ends
        (python/lynse/execution_layer/session.py)."""
        return InsertSession(self)

    def compact(self) -> int:
        """Physically drop the tombstoned rows and rebuild the store (``Collection::compact``); returns the rows removed.
        Rows keep their relative order, so ties still resolve the way they did."""
        with self._lock:
            self._flush_pending()
            dead_rows = set(self._ids.row_of(t) for t in self._tombstones)
            if not dead_rows:
                return 0
            n = len(self._ids)
            live = np.asarray([r for r in range(n) if r not in dead_rows], dtype=np.int64)
            vectors = self._store.read_rows(0, n)[live] if n else np.empty((0, self._dim or 0), np.float32)
            ids = [self._ids.id_of(int(r)) for r in live]
            fields = {new: self._fields[int(old)] for new, old in enumerate(live) if int(old) in self._fields}
            if self._ivf is not None:
                self._ivf.close()
                self._ivf = None
            self._store.close()
            self._store = None
            self._ids.replace(ids)
            self._fields = fields
            self._tombstones = set()
            self._dead_rows = None
            self._dead_masks.clear()
            if len(ids):
                self._ensure_store().append(np.ascontiguousarray(vectors, dtype=np.float32))
                if self._index_mode and not self._index_mode.startswith("IVF"):
                    self._store.prepare(self._metric)
            return len(dead_rows)

    def load_lynsedb_directory(self, collection_path, dtype: Optional[str] = None) -> int:
        """Load the vectors of an existing LynseDB collection directory (vector_manifest.json + segment files + id_map.bin,
        src/storage/vector_store.rs:24-60, :157-243) into this collection; one append per segment file.  Returns the rows added."""
        from . import storage_reader as R

        with self._lock:
            if self._dim is None:
                raise ValueError("collection dimension must be set to read a raw vector store")
            if len(self._ids):
                raise ValueError("load_lynsedb_directory needs an empty collection")
            dtype = dtype or self._dtypes
            segments, id_map_path = R.read_manifest(collection_path, self._dim, dtype)
            total = sum(r for _, r in segments)
            ids = R.read_id_map(id_map_path, total)
            ext = ids.tolist() if ids is not None else list(range(total))
            if len(set(ext)) != len(ext):
                raise IOError("id map holds duplicate ids")
            for path, rows in segments:
                if rows:
                    self._ensure_store().append(R.read_segment(path, rows, self._dim, dtype))
            self._ids.replace(ext)
        self._maybe_build_default_index()
        return total

    def search_range(self, vector, threshold: float, max_results: int = 1000) -> ResultView:
        """``ResultView`` (python/lynse/api/local_client.py:1370-1397) of every live row within ``threshold`` (<= for distances, >= for IP), best first, at most
        ``max_results`` — ``Collection::search_range`` (src/engine.rs:6410-6483): per-pair ``compute_distance_f32`` order."""
        q = np.ascontiguousarray(vector, dtype=np.float32).reshape(1, -1)
        max_results = int(max_results)
        with self._lock:
            if max_results <= 0 or self._dim is None or not len(self._ids):
                idx_type, dist_name = M.parse_index_mode(self._index_mode or "FLAT-IP")
                return ResultView(ids=np.empty(0, np.int64), distances=np.empty(0, np.float32), k=0, distance=dist_name,
                                  index=idx_type, result_type="search")
            if q.shape[1] != self._dim:
                raise ValueError(f"Dimension mismatch: expected {self._dim}, got {q.shape[1]}")
            self._flush_pending()
            want = min(max_results + len(self._tombstones), len(self._store))
            allow = None
            if want > NATIVE_MAX_K:
                # many deleted rows: mask them out and ask for max_results (the over-fetch only exists to survive the
                # tombstone filter, src/engine.rs:6410-6483)
                allow = self._live_mask(len(self._store), None)
                want = min(max_results, len(self._store))
            if want > NATIVE_MAX_K:
                raise ValueError("search_range is limited to max_results <= 2048 on this path")
            rows, dists, counts = self._store.search(q, want, self._metric, allow, pairwise=True)
            asc = M.is_ascending(self._metric)
            ids, out = [], []
            for r, d in zip(rows[0, :int(counts[0])].tolist(), dists[0, :int(counts[0])].tolist()):
                e = self._ids.id_of(r)
                if e in self._tombstones or not (d <= threshold if asc else d >= threshold):
                    continue
                ids.append(e)
                out.append(d)
                if len(ids) == max_results:
                    break
            all_int = all(isinstance(i, (int, np.integer)) for i in ids)
            idx_type, dist_name = M.parse_index_mode(self._index_mode or "FLAT-IP")
            return ResultView(ids=np.asarray(ids, dtype=np.int64 if all_int else object), distances=np.asarray(out, dtype=np.float32),
                              k=len(ids), distance=dist_name, index=idx_type, result_type="search")

    flush = commit

    def delete(self, ids) -> int:
        """Tombstone ids (soft delete; ``search`` asks for ``k + |tombstones|`` and filters)."""
        with self._lock:
            ids = ids if isinstance(ids, (list, tuple, np.ndarray, set)) else [ids]
            n = 0
            for i in ids:
                i = i.item() if isinstance(i, np.generic) else i
                if i in self._ids and i not in self._tombstones:
                    self._tombstones.add(i)
                    n += 1
            if n:
                self._dead_rows = None
                self._dead_masks.clear()
            return n

    def restore(self, ids) -> int:
        with self._lock:
            ids = ids if isinstance(ids, (list, tuple, np.ndarray, set)) else [ids]
            n = 0
            for i in ids:
                i = i.item() if isinstance(i, np.generic) else i
                if i in self._tombstones:
                    self._tombstones.discard(i)
                    n += 1
            if n:
                self._dead_rows = None
                self._dead_masks.clear()
            return n

    # ------------------------------------------------------------------ index
    def _maybe_build_default_index(self) -> None:
        if self._index_mode is None and self._default_index and len(self._ids):
            self.build_index(self._default_index)

    def build_index(self, index_mode: str = "FLAT-IP", **kwargs) -> None:
        """``build_index(index_mode, **kwargs)`` (local_client.py:701-845; src/engine.rs:4500-4660)."""
        for key, value in kwargs.items():
            if value is not None and key not in KNOWN_BUILD_KEYS:
                raise ValueError(f"unknown index build parameter {key!r}; supported keys: {', '.join(sorted(KNOWN_BUILD_KEYS))}")
        mode = str(index_mode).upper()
        family = _index_family(mode)
        # HNSW / DiskANN / SPANN modes are accepted and served by the exact FLAT scan of their metric: on a B200 the
        # full-bandwidth scan is faster than the reference's graph walks and returns the exact neighbours the graphs
        # approximate (the index_mode string and ResultView.index_type keep the requested family)
        if mode in ("FLAT", "IVF", "HNSW", "DISKANN", "SPANN"):
            raise ValueError(f"unknown index type '{index_mode}'")       # bare family names are rejected (src/index/mod.rs:827-836)
        metric = M.from_index_mode(mode)
        if metric is None:
            raise ValueError(f"unknown index type '{index_mode}'")
        if family == "IVF" and metric not in _DOMAIN_FREE and metric not in (M.TANIMOTO, M.DICE):
            raise ValueError(f"unsupported index/metric combination '{index_mode}'")
        if family not in ("FLAT", "IVF") and metric in (M.CANBERRA, M.BRAY_CURTIS):
            raise ValueError(f"unsupported index/metric combination '{index_mode}'")   # exact-only metrics have no graph index
        if any(tok in mode.split("-") for tok in ("SQ8", "PQ", "RABITQ", "POLARVEC")):
            if metric not in _DOMAIN_FREE:
                raise ValueError(f"unsupported index/metric combination '{index_mode}'")
            raise ValueError(f"quantized index '{index_mode}' is outside this package's scope")
        with self._lock:
            if self._dim is not None and not M.accepts_dimension(metric, self._dim):
                raise ValueError(f"metric '{M.NAMES[metric]}' requires dimension 2 as [longitude_degrees, latitude_degrees]")
            if self._ivf is not None:
                self._ivf.close()
                self._ivf = None
            self._metric = metric
            self._index_mode = mode
            self._ivf_params = {}
            if family == "IVF":
                n_clusters = kwargs.get("n_clusters") or kwargs.get("n_centroids") or DEFAULT_N_CLUSTERS
                self._ivf_params = {"n_clusters": int(n_clusters), "nprobe": int(kwargs.get("nprobe") or DEFAULT_NPROBE)}
                self._flush_pending()
                self._build_ivf()
            elif self._store is not None and len(self._store):
                self._store.prepare(metric)

    def _build_ivf(self) -> None:
        if self._store is None or len(self._store) == 0:
            return
        if isinstance(self._store, ShardedDeviceIndex) or self._store.dtype != "float32":
            # the inverted lists address the f32 rows of ONE device index: gather (and decode) the store on the first device
            single = DeviceIndex(self._dim, "float32", self._devices[0])
            if self._dtypes == "float16":
                single.set_segment_target(1 << 62)
            n, pos = len(self._store), 0
            for rows in self._store.segments():
                single.new_segment()
                single.append(self._store.read_rows(pos, rows))
                pos += rows
            assert pos == n
            self._store.close()
            self._store = single
            self._devices = self._devices[:1]
        self._ivf = IVFIndex(self._store, self._metric, n_clusters=self._ivf_params["n_clusters"],
                             nprobe=self._ivf_params["nprobe"])

    def remove_index(self, field_name: str = "default") -> None:
        with self._lock:
            if self._ivf is not None:
                self._ivf.close()
            self._ivf = None
            self._index_mode = None
            self._metric = M.IP

    # ------------------------------------------------------------------ search
    def _subset_rows(self, where, filter_ids) -> Optional[np.ndarray]:
        if where is None and filter_ids is None:
            return None
        rows = None
        if filter_ids is not None:
            rows = {self._ids.row_of(i) for i in filter_ids if i in self._ids}
        if where is not None:
            pred: Callable[[dict], bool]
            if isinstance(where, str):
                where = _compile_where(where)
            if isinstance(where, dict):
                cond = dict(where)
                pred = lambda f: all(f.get(k) == v for k, v in cond.items())  # noqa: E731
            else:
                pred = where
            matched = {r for r in range(len(self._ids)) if pred(self._fields.get(r, {}))}
            rows = matched if rows is None else rows & matched
        return np.fromiter(sorted(rows), dtype=np.uint64, count=len(rows))

    # ---- tombstones as a row mask ------------------------------------------------------------------------------------
    def _dead_row_array(self) -> np.ndarray:
        if self._dead_rows is None:
            self._dead_rows = np.sort(np.fromiter((self._ids.row_of(t) for t in self._tombstones), dtype=np.uint64, count=len(self._tombstones)))
        return self._dead_rows

    def _live_mask(self, n_rows: int, allow: Optional[np.ndarray], first_row: int = 0) -> np.ndarray:
        """Allow-bitset over rows [first_row, first_row + n_rows) with the tombstoned rows cleared (and ``allow`` applied).
        Masking the deleted rows out and asking for k gives the same live top-k as the reference's over-fetch of
        ``k + |tombstones|`` followed by the tombstone filter (src/engine.rs:4735-4741, :3286-3308)."""
        key = (first_row, n_rows)
        base = self._dead_masks.get(key)
        if base is None:
            dead = self._dead_row_array()
            dead = dead[(dead >= first_row) & (dead < first_row + n_rows)] - np.uint64(first_row)
            base = np.full((n_rows + 63) // 64, np.uint64(0xFFFFFFFFFFFFFFFF), dtype=np.uint64)
            if n_rows & 63:
                base[-1] = (np.uint64(1) << np.uint64(n_rows & 63)) - np.uint64(1)
            np.bitwise_and.at(base, (dead >> np.uint64(6)).astype(np.int64), ~(np.uint64(1) << (dead & np.uint64(63))))
            self._dead_masks[key] = base
        return base if allow is None else (base & allow)

    def _search_rows(self, q: np.ndarray, k: int, nprobe: int, subset: Optional[np.ndarray], single: bool = False):
        """(rows[nq, m] u64, dists[nq, m] f32, counts[nq]) over flushed + pending rows, before id mapping: scan,
        pending_search, merge_row_results.  Tombstoned rows may be present when the over-fetch path was taken."""
        nq = q.shape[0]
        n_store = len(self._store) if self._store is not None else 0
        n_dead = len(self._tombstones)
        search_k = k + n_dead                                # engine.rs:4735-4741
        rows = np.empty((nq, 0), np.uint64)
        dists = np.empty((nq, 0), np.float32)
        counts = np.zeros(nq, np.int64)
        if n_store and k > 0:
            allow = None
            if subset is not None:
                allow = make_allow_bits(n_store, subset[subset < n_store])
            if self._index_mode and self._index_mode.startswith("IVF"):
                if self._ivf is None:
                    self._build_ivf()
                stored_k = search_k
                if search_k > NATIVE_MAX_K:
                    allow = self._live_mask(n_store, allow)
                    stored_k = k
                r, d, c = self._ivf.search(q, stored_k, nprobe, allow)
            else:
                f16_rows = self._dtypes == "float16" and (single or subset is not None)
                stored_k = search_k
                if n_dead and (search_k > TENSOR_PLAN_MAX_K >= k or search_k > NATIVE_MAX_K):
                    # many deleted rows: asking for k + |tombstones| would push the search off the tensor-core plan (or
                    # past the native k limit): mask the deleted rows out and ask for k
                    allow = self._live_mask(n_store, allow)
                    stored_k = k
                r, d, c = self._store.search(q, stored_k, self._metric, allow, f16_rows=f16_rows)
            rows, dists, counts = r.astype(np.uint64), d, c.astype(np.int64)
        if self._pending_rows and k > 0:
            # pending_search (src/engine.rs:3310-3360): compute_distance_f32 against every un-flushed row, top-k, then
            # merge_row_results with the flushed hits
            n_p = self._pending_rows
            allow, any_allowed = None, True
            if subset is not None:
                local = subset[(subset >= n_store) & (subset < n_store + n_p)] - np.uint64(n_store)
                any_allowed = local.size > 0
                allow = make_allow_bits(n_p, local)
            if any_allowed:
                asc = M.is_ascending(self._metric)
                pk = min(search_k, n_p)
                if n_dead and pk > NATIVE_MAX_K:
                    allow = self._live_mask(n_p, allow, first_row=n_store)
                    pk = min(k, n_p)
                prow, pd, pc = self._pending_store().search(q, pk, self._metric, allow, pairwise=True)
                limit = max(search_k, rows.shape[1])
                merged = [_merge_row_results(rows[i, :int(counts[i])], dists[i, :int(counts[i])],
                                             prow[i, :int(pc[i])].astype(np.uint64) + np.uint64(n_store), pd[i, :int(pc[i])].copy(), limit, asc)
                          for i in range(nq)]
                width = max((len(m[0]) for m in merged), default=0)
                rows = np.zeros((nq, width), np.uint64)
                dists = np.zeros((nq, width), np.float32)
                counts = np.zeros(nq, np.int64)
                for i, (mr, md) in enumerate(merged):
                    rows[i, :len(mr)], dists[i, :len(mr)], counts[i] = mr, md, len(mr)
        return rows, dists, counts

    def _row_id_array(self) -> Optional[np.ndarray]:
        """row -> external id as an int64 array while every id is an integer (else ``None``: ids are mapped one by one)."""
        return self._ids.array()

    def _finish_batch(self, rows: np.ndarray, dists: np.ndarray, counts: np.ndarray, k: int, return_fields: bool) -> List[ResultView]:
        """row_to_user_id + filter_tombstoned_limit (src/engine.rs:3071-3073, :3286-3308) for a whole batch at once."""
        nq, width = rows.shape
        idx_type, dist_name = M.parse_index_mode(self._index_mode or "FLAT-IP")
        ids_np = self._row_id_array()
        fields_of = self._fields

        def lazy(row_block: np.ndarray, d_block: np.ndarray, m: int) -> List[ResultView]:
            # one id gather for the batch (none at all while ids are the row numbers); views are built when first touched
            id_block = row_block.astype(np.int64) if self._ids.identity else ids_np[row_block.astype(np.int64)]

            def make(i: int) -> ResultView:
                flds = [dict(fields_of.get(int(r), {})) for r in row_block[i]] if return_fields else []
                return ResultView(ids=id_block[i], distances=d_block[i], fields=flds, k=m, distance=dist_name, index=idx_type,
                                  result_type="search")

            return _LazyViews(nq, make)

        m = min(k, width)
        if ids_np is not None and nq and not self._tombstones and int(counts.min()) >= m:
            # the common case: nothing deleted, every query has its k hits -> no masks, no prefix sums
            return lazy(rows[:, :m], np.ascontiguousarray(dists[:, :m], dtype=np.float32), m)
        valid = np.arange(width)[None, :] < counts[:, None]
        if self._tombstones and width:
            dead = self._dead_row_array()
            pos = np.searchsorted(dead, rows)
            pos[pos >= len(dead)] = max(len(dead) - 1, 0)
            valid &= ~(dead[pos] == rows) if len(dead) else True
        keep = valid & (np.cumsum(valid, axis=1) <= k)
        n_keep = keep.sum(axis=1)
        out: List[ResultView] = []
        if ids_np is not None and bool((n_keep == np.minimum(counts, k)).all()) and bool(keep[:, :int(n_keep.max(initial=0))].all() if nq else True) \
                and (nq == 0 or int(n_keep.min()) == int(n_keep.max())):
            # nothing filtered out of the leading columns, every query has the same number of hits
            m = int(n_keep[0]) if nq else 0
            return lazy(rows[:, :m], np.ascontiguousarray(dists[:, :m], dtype=np.float32), m)
        for i in range(nq):
            sel = np.nonzero(keep[i])[0]
            r = rows[i, sel].astype(np.int64)
            if ids_np is not None:
                id_arr = ids_np[r]
            else:
                ids = [self._ids.id_of(int(x)) for x in r]
                all_int = all(isinstance(x, (int, np.integer)) and not isinstance(x, bool) for x in ids)
                id_arr = np.asarray(ids, dtype=np.int64) if all_int else np.asarray(ids, dtype=object)
            flds = [dict(self._fields.get(int(x), {})) for x in r] if return_fields else []
            out.append(ResultView(ids=id_arr, distances=np.asarray(dists[i, sel], dtype=np.float32), fields=flds, k=len(sel),
                                  distance=dist_name, index=idx_type, result_type="search"))
        return out

    def search(self, vector=None, k: int = 10, *, document=None, embed_func=None, where=None, return_fields: bool = False,
               vector_field: str = "default", reranker=None, rerank_k=None, rerank_with_fields: bool = False, nprobe: int = 10,
               approx: bool = False, eps: float = 1e-4, wire_dtype: str = "float32", filter_ids: Optional[Iterable] = None
               ) -> ResultView:
        del wire_dtype, rerank_with_fields
        if (vector is None) == (document is None):
            raise ValueError("search() requires exactly one of vector or document")
        if document is not None or embed_func is not None or reranker is not None or rerank_k is not None:
            raise NotImplementedError("document search and external rerankers are outside this package's scope")
        if vector_field != "default":
            raise NotImplementedError("named vector fields are outside this package's scope")
        result = self._batch_search(np.ascontiguousarray(vector, dtype=np.float32).reshape(1, -1), k, where, return_fields, nprobe,
                                    filter_ids, single=True)[0]
        # approx=True: the ids come from the exact GPU scan (it supersedes the CPU shortlist heuristics); what a caller
        # can observe of the reference's approximate mode is kept — on an unfiltered FLAT search with a metric that
        # supports it, distances are rounded to multiples of eps (src/engine.rs:4756-4763, :4817-4819)
        if (approx and where is None and filter_ids is None and _index_family(self._index_mode or "FLAT-IP") == "FLAT"
                and self._metric in _APPROX_METRICS and result.distances is not None and len(result.distances)):
            _round_to_eps(result.distances, eps)
        return result

    def search_profile(self, vector, k: int = 10, *, where=None, nprobe: int = 10, approx: bool = False, eps: float = 1e-4) -> dict:
        """``search_profile`` (python/lynse/api/local_client.py:1049-1059 -> ``Collection::search_with_profile``,
        src/engine.rs:5005-5054): ``{"items": {k, ids, scores, index}, "profile": {...}}`` with the reference's profile keys
        (query_kind, vector_field, index_path, total_vectors, filter_expression, filter_matches, scanned_vectors, result_count,
        filter_us, search_us, rerank_us, total_us) plus ``device`` — what the GPU did (plan, kernels, fallbacks, timings)."""
        import time

        started = time.perf_counter()
        q = np.ascontiguousarray(vector, dtype=np.float32).reshape(1, -1)
        filter_us, filter_matches = 0, None
        filter_ids = None
        with self._lock:
            if where is not None:
                t0 = time.perf_counter()
                rows = self._subset_rows(where, None)
                filter_ids = [self._ids.id_of(int(r)) for r in rows]
                filter_us = int((time.perf_counter() - t0) * 1e6)
                filter_matches = len(filter_ids)
            t0 = time.perf_counter()
            result = self.search(q[0], k, filter_ids=filter_ids, nprobe=nprobe, approx=approx, eps=eps) if where is not None \
                else self.search(q[0], k, nprobe=nprobe, approx=approx, eps=eps)
            search_us = int((time.perf_counter() - t0) * 1e6)
            total = len(self._ids)
            family = _index_family(self._index_mode or "FLAT-IP")
            index_path = "ann_index" if family != "FLAT" else ("flat_mmap_filtered" if where is not None else "flat_mmap")
            device = self._store.last_stats() if self._store is not None and len(self._store) else {}
        ids = result.ids.tolist() if result.ids is not None else []
        return {
            "items": {"k": int(k), "ids": ids, "scores": [float(x) for x in (result.distances if result.distances is not None else [])],
                      "index": self._index_mode or "FLAT-IP"},
            "profile": {"query_kind": "vector", "vector_field": "default", "index_path": index_path, "total_vectors": total,
                        "filter_expression": where if isinstance(where, str) else (None if where is None else repr(where)),
                        "filter_matches": filter_matches, "scanned_vectors": filter_matches if filter_matches is not None else total,
                        "result_count": len(ids), "filter_us": filter_us, "search_us": search_us, "rerank_us": 0,
                        "total_us": int((time.perf_counter() - started) * 1e6), "device": device},
        }

    def batch_search(self, vectors, k: int = 10, *, where=None, return_fields: bool = False, nprobe: int = 10, reranker=None,
                     rerank_k=None, rerank_with_fields: bool = False, wire_dtype: str = "float32",
                     filter_ids: Optional[Iterable] = None) -> List[ResultView]:
        del wire_dtype, rerank_with_fields
        if reranker is not None or rerank_k is not None:
            raise NotImplementedError("external rerankers are outside this package's scope")
        return self._batch_search(vectors, k, where, return_fields, nprobe, filter_ids, single=False)

    def _batch_search(self, vectors, k, where, return_fields, nprobe, filter_ids, single: bool) -> List[ResultView]:
        q = np.ascontiguousarray(vectors, dtype=np.float32)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        k = int(k)
        nq = q.shape[0]
        with self._lock:
            if self._dim is None or not len(self._ids) or k <= 0:      # empty collection / k = 0: empty result, not an error
                empty = (np.empty((nq, 0), np.uint64), np.empty((nq, 0), np.float32), np.zeros(nq, np.int64))
                return self._finish_batch(*empty, max(k, 0), return_fields)
            if q.shape[1] != self._dim:
                raise ValueError(f"Dimension mismatch: expected {self._dim}, got {q.shape[1]}")
            subset = self._subset_rows(where, filter_ids)
            rows, dists, counts = self._search_rows(q, k, int(nprobe), subset, single)
            return self._finish_batch(rows, dists, counts, k, return_fields)

    def __repr__(self) -> str:
        return f"Collection(name={self.name!r}, shape={self.shape}, index_mode={self._index_mode!r})"


_APPROX_METRICS = (M.IP, M.L2, M.COSINE, M.MANHATTAN, M.CHEBYSHEV, M.CANBERRA, M.BRAY_CURTIS)   # supports_flat_approx


def _round_to_eps(distances: np.ndarray, eps) -> None:
    """``round_distances_to_eps`` in place (src/storage/approx_search.rs:113-143): eps is normalised to a finite value
    >= 1e-8 (default 1e-4), each finite distance becomes ``round(d / eps) * eps`` in f32, halves away from zero."""
    try:
        e = np.float32(eps)
    except (TypeError, ValueError, OverflowError):
        e = np.float32("nan")
    e = np.float32(max(e, np.float32(1e-8))) if np.isfinite(e) and e > 0 else np.float32(1e-4)
    with np.errstate(all="ignore"):
        scaled = distances / e
        t = np.trunc(scaled)
        r = (t + np.sign(scaled) * (np.abs(scaled - t) >= np.float32(0.5))).astype(np.float32)
        rounded = (r * e).astype(np.float32)
        ok = np.isfinite(distances) & np.isfinite(scaled) & np.isfinite(rounded)
    distances[ok] = rounded[ok]


def _merge_row_results(l_rows, l_d, r_rows, r_d, limit: int, ascending: bool):
    """``Collection::merge_row_results`` (src/engine.rs:3362-3419): best score per row, sort by (score, row), truncate."""
    if len(r_rows) == 0:
        return l_rows, l_d
    if len(l_rows) == 0:
        return np.asarray(r_rows, np.uint64), np.asarray(r_d, np.float32)
    best: Dict[int, float] = {}
    for r, d in zip(l_rows.tolist(), l_d.tolist()):
        best[r] = d
    for r, d in zip(np.asarray(r_rows).tolist(), np.asarray(r_d).tolist()):
        if r not in best or (d < best[r] if ascending else d > best[r]):
            best[r] = d
    pairs = sorted(best.items(), key=(lambda p: (p[1], p[0])) if ascending else (lambda p: (-p[1], p[0])))[:limit]
    return np.asarray([p[0] for p in pairs], np.uint64), np.asarray([p[1] for p in pairs], np.float32)


class _LazyViews(list):
    """The ``list[ResultView]`` of a batch search whose views are built from the batch's id / distance blocks when they
    are first touched: a 1024-query batch costs two array gathers instead of 1024 object constructions unless the caller
    reads every view.  A real ``list`` (``isinstance`` holds); every read path materialises what it returns."""

    __slots__ = ("_make",)

    def __init__(self, n: int, make: Callable[[int], ResultView]):
        super().__init__([None] * n)
        self._make = make

    def _fill(self, i: int) -> ResultView:
        v = list.__getitem__(self, i)
        if v is None:
            v = self._make(i if i >= 0 else len(self) + i)
            list.__setitem__(self, i, v)
        return v

    def _fill_all(self) -> None:
        for i in range(len(self)):
            self._fill(i)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._fill(j) for j in range(*i.indices(len(self)))]
        return self._fill(i)

    def __iter__(self):
        for i in range(len(self)):
            yield self._fill(i)

    def __reversed__(self):
        for i in range(len(self) - 1, -1, -1):
            yield self._fill(i)

    def __contains__(self, item) -> bool:
        return any(v == item for v in self)

    def __eq__(self, other) -> bool:
        self._fill_all()
        if isinstance(other, _LazyViews):
            other._fill_all()
        return list.__eq__(self, other)

    __hash__ = None

    def __add__(self, other):
        return list(self) + list(other)

    def __mul__(self, n):
        return list(self) * n

    def copy(self):
        return list(self)

    def index(self, *args):
        self._fill_all()
        return list.index(self, *args)

    def count(self, item) -> int:
        self._fill_all()
        return list.count(self, item)

    def __repr__(self) -> str:
        self._fill_all()
        return list.__repr__(self)

    def __reduce__(self):
        return (list, (list(self),))


class InsertSession:
    """The reference's ``DataInsertionSession``: buffers ``add`` calls, applies them and commits on exit."""

    def __init__(self, coll: Collection):
        self.db = coll
        self._adds: List[tuple] = []

    def __enter__(self):
        return self

    def __getattr__(self, name):
        return getattr(self.db, name)

    def add(self, ids=None, *, vectors=None, documents=None, embed_func=None, fields=None, batch_size: int = 1000, wire_dtype: str = "float32"):
        if documents is not None or embed_func is not None:
            raise NotImplementedError("document embedding is outside this package's scope; pass vectors")
        self._adds.append((ids, vectors, fields, batch_size))

    def commit(self) -> None:
        for ids, vectors, fields, batch_size in self._adds:
            self.db.add(ids, vectors=vectors, fields=fields, batch_size=batch_size)
        self._adds.clear()
        self.db.commit()

    def __exit__(self, exc_type, exc, tb):
        if exc_type is None:
            self.commit()
        return False


_WHERE_TERM = None


def _compile_where(expr: str) -> Callable[[dict], bool]:
    """The conjunctive subset of the reference's SQL ``where`` strings that needs no metadata engine:
    ``"field" <op> literal [AND ...]`` with ``= == != <> < <= > >=``, numeric or quoted-string literals.  Anything else
    (OR, functions, LIKE, ...) belongs to the reference's SQL engine and raises ``NotImplementedError``."""
    import re

    global _WHERE_TERM
    if _WHERE_TERM is None:
        _WHERE_TERM = re.compile(r"""^\s*(?:"([^"]+)"|([A-Za-z_][A-Za-z_0-9]*))\s*(==|=|!=|<>|<=|>=|<|>)\s*(?:'([^']*)'|(-?\d+(?:\.\d+)?(?:[eE][-+]?\d+)?)|(true|false))\s*$""", re.I)
    import operator

    ops = {"=": operator.eq, "==": operator.eq, "!=": operator.ne, "<>": operator.ne, "<": operator.lt, "<=": operator.le,
           ">": operator.gt, ">=": operator.ge}
    terms = []
    for part in re.split(r"\s+AND\s+", expr.strip(), flags=re.I):
        m = _WHERE_TERM.match(part)
        if not m:
            raise NotImplementedError(f"where expression {expr!r} needs the reference's SQL metadata engine; pass a callable or a dict")
        name = m.group(1) or m.group(2)
        if m.group(4) is not None:
            lit: Any = m.group(4)
        elif m.group(5) is not None:
            lit = float(m.group(5)) if any(c in m.group(5) for c in ".eE") else int(m.group(5))
        else:
            lit = m.group(6).lower() == "true"
        terms.append((name, ops[m.group(3)], lit))

    def pred(fields: dict) -> bool:
        for name, op, lit in terms:
            if name not in fields or fields[name] is None:
                return False
            try:
                if not op(fields[name], lit):
                    return False
            except TypeError:
                return False
        return True

    return pred


class Database:
    """``LocalClient``: the collections of one database."""

    def __init__(self, manager: "VectorDBClient", database_name: str):
        self._manager = manager
        self.database_name = database_name

    @property
    def _colls(self) -> Dict[str, Collection]:
        return self._manager._dbs[self.database_name]

    def require_collection(self, collection: str, dim: Optional[int] = None, n_threads: Optional[int] = 10, warm_up: bool = False,
                           drop_if_exists: bool = False, description: Optional[str] = None, dtypes: str = "float32",
                           default_index: Optional[str] = "FLAT-IP") -> Collection:
        del n_threads, warm_up
        if drop_if_exists and collection in self._colls:
            self.drop_collection(collection)
        if collection not in self._colls:
            self._colls[collection] = Collection(collection, dim, dtypes=dtypes, default_index=default_index,
                                                 description=description, device=self._manager._device, devices=self._manager._devices)
        coll = self._colls[collection]
        if dim is not None and coll._dim is not None and int(dim) != coll._dim:
            raise ValueError(f"collection {collection!r} has dimension {coll._dim}, not {dim}")
        return coll

    def get_collection(self, collection: str, warm_up: bool = True) -> Collection:
        del warm_up
        if collection not in self._colls:
            raise ValueError(f"Collection '{collection}' does not exist.")
        return self._colls[collection]

    def drop_collection(self, collection: str) -> None:
        coll = self._colls.pop(collection, None)
        if coll is not None:
            coll.close()

    def show_collections(self) -> List[str]:
        return sorted(self._colls)

    def show_collections_details(self) -> List[dict]:
        return [self._colls[c].stats() for c in self.show_collections()]

    def database_exists(self) -> bool:
        return self.database_name in self._manager._dbs

    def drop_database(self) -> None:
        self._manager.drop_database(self.database_name)

    def __repr__(self) -> str:
        return f"Database(name={self.database_name!r}, collections={self.show_collections()})"


class VectorDBClient:
    """``VectorDBClient(uri=None)``: local, in-memory databases whose vectors live in the HBM of this process's B200s.
    ``devices`` (or ``LYNSE_B200_DEVICES=0,1,...``) names the GPUs a collection's segments are spread over; the default is
    ``device`` alone."""

    def __init__(self, uri: Union[str, None] = None, api_key: Optional[str] = None, read_only: bool = False, device: int = 0,
                 devices: Optional[Sequence[int]] = None):
        del api_key
        if uri is not None and str(uri).startswith(("http://", "https://")):
            raise NotImplementedError("the HTTP client is outside this package's scope")
        if read_only:
            raise NotImplementedError("read_only opens persisted storage, which is outside this package's scope")
        self._uri = None if uri is None else str(uri)
        self._device = int(device)
        import os

        if devices is None and os.environ.get("LYNSE_B200_DEVICES", "").strip():
            devices = visible_devices()
        self._devices = [int(d) for d in devices] if devices else [self._device]
        self._dbs: Dict[str, Dict[str, Collection]] = {}

    def create_database(self, database_name: str, drop_if_exists: bool = False) -> Database:
        if len(self._dbs) >= MAX_DATABASES and database_name not in self._dbs:
            raise ValueError("The maximum number of databases created is 64.")
        if drop_if_exists and database_name in self._dbs:
            self.drop_database(database_name)
        self._dbs.setdefault(database_name, {})
        return Database(self, database_name)

    def create_collection(self, database_name: str, collection: str, dim: Optional[int] = None, n_threads: Optional[int] = 10,
                          warm_up: bool = False, drop_if_exists: bool = False, description: Optional[str] = None,
                          dtypes: str = "float32", default_index: Optional[str] = "FLAT-IP",
                          drop_database_if_exists: bool = False) -> Collection:
        if drop_database_if_exists or database_name not in self._dbs:
            db = self.create_database(database_name, drop_if_exists=drop_database_if_exists)
        else:
            db = self.get_database(database_name)
        return db.require_collection(collection=collection, dim=dim, n_threads=n_threads, warm_up=warm_up,
                                     drop_if_exists=drop_if_exists, description=description, dtypes=dtypes,
                                     default_index=default_index)

    def get_database(self, database_name: str) -> Database:
        if database_name not in self._dbs:
            raise ValueError(f"{database_name} does not exist.")
        return Database(self, database_name)

    def list_databases(self) -> List[str]:
        return sorted(self._dbs)

    def drop_database(self, database_name: str) -> None:
        for coll in self._dbs.pop(database_name, {}).values():
            coll.close()

    def close(self) -> None:
        for name in list(self._dbs):
            self.drop_database(name)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __repr__(self) -> str:
        return f"VectorDBClient(uri={self._uri!r}, databases={self.list_databases()})"
