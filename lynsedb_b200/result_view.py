"""ResultView — the container ``search`` / ``batch_search`` return.

Keeps the contract of the reference's python/lynse/result_view.py:75-320 for
search results: keyword-only constructor, ``ids`` / ``distances`` / ``fields``
accessors, string-key indexing, tuple unpacking as ``(ids, distances, fields)``,
equality, truthiness and the ``to_*`` conversions.  Written against that
contract, not copied from it; only what the search path produces is kept
(``result_type`` "search", plus "data"/"query" shapes for completeness).
"""
from __future__ import annotations

import json
from typing import Any, Dict, List, Optional

import numpy as np

from .metrics import parse_index_mode as _parse_index_mode  # noqa: F401  (re-exported, same name as the reference)


class ResultView:
    __slots__ = ("_ids", "_distances", "_vectors", "_fields", "_k", "_distance", "_index", "_result_type")

    def __init__(self, *, ids=None, distances=None, vectors=None, fields: Optional[List[Dict[str, Any]]] = None,
                 k: Optional[int] = None, distance: Optional[str] = None, index: Optional[str] = None,
                 result_type: str = "search"):
        self._ids = ids
        self._distances = distances
        self._vectors = vectors
        self._fields = fields if fields is not None else []
        self._k = k
        self._distance = distance
        self._index = index
        self._result_type = result_type

    # components in the order tuple-unpacking yields them
    def _components(self) -> list:
        if self._result_type == "search":
            return [self._ids, self._distances, self._fields]
        if self._result_type == "data":
            return [self._vectors, self._ids, self._fields]
        return [self._ids, self._fields] if self._fields else [self._ids]

    ids = property(lambda self: self._ids)
    distances = property(lambda self: self._distances)
    vectors = property(lambda self: self._vectors)
    fields = property(lambda self: self._fields)
    k = property(lambda self: self._k)
    distance_metric = property(lambda self: self._distance)
    index_type = property(lambda self: self._index)
    result_type = property(lambda self: self._result_type)

    def __len__(self) -> int:
        for part in (self._ids, self._distances):
            if part is not None:
                return len(part)
        if self._vectors is not None:
            return self._vectors.shape[0]
        return len(self._fields) if self._fields else 0

    def __getitem__(self, key):
        if not isinstance(key, str):
            raise TypeError(f"ResultView indices must be strings, not {type(key).__name__}")
        table = {"ids": self._ids, "distance": self._distances, "distances": self._distances,
                 "vectors": self._vectors, "fields": self._fields, "k": self._k, "measure": self._distance,
                 "index": self._index}
        if key in table:
            return table[key]
        if key == "n":
            return len(self)
        raise KeyError(f"ResultView has no key {key!r}")

    def __iter__(self):
        yield from self._components()

    def __eq__(self, other) -> bool:
        if not isinstance(other, ResultView):
            return NotImplemented
        if self._result_type != other._result_type or len(self) != len(other):
            return False
        for a, b in ((self._ids, other._ids), (self._distances, other._distances), (self._vectors, other._vectors)):
            if (a is None) != (b is None):
                return False
            if a is not None and not np.array_equal(a, b):
                return False
        return self._fields == other._fields

    __hash__ = None

    def __bool__(self) -> bool:
        return len(self) > 0

    def __repr__(self) -> str:
        bits = []
        if self._ids is not None:
            bits.append(f"ids={np.asarray(self._ids)[:8].tolist()}{'...' if len(self) > 8 else ''}")
        if self._distances is not None:
            bits.append(f"distances={np.asarray(self._distances)[:8].tolist()}{'...' if len(self) > 8 else ''}")
        if self._vectors is not None:
            bits.append(f"vectors=array{tuple(self._vectors.shape)}")
        if self._k is not None:
            bits.append(f"k={self._k}")
        if self._distance:
            bits.append(f"measure={self._distance!r}")
        if self._index:
            bits.append(f"index={self._index!r}")
        if self._fields:
            bits.append(f"fields=[{len(self._fields)} rows]")
        return "ResultView(" + ", ".join(bits) + ")"

    # conversions -----------------------------------------------------------------
    def to_tuple(self) -> tuple:
        return tuple(self._components())

    def to_numpy(self) -> Dict[str, np.ndarray]:
        out = {}
        for name, part in (("ids", self._ids), ("distances", self._distances), ("vectors", self._vectors)):
            if part is not None:
                out[name] = part
        return out

    def to_dict(self) -> Dict[str, Any]:
        d: Dict[str, Any] = {}
        order = {"search": ("ids", "distances"), "data": ("vectors", "ids")}.get(self._result_type, ("ids",))
        parts = {"ids": self._ids, "distances": self._distances, "vectors": self._vectors}
        for name in order:
            part = parts[name]
            if part is not None:
                d[name] = part.tolist() if isinstance(part, np.ndarray) else list(part)
        if self._fields:
            keys = sorted({key for row in self._fields if row for key in row})
            for key in keys:
                d[key] = [row.get(key) if row else None for row in self._fields]
        return d

    def to_list(self) -> List[Dict[str, Any]]:
        rows = []
        for i in range(len(self)):
            row: Dict[str, Any] = {}
            if self._ids is not None:
                v = self._ids[i]
                row["id"] = v.item() if hasattr(v, "item") else v
            if self._distances is not None:
                row["distance"] = float(self._distances[i])
            if self._vectors is not None:
                row["vector"] = self._vectors[i].tolist()
            if self._fields and i < len(self._fields) and self._fields[i]:
                row.update(self._fields[i])
            rows.append(row)
        return rows

    def to_json(self, orient: str = "records", indent: Optional[int] = None) -> str:
        payload = self.to_list() if orient == "records" else self.to_dict()
        return json.dumps(payload, indent=indent, default=str)

    def to_pandas(self):
        import pandas as pd

        return pd.DataFrame(self.to_dict())

    def to_arrow(self):
        """``pyarrow.Table``: ``id`` / ``distance`` (search), ``id`` / ``vector`` (data), then one column per field
        name (python/lynse/result_view.py:470-520).  Ids that are not all integers become strings."""
        try:
            import pyarrow as pa
        except ImportError as e:
            raise ImportError("pyarrow is required for to_arrow(). Install it with: pip install pyarrow") from e

        def id_array(ids):
            vals = [v.item() if hasattr(v, "item") else v for v in ids]
            if all(isinstance(v, int) and not isinstance(v, bool) for v in vals):
                return pa.array(vals, type=pa.int64())
            return pa.array([str(v) for v in vals], type=pa.string())

        arrays = {}
        if self._ids is not None:
            arrays["id"] = id_array(self._ids)
        if self._result_type == "search" and self._distances is not None:
            arrays["distance"] = pa.array(np.asarray(self._distances, dtype=np.float32), type=pa.float32())
        if self._result_type == "data" and self._vectors is not None:
            arrays["vector"] = pa.array(np.asarray(self._vectors, dtype=np.float32).tolist(), type=pa.list_(pa.float32()))
        if self._fields:
            for key in sorted({key for row in self._fields if row for key in row}):
                arrays[key] = pa.array([row.get(key) if row else None for row in self._fields])
        return pa.table(arrays)

    def to_polars(self):
        """``polars.DataFrame`` (python/lynse/result_view.py: requires polars)."""
        try:
            import polars as pl
        except ImportError as e:
            raise ImportError("polars is required for to_polars(). Install it with: pip install polars") from e
        return pl.from_arrow(self.to_arrow())
