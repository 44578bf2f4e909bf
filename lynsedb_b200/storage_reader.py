"""Reader for LynseDB's on-disk vector store, so an existing collection directory can be loaded straight into HBM.

Format (reference src/storage/vector_store.rs:24-60, :157-243; src/storage/dtype.rs):
  <collection>/vector_manifest.json   {"version": 1, "generation": g, "id_map_file": "id_map.bin",
                                       "segments": [{"file": "vector_segments/...", "rows": n}, ...]}
  segment files                        raw little-endian row-major rows, f32 (4 B) or f16 (2 B) per value; the row
                                       count is re-derived from the file length, exactly as the reference does
  <id_map_file>                        one little-endian u64 external id per row, in row order
  legacy layout                        no manifest: a single ``vectors.bin``
Only reading is implemented; writing, the update journal and compaction belong to the storage engine (out of scope).
"""
from __future__ import annotations

import json
import os
from pathlib import Path, PurePosixPath
from typing import List, Optional, Tuple

import numpy as np

MANIFEST_FILE = "vector_manifest.json"
MANIFEST_VERSION = 1
DEFAULT_ID_MAP_FILE = "id_map.bin"


def _validate_relative(path: str, what: str) -> None:
    """``validate_manifest_path``: manifest paths must stay inside the collection directory."""
    p = PurePosixPath(path.replace("\\", "/"))
    if not path or p.is_absolute() or any(part in ("..", "") for part in p.parts) or ":" in p.parts[0]:
        raise IOError(f"vector manifest {what} {path!r} escapes the collection directory")


def read_manifest(collection_path, dim: int, dtype: str = "float32") -> Tuple[List[Tuple[Path, int]], Path]:
    """``[(segment file, rows)]`` in manifest order and the id-map path."""
    root = Path(collection_path)
    width = int(dim) * (4 if dtype == "float32" else 2)
    manifest_path = root / MANIFEST_FILE
    if manifest_path.exists():
        m = json.loads(manifest_path.read_text())
        if int(m.get("version", 0)) > MANIFEST_VERSION:
            raise IOError(f"vector manifest version {m['version']} is newer than supported version {MANIFEST_VERSION}")
        id_map_file = m.get("id_map_file", DEFAULT_ID_MAP_FILE)
        seg_files = [s["file"] for s in m.get("segments", [])]
    else:
        id_map_file = DEFAULT_ID_MAP_FILE
        legacy = root / "vectors.bin"
        seg_files = ["vectors.bin"] if legacy.exists() and width and legacy.stat().st_size >= width else []
    _validate_relative(id_map_file, "ID-map path")
    if len(set(seg_files)) != len(seg_files):
        raise IOError("vector manifest contains a duplicate segment path")
    segments = []
    for f in seg_files:
        _validate_relative(f, "segment path")
        path = root / f
        if not path.exists():
            raise IOError(f"vector manifest segment {path} is unavailable")
        segments.append((path, (path.stat().st_size // width) if width else 0))
    return segments, root / id_map_file


def read_segment(path, rows: int, dim: int, dtype: str = "float32") -> np.ndarray:
    """One segment as an f32 ``[rows, dim]`` array (f16 rows widen exactly, IEEE binary16 -> binary32)."""
    if rows == 0:
        return np.empty((0, dim), dtype=np.float32)
    raw = np.memmap(path, dtype="<f4" if dtype == "float32" else "<f2", mode="r", shape=(rows, dim))
    return np.ascontiguousarray(raw, dtype=np.float32)


def read_id_map(path, n_rows: int) -> Optional[np.ndarray]:
    if not os.path.exists(path):
        return None
    ids = np.fromfile(path, dtype="<u8")
    if ids.size < n_rows:
        raise IOError(f"{path}: {ids.size} ids for {n_rows} rows")
    return ids[:n_rows]
