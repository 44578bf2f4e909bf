"""The on-disk vector-store reader against files written in the reference's format
(src/storage/vector_store.rs:24-60, :157-243): manifest + raw little-endian segments + u64 id map, and the legacy
single-file layout.  CPU only."""
import json

import numpy as np
import pytest

from lynsedb_b200 import storage_reader as R


def _write_store(root, blocks, dtype="<f4", ids=None, manifest=True):
    (root / "vector_segments").mkdir(parents=True, exist_ok=True)
    segs = []
    for i, b in enumerate(blocks):
        name = f"vector_segments/seg_{i:06d}.bin"
        b.astype(dtype).tofile(root / name)
        segs.append({"file": name, "rows": 0})          # the reader re-derives rows from the file length, as the reference does
    if manifest:
        (root / "vector_manifest.json").write_text(json.dumps(
            {"version": 1, "generation": 3, "id_map_file": "id_map.bin", "segments": segs}))
    if ids is not None:
        np.asarray(ids, dtype="<u8").tofile(root / "id_map.bin")


def test_manifest_segments_and_id_map(tmp_path):
    rng = np.random.default_rng(0)
    blocks = [rng.random((5, 6), dtype=np.float32), rng.random((3, 6), dtype=np.float32)]
    _write_store(tmp_path, blocks, ids=[10, 11, 12, 13, 14, 20, 21, 22])
    segments, id_path = R.read_manifest(tmp_path, 6)
    assert [r for _, r in segments] == [5, 3]
    got = np.concatenate([R.read_segment(p, r, 6) for p, r in segments])
    assert np.array_equal(got, np.concatenate(blocks))
    assert R.read_id_map(id_path, 8).tolist() == [10, 11, 12, 13, 14, 20, 21, 22]


def test_f16_segments_widen_exactly(tmp_path):
    rng = np.random.default_rng(1)
    block = rng.random((4, 8), dtype=np.float32).astype(np.float16)
    _write_store(tmp_path, [block], dtype="<f2")
    (p, r), = R.read_manifest(tmp_path, 8, "float16")[0]
    assert r == 4 and np.array_equal(R.read_segment(p, r, 8, "float16"), block.astype(np.float32))


def test_legacy_single_file_and_missing_id_map(tmp_path):
    block = np.arange(12, dtype=np.float32).reshape(3, 4)
    block.astype("<f4").tofile(tmp_path / "vectors.bin")
    segments, id_path = R.read_manifest(tmp_path, 4)
    assert [(p.name, r) for p, r in segments] == [("vectors.bin", 3)]
    assert R.read_id_map(id_path, 3) is None


def test_manifest_paths_must_stay_inside_the_collection(tmp_path):
    (tmp_path / "vector_manifest.json").write_text(json.dumps(
        {"version": 1, "generation": 0, "id_map_file": "id_map.bin", "segments": [{"file": "../outside.bin", "rows": 1}]}))
    with pytest.raises(IOError):
        R.read_manifest(tmp_path, 4)
    (tmp_path / "vector_manifest.json").write_text(json.dumps({"version": 2, "generation": 0, "id_map_file": "id_map.bin", "segments": []}))
    with pytest.raises(IOError, match="newer than supported"):
        R.read_manifest(tmp_path, 4)
