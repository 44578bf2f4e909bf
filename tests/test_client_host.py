"""Host-side logic of the Collection object model (no GPU): id mapping, tombstones, pending rows, merge rules,
build_index validation.  The device index is replaced by a numpy stand-in with the same method surface, so only the
Python layer (the restated engine post-processing, src/engine.rs:3286-3419, :4718-4833) is under test here."""
import numpy as np
import pytest

import lynsedb_b200.client as client_mod
from lynsedb_b200 import metrics as M
from lynsedb_b200.client import Collection, VectorDBClient, _merge_row_results


class FakeIndex:
    """numpy brute force with DeviceIndex's surface: (score best-first, row ascending)."""

    def __init__(self, dim, dtype="float32", device=0):
        self.dim = dim
        self.rows = np.empty((0, dim), np.float32)
        self.appends = []

    def __len__(self):
        return self.rows.shape[0]

    def append(self, block):
        self.appends.append(block.shape[0])
        self.rows = np.concatenate([self.rows, block])

    def segments(self):
        return list(self.appends)

    def set_segment_target(self, n_bytes):
        self.segment_target = n_bytes

    def prepare(self, metric):
        pass

    def close(self):
        pass

    def search(self, q, k, metric, allow_bits=None, f16_rows=False, pairwise=False):
        assert metric in (M.IP, M.L2)
        self.f16_calls = getattr(self, "f16_calls", []) + [bool(f16_rows)]
        nq, n = q.shape[0], len(self)
        rows = np.full((nq, k), 0xFFFFFFFF, np.uint32)
        dists = np.full((nq, k), np.nan, np.float32)
        counts = np.zeros(nq, np.uint32)
        for i in range(nq):
            s = self.rows @ q[i] if metric == M.IP else ((self.rows - q[i]) ** 2).sum(1)
            cand = np.arange(n)
            if allow_bits is not None:
                ok = np.array([(int(allow_bits[r >> 6]) >> (r & 63)) & 1 for r in range(n)], bool)
                cand = cand[ok]
            order = sorted(cand.tolist(), key=lambda r: ((-s[r]) if metric == M.IP else s[r], r))[:k]
            counts[i] = len(order)
            rows[i, :len(order)] = order
            dists[i, :len(order)] = s[order]
        return rows, dists, counts


@pytest.fixture()
def fake(monkeypatch):
    monkeypatch.setattr(client_mod, "DeviceIndex", FakeIndex)


def test_client_object_model(fake):
    c = VectorDBClient()
    coll = c.create_collection("db", "docs", dim=4)
    assert c.list_databases() == ["db"] and c.get_database("db").show_collections() == ["docs"]
    assert coll.index_mode is None and coll.shape == (0, 4)
    with pytest.raises(ValueError):
        c.get_database("nope")
    r = coll.search(np.ones(4, np.float32), k=3)          # empty collection -> empty result, not an error
    assert len(r) == 0 and r.distance_metric == "IP" and r.index_type == "Flat"
    ids = coll.add(vectors=np.eye(4, dtype=np.float32))
    assert ids == [0, 1, 2, 3] and coll.index_mode == "FLAT-IP"      # default index auto-built after the first write
    assert coll.add(vectors=[0, 0, 0, 2.0]) == 4
    assert coll.add("x", vectors=[0, 0, 3.0, 0]) == "x"
    with pytest.raises(ValueError):
        coll.add("x", vectors=[1, 0, 0, 0])
    with pytest.raises(ValueError):
        coll.add(vectors=np.ones((1, 5), np.float32))
    ids_, d, f = coll.search([0, 0, 1.0, 1.0], k=3)        # tuple-unpacks as (ids, distances, fields)
    assert list(ids_) == ["x", 4, 2] and ids_.dtype == object and np.allclose(d, [3, 2, 1]) and f == []
    c.close()


def test_pending_rows_are_searchable_and_flush_thresholds(fake):
    coll = Collection("c", 8)
    rng = np.random.default_rng(1)
    a = rng.random((9000, 8), dtype=np.float32)
    coll.add(vectors=a, batch_size=4000)
    assert coll.stats()["pending_rows"] == 9000 and coll.stats()["segments"] == []       # below 10 000 rows / 32 MiB
    q = rng.random(8, dtype=np.float32)
    want = np.argsort(-(a @ q), kind="stable")[:5]
    assert list(coll.search(q, k=5).ids) == want.tolist()                                 # un-committed rows are found
    coll.add(vectors=rng.random((2000, 8), dtype=np.float32), batch_size=1000)
    assert coll.stats()["segments"] == [10000] and coll.stats()["pending_rows"] == 1000   # one flush at the threshold
    coll.commit()
    assert coll.stats()["segments"] == [10000, 1000] and coll.COMMIT_FLAG    # (the stand-in records appends, not segments)


def test_tombstones_ask_for_k_plus_deleted_and_refill(fake):
    coll = Collection("c", 2, default_index="FLAT-L2")
    pts = np.array([[i, 0] for i in range(10)], np.float32)
    coll.add(vectors=pts)
    coll.commit()
    assert list(coll.search([0, 0], k=3).ids) == [0, 1, 2]
    assert coll.delete([0, 1]) == 2 and not coll.is_id_exists(0)
    r = coll.search([0, 0], k=3)
    assert list(r.ids) == [2, 3, 4] and np.allclose(r.distances, [4, 9, 16])             # refilled past the tombstones
    assert coll.restore(0) == 1 and list(coll.search([0, 0], k=2).ids) == [0, 2]
    assert len(coll.search([0, 0], k=50)) == 9                                            # k > n


def test_filters_by_fields_and_ids(fake):
    coll = Collection("c", 2, default_index="FLAT-L2")
    coll.add(vectors=np.array([[i, 0] for i in range(8)], np.float32), fields=[{"g": i % 2} for i in range(8)])
    assert list(coll.search([0, 0], k=3, where={"g": 1}).ids) == [1, 3, 5]
    assert list(coll.search([0, 0], k=3, where=lambda f: f.get("g") == 0, filter_ids=[2, 4, 6, 7]).ids) == [2, 4, 6]
    r = coll.search([0, 0], k=2, where={"g": 1}, return_fields=True)
    assert r.fields == [{"g": 1}, {"g": 1}]
    assert list(coll.search([0, 0], k=3, where='"g" = 1').ids) == [1, 3, 5]          # the conjunctive subset of the SQL strings
    assert list(coll.search([0, 0], k=3, where="g = 0 AND g <> 1").ids) == [0, 2, 4]
    with pytest.raises(NotImplementedError):
        coll.search([0, 0], k=1, where="g = 1 OR g = 0")


def test_build_index_validation(fake):
    coll = Collection("c", 3, default_index=None)
    coll.add(vectors=np.eye(3, dtype=np.float32))
    with pytest.raises(ValueError, match="unknown index build parameter"):
        coll.build_index("FLAT-IP", bogus=1)
    with pytest.raises(ValueError, match="unknown index type"):
        coll.build_index("FLAT")
    with pytest.raises(ValueError, match="requires dimension 2"):
        coll.build_index("FLAT-HAVERSINE")
    with pytest.raises(ValueError, match="unsupported index/metric combination"):
        coll.build_index("IVF-WASSERSTEIN")
    with pytest.raises(ValueError, match="unsupported index/metric combination"):
        coll.build_index("HNSW-CANBERRA")                  # graph indexes only exist for the domain-free metrics
    with pytest.raises(ValueError, match="outside this package"):
        coll.build_index("FLAT-IP-SQ8")
    coll.build_index("HNSW-IP")                            # accepted: served by the exact scan, reported as HNSW
    assert M.parse_index_mode(coll.index_mode) == ("HNSW", "IP")
    coll.build_index("FLAT-COS", n_clusters=4)              # FLAT ignores the shared kwargs
    assert coll.index_mode == "FLAT-COS"
    assert M.parse_index_mode(coll.index_mode) == ("Flat", "Cosine")


def test_merge_row_results_rule():
    rows, d = _merge_row_results(np.array([5, 1], np.uint64), np.array([0.5, 0.7], np.float32),
                                 np.array([9, 1, 3], np.uint64), np.array([0.5, 0.6, 0.9], np.float32), 3, True)
    assert rows.tolist() == [5, 9, 1] and np.allclose(d, [0.5, 0.5, 0.6])          # best per row, ties by row, truncated
    rows, d = _merge_row_results(np.array([5], np.uint64), np.array([0.5], np.float32),
                                 np.array([2], np.uint64), np.array([0.9], np.float32), 5, False)
    assert rows.tolist() == [2, 5]                                                   # IP: higher first


def test_database_limit(fake):
    c = VectorDBClient()
    for i in range(64):
        c.create_database(f"d{i}")
    with pytest.raises(ValueError, match="maximum number of databases"):
        c.create_database("one_too_many")


def test_float16_collection_rounds_like_the_reference_encoder(fake):
    # src/engine.rs:8043-8077 (f16_collection_batch_search_reuses_decoded_candidates)
    coll = Collection("c", 4, dtypes="float16")
    coll.add([10, 11, 12], vectors=np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0.5, 0.5, 0, 0]], np.float32))
    res = coll.batch_search(np.array([[1, 0, 0, 0], [0, 1, 0, 0]], np.float32), k=2)
    assert list(res[0].ids) == [10, 12] and list(res[1].ids) == [11, 12]
    assert coll.vector_dtype() == "float16"
    coll2 = Collection("d", 2, dtypes="float16", default_index="FLAT-L2")
    v = np.array([[0.1, 0.2]], np.float32)
    coll2.add(vectors=v)
    stored = v.astype(np.float16).astype(np.float32)
    d = coll2.search(v[0], k=1).distances[0]
    assert d == pytest.approx(float(((stored[0] - v[0]) ** 2).sum()), rel=1e-6) and d > 0   # the f16 rounding is visible
    with pytest.raises(ValueError):
        Collection("e", 2, dtypes="int8")


def test_float16_collection_routes_single_and_filtered_searches_to_the_scalar_kernels(fake):
    # FlatMmap::search / search_filtered on F16 storage use the scalar f16 kernels (flat_mmap.rs:905-907, :511-520);
    # only the unfiltered batch path decodes and runs the f32 kernels (engine.rs:5440-5474)
    rng = np.random.default_rng(0)
    coll = Collection("c", 4, dtypes="float16")
    coll.add(list(range(8)), vectors=rng.random((8, 4), dtype=np.float32))
    coll.commit()
    q = rng.random((2, 4), dtype=np.float32)
    coll.search(q[0], 3)
    coll.batch_search(q, 3)
    coll.batch_search(q[:1], 3)
    coll.batch_search(q, 3, filter_ids=[1, 2, 3])
    assert coll._store.f16_calls == [True, False, False, True]
    plain = Collection("d", 4)
    plain.add(list(range(8)), vectors=rng.random((8, 4), dtype=np.float32))
    plain.commit()
    plain.search(q[0], 3)
    plain.batch_search(q, 3, filter_ids=[1, 2, 3])
    assert plain._store.f16_calls == [False, False]


def test_many_tombstones_become_a_row_mask_with_the_same_result(fake):
    # k + |tombstones| beyond the tensor plan's k: the deleted rows are masked out and k results are asked for;
    # the answer is the live top-k either way (src/engine.rs:4735-4741, :3286-3308)
    rng = np.random.default_rng(4)
    vecs = rng.random((1000, 4), dtype=np.float32)
    coll = Collection("c", 4)
    coll.add(list(range(1000)), vectors=vecs)
    coll.commit()
    order = np.argsort(-(vecs @ vecs[0]), kind="stable")
    coll.delete([int(i) for i in order[1:301]])          # the 300 best after the query itself
    calls = []
    real = coll._store.search

    def spy(q, k, metric, allow_bits=None, **kw):
        calls.append((k, allow_bits is not None))
        return real(q, k, metric, allow_bits, **kw)

    coll._store.search = spy
    got = coll.search(vecs[0], 10)
    assert calls == [(10, True)]
    live = [int(i) for i in order if int(i) not in set(int(j) for j in order[1:301])][:10]
    assert got.ids.tolist() == live
    coll.restore([int(i) for i in order[1:295]])         # 6 tombstones left: back to the over-fetch
    calls.clear()
    coll.search(vecs[0], 10)
    assert calls == [(16, False)]


def test_approx_search_rounds_distances_to_eps_and_search_range_returns_a_view(fake):
    # tests/standard_tests/test_search.py:30-43 (approx ids == exact ids, distances on the eps grid, non-finite eps falls
    # back to 1e-4) and :594-616 (search_range returns a ResultView)
    from lynsedb_b200.client import _round_to_eps
    rng = np.random.default_rng(12)
    vecs = rng.random((200, 4), dtype=np.float32)
    coll = Collection("c", 4)
    coll.add(list(range(200)), vectors=vecs)
    coll.commit()
    exact = coll.search(vecs[3], k=5, approx=False)
    approx = coll.search(vecs[3], k=5, approx=True, eps=1e-4)
    assert approx.ids.tolist() == exact.ids.tolist()
    assert np.max(np.abs(approx.distances - exact.distances)) <= 1e-4
    scaled = approx.distances / 1e-4
    assert np.allclose(scaled, np.round(scaled), atol=1e-3)
    inf_eps = coll.search(vecs[3], k=5, approx=True, eps=float("inf"))
    assert len(inf_eps.ids) == 5 and np.all(np.isfinite(inf_eps.distances))
    assert np.array_equal(inf_eps.distances, approx.distances)                      # default eps 1e-4
    filtered = coll.search(vecs[3], k=5, approx=True, filter_ids=list(range(50)))   # filtered searches stay exact
    assert np.array_equal(filtered.distances, coll.search(vecs[3], k=5, filter_ids=list(range(50))).distances)
    d = np.array([0.25, -0.25, 0.35, np.inf, np.nan], dtype=np.float32)
    _round_to_eps(d, 0.5)
    assert d[:3].tolist() == [0.5, -0.5, 0.5] and np.isinf(d[3]) and np.isnan(d[4])  # halves away from zero
    view = coll.search_range(vecs[3], threshold=-1e6)
    assert isinstance(view, client_mod.ResultView) and len(view.ids) > 0
    ids, dists, fields = view
    assert len(ids) == len(dists)
    assert len(coll.search_range(vecs[3], threshold=1e6).ids) == 0
    assert len(coll.search_range(vecs[3], threshold=-1e6, max_results=3).ids) <= 3


def test_result_view_arrow_export():
    # tests/standard_tests/test_result_view.py: to_arrow columns (id, distance, one per field); polars is optional
    pa = pytest.importorskip("pyarrow")
    rv = client_mod.ResultView(ids=np.array([3, 1], dtype=np.int64), distances=np.array([0.5, 0.25], dtype=np.float32),
                               fields=[{"g": 1}, {"g": 2}], k=2, distance="IP", index="Flat")
    t = rv.to_arrow()
    assert t.column_names == ["id", "distance", "g"] and t.num_rows == 2
    assert t.schema.field("id").type == pa.int64() and t.schema.field("distance").type == pa.float32()
    mixed = client_mod.ResultView(ids=np.array(["a", 7], dtype=object), distances=np.array([0.5, 0.25], dtype=np.float32), k=2)
    assert mixed.to_arrow().column("id").to_pylist() == ["a", "7"]
    data = client_mod.ResultView(ids=np.array([1, 2]), vectors=np.zeros((2, 3), np.float32), result_type="data")
    assert data.to_arrow().column_names == ["id", "vector"]
