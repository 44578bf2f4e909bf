"""GPU parity: IVF build + search against the oracle's restatement of src/index/kmeans.rs / ivf.rs, and the
Collection object model end to end (same shapes as the reference's own tests: src/index/ivf.rs:545-679,
tests/standard_tests/test_search.py, test_collection.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    import lynsedb_b200 as L
    return L


@pytest.fixture(scope="module")
def oracle():
    import oracle
    return oracle


def _formula(n=800, dim=32):
    i, j = np.arange(n).reshape(-1, 1), np.arange(dim).reshape(1, -1)
    return (((i * 131 + j * 17 + 1) % 997).astype(np.float32) / np.float32(997.0) + np.float32(0.01)).astype(np.float32)


@pytest.mark.parametrize("metric,n,dim,nc", [("ip", 800, 32, 32), ("l2", 3000, 24, 40), ("cosine", 2000, 16, 17), ("l2", 5000, 20, 300)])
def test_kmeans_training_is_bit_identical_to_the_oracle(L, oracle, metric, n, dim, nc):
    data = _formula(n, dim) if metric == "ip" else np.random.default_rng(5).random((n, dim), dtype=np.float32)
    cent, assign = oracle.kmeans_train(data, nc, metric)
    with L.DeviceIndex(dim) as idx:
        idx.append(data)
        with L.IVFIndex(idx, metric, n_clusters=nc) as ivf:
            assert ivf.n_centroids == cent.shape[0]
            assert np.array_equal(ivf.assignments(), assign)
            assert np.array_equal(ivf.centroids().view(np.uint32), cent.view(np.uint32))


@pytest.mark.parametrize("metric", ["ip", "l2", "cosine", "l1", "hamming", "tanimoto"])
@pytest.mark.parametrize("nprobe", [1, 3, 64])
def test_ivf_search_matches_oracle(L, oracle, metric, nprobe):
    rng = np.random.default_rng(9)
    n, dim, nc, k = 4000, 40, 24, 10
    data = rng.random((n, dim), dtype=np.float32)
    queries = rng.random((7, dim), dtype=np.float32)
    routing = "l2" if metric in ("hamming", "tanimoto") else metric
    cent, assign = oracle.kmeans_train(data, nc, routing)
    with L.DeviceIndex(dim) as idx:
        idx.append(data)
        with L.IVFIndex(idx, metric, centroids=cent, assignments=assign) as ivf:
            rows, dists, counts = ivf.search(queries, k, nprobe)
    for qi in range(queries.shape[0]):
        o_ids, o_d = oracle.ivf_search(data, cent, assign, queries[qi], k, nprobe, metric)
        c = int(counts[qi])
        assert c == len(o_ids)
        if metric in ("hamming", "tanimoto"):
            assert np.array_equal(dists[qi, :c], o_d)       # massive ties: the reference does not order ids within a tie
        else:
            assert np.array_equal(rows[qi, :c], o_ids), (metric, nprobe, qi)
            if metric == "l1" or metric in ("ip", "l2", "cosine"):
                assert np.array_equal(dists[qi, :c].view(np.uint32), o_d.view(np.uint32))


def test_ivf_reference_cases(L, oracle):
    # full probe == exact (ivf.rs: ivf_ip_recall_improves_with_nprobe)
    data = _formula()
    with L.DeviceIndex(32) as idx:
        idx.append(data)
        with L.IVFIndex(idx, "ip", n_clusters=32) as ivf:
            high, _, _ = ivf.search(data[0], 10, 32)
            low, _, _ = ivf.search(data[0], 10, 2)
        flat, _, _ = idx.search(data[0], 10, "ip")
    assert set(high[0].tolist()) == set(flat[0].tolist())
    assert len(set(low[0].tolist()) & set(flat[0].tolist())) <= 10
    # filtered search with an empty probe falls back to the filtered corpus
    pts = np.array([[0, 0], [0.1, 0], [10, 10], [10.1, 10]], dtype=np.float32)
    with L.DeviceIndex(2) as idx:
        idx.append(pts)
        with L.IVFIndex(idx, "l2", n_clusters=2) as ivf:
            rows, _, counts = ivf.search(np.zeros(2, np.float32), 2, 1, L.make_allow_bits(4, [2, 3]))
    assert counts[0] > 0 and set(rows[0, :counts[0]].tolist()) <= {2, 3}
    # Hamming full probe == flat distances (ivf.rs: ivf_hamming_binary_full_probe_matches_flat_distances)
    i, j = np.arange(256).reshape(-1, 1), np.arange(32).reshape(1, -1)
    bits = (((i * 17 + j * 3) % 2) == 0).astype(np.float32)
    with L.DeviceIndex(32) as idx:
        idx.append(bits)
        with L.IVFIndex(idx, "hamming", n_clusters=16) as ivf:
            _, got, _ = ivf.search(bits[0], 10, 16)
        _, want, _ = idx.search(bits[0], 10, "hamming")
    assert np.array_equal(got, want)


FLAT_MODES = ["FLAT-IP", "FLAT-L2", "FLAT-COS", "FLAT-L1", "FLAT-CORRELATION", "FLAT-HELLINGER", "FLAT-WASSERSTEIN",
              "FLAT-JENSEN-SHANNON", "FLAT-CHEBYSHEV", "FLAT-CANBERRA", "FLAT-BRAY-CURTIS"]


@pytest.mark.parametrize("mode", FLAT_MODES)
def test_collection_every_flat_metric_finds_self(L, mode):
    # tests/standard_tests/test_search.py:30-98 (default_rng(20260620), 32 x 16)
    rng = np.random.default_rng(20260620)
    data = rng.random((32, 16), dtype=np.float32) + np.float32(0.05)
    if mode == "FLAT-IP":
        data[7] *= 4.0                                      # IP has no self-hit property: make row 7 dominant for itself
    with L.VectorDBClient() as client:
        coll = client.create_collection("db", "c", dim=16, default_index=mode)
        coll.add(vectors=data)
        r = coll.search(data[7], k=3)
        assert r.ids[0] == 7 and coll.index_mode == mode
        if mode != "FLAT-IP":
            assert abs(float(r.distances[0])) <= 1e-5
        assert r.index_type == "Flat"


def test_collection_matches_flat_oracle_with_pending_and_tombstones(L, oracle):
    rng = np.random.default_rng(3)
    dim, k = 24, 8
    a, b = rng.random((12000, dim), dtype=np.float32), rng.random((500, dim), dtype=np.float32)
    queries = rng.random((5, dim), dtype=np.float32)
    with L.VectorDBClient() as client:
        coll = client.create_collection("db", "c", dim=dim, default_index="FLAT-L2")
        coll.add(vectors=a, batch_size=5000)                # 10 000 rows flush as one segment, 2 000 stay pending
        coll.add(vectors=b)
        assert coll.stats()["segments"] == [10000] and coll.stats()["pending_rows"] == 2500
        allv = np.concatenate([a, b])
        res = coll.batch_search(queries, k)
        o_ids, o_d, _ = oracle.store_batch_search(allv, queries, k, "l2", n_threads=1)
        for i, r in enumerate(res):
            assert np.array_equal(r.ids, o_ids[i].astype(np.int64))
            np.testing.assert_allclose(r.distances, o_d[i], rtol=1e-5)
        dead = [int(x) for x in res[0].ids[:3]]
        coll.delete(dead)
        r = coll.search(queries[0], k)
        want = [int(x) for x in oracle.store_batch_search(allv, queries[:1], k + 3, "l2", n_threads=1)[0][0] if int(x) not in dead][:k]
        assert list(r.ids) == want
        coll.commit()
        assert coll.stats()["segments"] == [12500]          # the flush joins the open segment (under the 256 MiB target)
        assert list(coll.search(queries[0], k).ids) == want


def test_collection_with_hundreds_of_deleted_rows_stays_exact(L, oracle):
    # k + |tombstones| > 256: the client masks the deleted rows out and asks for k (tensor-core plan with a row filter)
    rng = np.random.default_rng(8)
    n, dim, k = 30000, 48, 10
    data = rng.random((n, dim), dtype=np.float32) - np.float32(0.5)      # signed: the queries' best rows barely overlap
    queries = rng.random((20, dim), dtype=np.float32) - np.float32(0.5)
    with L.VectorDBClient() as client:
        coll = client.create_collection("db", "c", dim=dim, default_index="FLAT-IP")
        coll.add(vectors=data)
        coll.commit()
        o_ids, _, _ = oracle.store_batch_search(data, queries, 400, "ip", n_threads=1)
        dead = sorted({int(x) for x in o_ids[:, :15].ravel()})           # the best 15 of every query: ~300 rows
        assert len(dead) + k > 256
        coll.delete(dead)
        res = coll.batch_search(queries, k)
        assert coll._store.last_stats()["plan_used"] == 1
        for i, r in enumerate(res):
            want = [int(x) for x in o_ids[i] if int(x) not in set(dead)][:k]
            assert r.ids.tolist() == want


def test_collection_ivf_mode(L):
    rng = np.random.default_rng(11)
    data = rng.random((3000, 32), dtype=np.float32)
    with L.VectorDBClient() as client:
        coll = client.create_collection("db", "ivf", dim=32, default_index=None)
        coll.add(vectors=data)
        coll.build_index("IVF-IP", n_clusters=16, nprobe=4)
        exact = np.argsort(-(data @ data[5]), kind="stable")[:10]
        full = coll.search(data[5], k=10, nprobe=16)        # nprobe == n_clusters -> exact
        assert set(full.ids.tolist()) == set(exact.tolist()) and full.index_type == "IVF"
        some = coll.search(data[5], k=10, nprobe=2)
        assert len(some) == 10


def test_load_lynsedb_directory_and_search_range(L, oracle, tmp_path):
    import json

    rng = np.random.default_rng(21)
    dim = 12
    blocks = [rng.random((700, dim), dtype=np.float32), rng.random((301, dim), dtype=np.float32)]
    (tmp_path / "vector_segments").mkdir()
    segs = []
    for i, b in enumerate(blocks):
        name = f"vector_segments/seg_{i}.bin"
        b.astype("<f4").tofile(tmp_path / name)
        segs.append({"file": name, "rows": b.shape[0]})
    (tmp_path / "vector_manifest.json").write_text(json.dumps({"version": 1, "generation": 1, "id_map_file": "id_map.bin", "segments": segs}))
    ids = np.arange(5000, 6001, dtype="<u8")
    ids.tofile(tmp_path / "id_map.bin")
    allv = np.concatenate(blocks)
    with L.VectorDBClient() as client:
        coll = client.create_collection("db", "disk", dim=dim, default_index="FLAT-L2")
        assert coll.load_lynsedb_directory(tmp_path) == 1001 and coll.shape == (1001, dim)
        q = rng.random(dim, dtype=np.float32)
        r = coll.search(q, k=5)
        want = oracle.store_batch_search(allv, q, 5, "l2", segment_rows=[700, 301], n_threads=1)
        assert list(r.ids) == [int(x) + 5000 for x in want[0][0]]
        # search_range: every live row within the threshold, per-pair kernel order, best first
        d_all = np.array([oracle.compute_distance(q, v, "l2") for v in allv], dtype=np.float32)
        thr = float(np.sort(d_all)[40])
        coll.delete([int(np.argsort(d_all, kind="stable")[3]) + 5000])
        got = coll.search_range(q, thr, max_results=100)
        assert isinstance(got, L.ResultView)
        got_ids, got_d = got.ids, got.distances
        order = [i for i in np.argsort(d_all, kind="stable") if d_all[i] <= thr and i != np.argsort(d_all, kind="stable")[3]]
        assert got_ids.tolist() == [int(i) + 5000 for i in order]
        assert np.array_equal(got_d.view(np.uint32), d_all[order].view(np.uint32))
        assert len(coll.search_range(q, thr, max_results=7).ids) == 7


# ---- standalone IVF_FLAT index: `_core.IvfFlatIndex` (src/python/mod.rs:2049-2156, src/storage/ivf_flat_mmap.rs) ----
_IVF_FLAT_12 = np.array([
    1.0, 0.1, 0.0, 0.0, 0.9, 0.0, 0.1, 0.0, 1.0, 0.0, 0.0, 0.1, 0.8, 0.1, 0.1, 0.0,
    0.0, 1.0, 0.1, 0.0, 0.1, 0.9, 0.0, 0.0, 0.0, 1.0, 0.0, 0.1, 0.1, 0.8, 0.1, 0.0,
    0.0, 0.0, 1.0, 0.1, 0.0, 0.1, 0.9, 0.0, 0.1, 0.0, 1.0, 0.0, 0.0, 0.0, 0.8, 0.1], dtype=np.float32).reshape(12, 4)


def test_ivf_flat_index_reference_cases(L, oracle, tmp_path):
    # test_ivf_flat_build_and_search (ivf_flat_mmap.rs:674-719)
    idx = L.IvfFlatIndex.build(str(tmp_path / "vectors.bin"), _IVF_FLAT_12, dim=4, n_partitions=3, n_iters=10, metric="ip")
    assert len(idx) == 12 and idx.dim == 4 and idx.n_partitions == 3
    q = np.array([1.0, 0.0, 0.0, 0.0], dtype=np.float32)
    ids, dists = idx.search(q, 3, 1, "ip")
    assert ids.dtype == np.uint32 and dists.dtype == np.float32 and len(ids) == 3 and ids[0] <= 3
    ids0, dists0 = idx.search(q, 3, 0, "ip")
    assert np.array_equal(ids0, ids) and np.array_equal(dists0, dists)
    cent, assign = oracle.kmeans_train(_IVF_FLAT_12, 3, "l2", max_iter=10)
    want_ids, want_d = oracle.ivf_flat_search(_IVF_FLAT_12, cent, assign, q, 3, 1, "ip")
    assert np.array_equal(ids, want_ids) and np.array_equal(dists.view(np.uint32), want_d.view(np.uint32))
    # test_ivf_flat_reopen (:751-773)
    data = np.array([1.0, 0.0, 0.0, 1.0, -1.0, 0.0, 0.0, -1.0], dtype=np.float32).reshape(4, 2)
    path = str(tmp_path / "small.bin")
    L.IvfFlatIndex.build(path, data, 2, 2, 5, "ip").close()
    again = L.IvfFlatIndex.open(path, 2)
    assert len(again) == 4 and again.n_partitions == 2
    assert again.search(np.array([1.0, 0.0], dtype=np.float32), 1, 2, "ip")[0][0] == 0
    with pytest.raises(IOError, match="dimension mismatch"):
        L.IvfFlatIndex.open(path, 3)
    # test_ivf_flat_rejects_invalid_build_inputs (:721-749) + the pyo3 wrapper's own checks (mod.rs:2075-2090)
    with pytest.raises(IOError):
        L.IvfFlatIndex.build(str(tmp_path / "x.bin"), np.zeros((2, 2), dtype=np.float32), 2, 0, 5, "l2")
    with pytest.raises(IOError):
        L.IvfFlatIndex.build(str(tmp_path / "x.bin"), np.zeros((2, 2), dtype=np.float32), 2, 3, 5, "l2")
    with pytest.raises(ValueError, match="dimension mismatch"):
        L.IvfFlatIndex.build(str(tmp_path / "x.bin"), np.zeros((2, 3), dtype=np.float32), 2, 1, 5, "l2")
    with pytest.raises(ValueError, match="Unknown metric"):
        L.IvfFlatIndex.build(str(tmp_path / "x.bin"), np.zeros((2, 2), dtype=np.float32), 2, 1, 5, "nope")
    with pytest.raises(ValueError, match="dimension mismatch"):
        idx.search(np.zeros(5, dtype=np.float32), 3, 1, "ip")


@pytest.mark.parametrize("metric,dim,nc,nprobe", [("ip", 64, 64, 8), ("ip", 96, 80, 40), ("ip", 32, 64, 6), ("l2", 64, 64, 8),
                                                  ("cosine", 48, 20, 5), ("hamming", 64, 16, 4), ("l1", 24, 12, 3),
                                                  ("ip", 64, 64, 64)])
def test_ivf_flat_index_matches_oracle(L, oracle, tmp_path, metric, dim, nc, nprobe):
    rng = np.random.default_rng(dim * 1000 + nc)
    data = rng.random((6000, dim), dtype=np.float32)
    data[:, 3] *= 5.0
    queries = rng.random((12, dim), dtype=np.float32)
    path = str(tmp_path / "ivf.bin")
    idx = L.IvfFlatIndex.build(path, data, dim, n_partitions=nc, n_iters=6, metric=metric)
    cent, assign = oracle.kmeans_train(data, nc, "l2", max_iter=6)
    assert np.array_equal(idx._ivf.centroids().view(np.uint32), cent.view(np.uint32))
    assert np.array_equal(idx._ivf.assignments(), assign)
    reopened = L.IvfFlatIndex.open(path, dim)
    for q in queries:
        want_ids, want_d = oracle.ivf_flat_search(data, cent, assign, q, 10, nprobe, metric)
        for index in (idx, reopened):
            ids, dists = index.search(q, 10, nprobe, metric)
            assert np.array_equal(dists.view(np.uint32), want_d.view(np.uint32))
            # tied distances: the reference's order is arbitrary; ours is the lower row first, as the oracle's
            assert np.array_equal(ids, want_ids)


def test_ivf_flat_index_files_follow_the_reference_layout(L, tmp_path):
    # save_metadata (ivf_flat_mmap.rs:450-483): u64 dim, n, partitions; f32 centroids; u64 offsets[p+1]; u32 original ids;
    # the data file holds the rows partition by partition, in row order inside a partition (:116-131)
    rng = np.random.default_rng(3)
    data = rng.random((500, 8), dtype=np.float32)
    path = tmp_path / "layout.bin"
    idx = L.IvfFlatIndex.build(str(path), data, 8, n_partitions=7, n_iters=4)
    meta = (tmp_path / "layout.ivf_meta.bin").read_bytes()
    dim, n, p = np.frombuffer(meta[:24], dtype="<u8")
    assert (dim, n, p) == (8, 500, 7)
    cent = np.frombuffer(meta[24:24 + 4 * 7 * 8], dtype="<f4").reshape(7, 8)
    off = np.frombuffer(meta[24 + 224:24 + 224 + 64], dtype="<u8")
    orig = np.frombuffer(meta[24 + 224 + 64:], dtype="<u4")
    assert orig.size == 500 and off[0] == 0 and off[-1] == 500 and np.all(np.diff(off.astype(np.int64)) >= 0)
    assign = idx._ivf.assignments()
    assert np.array_equal(cent, idx._ivf.centroids())
    for part in range(7):
        members = orig[int(off[part]):int(off[part + 1])]
        assert np.array_equal(members, np.flatnonzero(assign == part).astype(np.uint32))
    stored = np.fromfile(path, dtype="<f4").reshape(500, 8)
    assert np.array_equal(stored, data[orig])


def test_flat_index_write_appends(L, tmp_path):
    # test_flat_mmap_append (src/storage/flat_mmap.rs:6057-6075): write() appends rows, to the index and to the file
    path = tmp_path / "vectors.bin"
    store = L.FlatIndex(str(path), 2)
    assert len(store) == 0
    store.write(np.array([[1.0, 2.0], [3.0, 4.0]], dtype=np.float32))
    assert len(store) == 2
    store.write(np.array([[5.0, 6.0]], dtype=np.float32))
    assert len(store) == 3
    assert np.array_equal(np.fromfile(path, dtype="<f4"), np.array([1, 2, 3, 4, 5, 6], dtype=np.float32))
    ids, dists = store.search(np.array([1.0, 1.0], dtype=np.float32), 3, "ip")
    assert ids.tolist() == [2, 1, 0] and dists.tolist() == [11.0, 7.0, 3.0]
    reopened = L.FlatIndex(str(path), 2)
    assert len(reopened) == 3
