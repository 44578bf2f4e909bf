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
        got_ids, got_d = coll.search_range(q, thr, max_results=100)
        order = [i for i in np.argsort(d_all, kind="stable") if d_all[i] <= thr and i != np.argsort(d_all, kind="stable")[3]]
        assert got_ids.tolist() == [int(i) + 5000 for i in order]
        assert np.array_equal(got_d.view(np.uint32), d_all[order].view(np.uint32))
        assert len(coll.search_range(q, thr, max_results=7)[0]) == 7
