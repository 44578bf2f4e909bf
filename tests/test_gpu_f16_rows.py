"""GPU parity for float16 collections: FlatMmap::search / search_filtered on F16 storage score every pair with the
scalar f32-query x f16-row kernels (src/distance/simd.rs:805-1092 via compute_distance_f16, src/distance/mod.rs:217-237;
src/storage/flat_mmap.rs:905-907, :1259-1281, :5047-5180, :511-520, :5329-5437).  `lb_index_search_f16_rows` against the
oracle's restatement: ids and order exact; scores bit-exact for the f32-arithmetic metrics, 1e-5 for the f64 ones."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

F32_ARITH = ["ip", "l2", "cosine", "l1", "chebyshev", "canberra", "bray_curtis"]
F64_ARITH = ["correlation", "hellinger", "wasserstein", "jensen_shannon"]
BINARY = ["hamming", "jaccard", "tanimoto", "dice"]
REL_TOL = 1e-5


@pytest.fixture(scope="module")
def L():
    import lynsedb_b200
    return lynsedb_b200


def _f16(x):
    return np.ascontiguousarray(x.astype(np.float16).astype(np.float32))


def _compare(metric, rows, dists, want_ids, want_d):
    assert np.array_equal(rows, want_ids.astype(np.uint32)), (metric, rows, want_ids)
    if metric in F64_ARITH:
        np.testing.assert_allclose(dists, want_d, rtol=REL_TOL, atol=1e-7)
    else:
        assert np.array_equal(dists.view(np.uint32), want_d.view(np.uint32)), (metric, dists, want_d)


@pytest.mark.parametrize("metric", F32_ARITH + F64_ARITH + BINARY)
@pytest.mark.parametrize("n,dim", [(700, 19), (6000, 64), (9000, 100)])
def test_f16_rows_search_matches_oracle(L, oracle, metric, n, dim):
    rng = np.random.default_rng(n + dim)
    data = _f16(rng.random((n, dim), dtype=np.float32))
    queries = rng.random((3, dim), dtype=np.float32)      # queries stay f32 (simd.rs: f32 query x f16 candidate)
    queries[0] = data[5]
    idx = L.DeviceIndex(dim)
    idx.append(data)
    rows, dists, counts = idx.search(queries, 10, metric, f16_rows=True)
    for i in range(3):
        want_ids, want_d = oracle.store_search_f16(data, queries[i], 10, metric)
        assert counts[i] == len(want_ids)
        _compare(metric, rows[i, :counts[i]], dists[i, :counts[i]], want_ids, want_d)
    idx.close()


def test_f16_rows_order_differs_from_the_f32_kernels(L, oracle):
    # the scalar order is observable: on the same rounded rows the two kernel families disagree in the last bits
    rng = np.random.default_rng(2)
    data = _f16(rng.random((5000, 256), dtype=np.float32))
    q = rng.random((1, 256), dtype=np.float32)
    idx = L.DeviceIndex(256)
    idx.append(data)
    _, d_scalar, _ = idx.search(q, 50, "ip", f16_rows=True)
    _, d_flat, _ = idx.search(q, 50, "ip")
    assert not np.array_equal(d_scalar.view(np.uint32), d_flat.view(np.uint32))
    np.testing.assert_allclose(d_scalar, d_flat, rtol=1e-5)
    idx.close()


@pytest.mark.parametrize("metric", ["ip", "l2", "cosine", "l1", "jensen_shannon", "hamming"])
def test_f16_rows_filtered_search_matches_per_pair_oracle(L, oracle, metric):
    # search_filtered_f16: the same kernels over the allowed rows, (distance, row) order
    rng = np.random.default_rng(9)
    n, dim = 5000, 48
    data = _f16(rng.random((n, dim), dtype=np.float32))
    q = rng.random(dim, dtype=np.float32)
    allowed = np.sort(rng.choice(n, 900, replace=False))
    idx = L.DeviceIndex(dim)
    idx.append(data)
    rows, dists, counts = idx.search(q.reshape(1, -1), 12, metric, L.make_allow_bits(n, allowed), f16_rows=True)
    score = np.array([oracle.compute_distance_f16(q, data[r], metric) for r in allowed], dtype=np.float32)
    key = -score if metric == "ip" else score
    order = np.lexsort((allowed, key))[:12]
    assert counts[0] == 12
    _compare(metric, rows[0], dists[0], allowed[order], score[order])
    idx.close()


def test_f16_collection_search_takes_the_scalar_path_and_batch_the_decoded_one(L, oracle):
    # the reference's own f16 collection cases (src/engine.rs:7972-8076) + parity of both arithmetic paths
    coll = L.Collection("c", 4, dtypes="float16")
    coll.add([10, 11, 12], vectors=np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0.5, 0.5, 0, 0]], dtype=np.float32))
    coll.commit()
    assert coll.search(np.array([1, 0, 0, 0], dtype=np.float32), 2).ids.tolist() == [10, 12]
    res = coll.batch_search(np.array([[1, 0, 0, 0], [0, 1, 0, 0]], dtype=np.float32), 2)
    assert res[0].ids.tolist() == [10, 12] and res[1].ids.tolist() == [11, 12]
    coll.close()

    rng = np.random.default_rng(77)
    n, dim = 6000, 72
    raw = rng.random((n, dim), dtype=np.float32)
    stored = _f16(raw)
    q = rng.random((2, dim), dtype=np.float32)
    for mode, metric in (("FLAT-IP", "ip"), ("FLAT-L2", "l2"), ("FLAT-COS", "cosine")):
        coll = L.Collection("big", dim, dtypes="float16", default_index=mode)
        coll.add(list(range(n)), vectors=raw)
        coll.commit()
        segments = coll._store.segments()
        for i in range(2):
            single = coll.search(q[i], 10)
            want_ids, want_d = oracle.store_search_f16(stored, q[i], 10, metric, segments=segments)
            assert single.ids.tolist() == want_ids.tolist()
            assert np.array_equal(single.distances.view(np.uint32), want_d.view(np.uint32))
        coll.close()


def test_f16_collection_batch_path_scans_the_store_as_one_array(L, oracle):
    # engine.rs:5448-5453: batch_exact_flat_search_f32 over read_all_f32() — a 100-row flush followed by a 6000-row one
    # is ONE 6100-row array, so every row takes the batch-8 inner-product order (no small-segment rule)
    rng = np.random.default_rng(31)
    dim = 40
    raw = rng.random((6100, dim), dtype=np.float32)
    stored = _f16(raw)
    coll = L.Collection("c", dim, dtypes="float16")
    coll.add(list(range(100)), vectors=raw[:100])
    coll.commit()
    coll.add(list(range(100, 6100)), vectors=raw[100:])
    coll.commit()
    assert coll._store.segments() == [6100]
    q = rng.random((3, dim), dtype=np.float32)
    res = coll.batch_search(q, 10)
    ids, dists, counts = oracle.store_batch_search(stored, q, 10, "ip", segment_rows=[6100])
    for i in range(3):
        assert res[i].ids.tolist() == ids[i, :counts[i]].tolist()
        assert np.array_equal(res[i].distances.view(np.uint32), dists[i, :counts[i]].view(np.uint32))
    coll.close()
