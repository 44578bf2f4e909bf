"""Pins the CPU oracle against every known-answer test the reference holds for
the distance + top-k path.  The reference's assertions are
replayed here as DATA (inputs, expected values, tolerances); citations are
relative to the reference tree.  CPU-only; runs in seconds.
"""
import math

import numpy as np
import pytest

F = np.float32


def arr(*v):
    return np.asarray(v, dtype=F)


# ---------------------------------------------------------------- simd.rs:2906-2987
def test_ip_l2_cosine_known_answers(oracle):
    assert abs(oracle.compute_distance(arr(1, 2, 3, 4), arr(4, 3, 2, 1), "ip") - 20.0) < 1e-5
    assert abs(oracle.compute_distance(arr(1, 0, 0), arr(0, 1, 0), "l2") - 2.0) < 1e-5
    assert abs(oracle.compute_distance(arr(1, 0, 0), arr(1, 0, 0), "cosine")) < 1e-5
    assert abs(oracle.compute_distance(arr(1, 0), arr(0, 1), "cosine") - 1.0) < 1e-5


def test_domain_distances_known_answers(oracle):
    d = oracle.compute_distance
    assert abs(d(arr(1, 2), arr(4, 0), "l1") - 5.0) < 1e-6
    assert d(arr(1, 2, 3), arr(2, 4, 6), "correlation") < 1e-6
    assert abs(d(arr(1, 2, 3), arr(3, 2, 1), "correlation") - 2.0) < 1e-6
    assert d(arr(1, 0), arr(1, 0), "hellinger") < 1e-6
    assert abs(d(arr(1, 0), arr(0, 1), "hellinger") - 1.0) < 1e-6
    assert abs(d(arr(1, 0, 0), arr(0, 0, 1), "wasserstein") - 2.0) < 1e-6
    assert abs(d(arr(1, 1, 0), arr(1, 0, 1), "dice") - 0.5) < 1e-6
    assert abs(d(arr(1, 0), arr(0, 1), "jensen_shannon") - math.sqrt(F(math.log(2.0)))) < 1e-6
    assert abs(d(arr(1, 2, 3), arr(4, 0, 3), "chebyshev") - 3.0) < 1e-6
    assert abs(d(arr(1, 0, 3), arr(2, 0, 1), "canberra") - 5.0 / 6.0) < 1e-6
    assert abs(d(arr(1, 2), arr(2, 4), "bray_curtis") - 1.0 / 3.0) < 1e-6
    assert math.isinf(d(arr(-1, 2), arr(1, 2), "jensen_shannon"))


# ---------------------------------------------------------------- test_backend.py:71-104
@pytest.mark.parametrize(
    ("metric", "a", "b", "expected"),
    [
        ("l1", [1, 2, 3], [3, 0, 4], 5.0),
        ("correlation", [1, 2, 3], [2, 4, 6], 0.0),
        ("hellinger", [1, 0], [0, 1], 1.0),
        ("wasserstein", [1, 0, 0], [0, 0, 1], 2.0),
        ("dice", [1, 1, 0], [1, 0, 1], 0.5),
        ("tanimoto", [1, 1, 0], [1, 0, 1], 2.0 / 3.0),
        ("jensen_shannon", [1, 0], [0, 1], math.sqrt(math.log(2.0))),
        ("chebyshev", [1, 2, 3], [4, 0, 3], 3.0),
        ("canberra", [1, 0, 3], [2, 0, 1], 5.0 / 6.0),
        ("bray_curtis", [1, 2], [2, 4], 1.0 / 3.0),
    ],
)
def test_python_operator_known_answers(oracle, metric, a, b, expected):
    assert oracle.compute_distance(arr(*a), arr(*b), metric) == pytest.approx(expected, abs=1e-5)


def test_haversine_geojson_order_and_meters(oracle):
    shanghai, beijing = arr(121.4737, 31.2304), arr(116.4074, 39.9042)
    meters = oracle.compute_distance(shanghai, beijing, "haversine")
    assert abs(meters - 1_067_000.0) < 10_000.0
    assert oracle.compute_distance(shanghai, shanghai, "haversine") == 0.0
    assert math.isinf(oracle.compute_distance(arr(1, 1, 1), arr(1, 1, 1), "haversine"))
    assert math.isinf(oracle.compute_distance(arr(0, 91), arr(0, 0), "haversine"))


# ---------------------------------------------------------------- simd.rs:2989-3088
def _js_reference_f64(a, b):
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    sa, sb = a64.sum(), b64.sum()
    div = 0.0
    for av, bv in zip(a64, b64):
        p, q = av / sa, bv / sb
        m = 0.5 * (p + q)
        if p > 0:
            div += 0.5 * p * math.log(p / m)
        if q > 0:
            div += 0.5 * q * math.log(q / m)
    return F(math.sqrt(max(div, 0.0)))


def test_jensen_shannon_simd_matches_f64_reference(oracle):
    state = 0x12345678
    M = 0xFFFFFFFF
    for dim in [1, 3, 4, 7, 8, 16, 127, 128, 257]:
        for rnd in range(20):
            a, b = np.zeros(dim, F), np.zeros(dim, F)
            for i in range(dim):
                state = (state * 1664525 + 1013904223) & M
                av = F(F(state >> 8) + F(1.0)) / F(16777216.0)
                state = (state * 1664525 + 1013904223) & M
                bv = F(F(state >> 8) + F(1.0)) / F(16777216.0)
                a[i] = 0.0 if (i + rnd) % 11 == 0 else av
                b[i] = 0.0 if (i + rnd) % 13 == 0 else bv
            if not a.any():
                a[0] = 1.0
            if not b.any():
                b[0] = 1.0
            actual = oracle.compute_distance(a, b, "jensen_shannon")
            expected = float(_js_reference_f64(a, b))
            assert abs(actual - expected) <= 2e-5, (dim, actual, expected)
            assert abs(actual - oracle.compute_distance(b, a, "jensen_shannon")) <= 2e-5
            (inv_a, ent_a), (inv_b, ent_b) = oracle.probability_row_stats(np.stack([a, b]))
            na = a * inv_a
            cached = oracle.jensen_shannon_precomputed(na, b, ent_a, inv_b, ent_b)
            assert abs(cached - expected) <= 3e-5, (dim, cached, expected)
            b0 = math.sqrt(oracle.jensen_shannon_precomputed(na, b, ent_a, inv_b, ent_b, divergence=True))
            b1 = math.sqrt(oracle.jensen_shannon_precomputed(na, a, ent_a, inv_a, ent_a, divergence=True))
            assert abs(b0 - expected) <= 3e-5
            assert b1 <= 1e-6

    near_a = np.arange(1, 129, dtype=F)
    near_b = np.asarray([v * F(1.0 + (F(i) % F(3.0) - F(1.0)) * F(1e-4)) for i, v in enumerate(near_a)], dtype=F)
    (inv_a, ent_a), (inv_b, ent_b) = oracle.probability_row_stats(np.stack([near_a, near_b]))
    near = oracle.jensen_shannon_precomputed(near_a * inv_a, near_b, ent_a, inv_b, ent_b)
    assert abs(near - float(_js_reference_f64(near_a, near_b))) <= 2e-5

    tiny = np.frombuffer(np.uint32(1).tobytes(), dtype=F)[0]
    assert abs(oracle.compute_distance(arr(tiny, 0), arr(0, tiny), "jensen_shannon") - math.sqrt(F(math.log(2.0)))) <= 1e-6


# ---------------------------------------------------------------- simd.rs:3090-3114
def test_high_dim_ip_matches_f64_scalar(oracle):
    dim = 768
    a = (np.arange(dim, dtype=F) * F(0.001)).astype(F)
    b = ((dim - np.arange(dim)).astype(F) * F(0.001)).astype(F)
    expected = float(np.dot(a.astype(np.float64), b.astype(np.float64)))
    assert abs(oracle.compute_distance(a, b, "ip") - expected) < 1e-2
    assert abs(oracle.inner_product_batch8_order(a, b) - oracle.compute_distance(a, b, "ip")) < 1e-3


# ---------------------------------------------------------------- distance/mod.rs:502-621
def test_top_k_ip_l2(oracle):
    cands = np.asarray([[1, 0, 0, 0], [0.5, 0.5, 0, 0], [0, 1, 0, 0]], dtype=F)
    ids, dists = oracle.top_k_search(arr(1, 0, 0, 0), cands, "ip", 2)
    assert len(ids) == 2 and ids[0] == 0 and abs(dists[0] - 1.0) < 1e-6
    cands = np.asarray([[1, 0, 0], [0.1, 0, 0], [2, 0, 0]], dtype=F)
    ids, _ = oracle.top_k_search(arr(0, 0, 0), cands, "l2", 2)
    assert ids[0] == 1


def test_top_k_larger(oracle):
    dim, n = 16, 1000
    cands = (np.arange(n * dim, dtype=F) * F(0.001)).reshape(n, dim)
    ids, dists = oracle.top_k_search(cands[0], cands, "l2", 5)
    assert len(ids) == 5 and ids[0] == 0 and dists[0] < 1e-6
    assert np.all(np.diff(dists) >= 0)


def test_top_k_empty_zero_k_and_clamp(oracle):
    ids, dists = oracle.top_k_search(arr(1, 2), np.zeros((0, 2), F), "l2", 5)
    assert len(ids) == 0 and len(dists) == 0
    ids, dists = oracle.top_k_search(arr(1, 2), np.asarray([[1, 2], [3, 4]], F), "l2", 0)
    assert len(ids) == 0
    ids, dists = oracle.top_k_search(arr(0, 0), np.asarray([[2, 0], [1, 0]], F), "l2", 10)
    assert ids.tolist() == [1, 0] and dists[0] <= dists[1]


def test_top_k_binary_metrics_sorted_by_lower_distance(oracle):
    q = arr(1, 0, 1, 0)
    cands = np.asarray([[1, 0, 1, 0], [1, 1, 1, 0], [0, 1, 0, 1]], dtype=F)
    ids, dists = oracle.top_k_search(q, cands, "hamming", 3)
    assert ids.tolist() == [0, 1, 2] and dists.tolist() == [0.0, 1.0, 4.0]
    ids, dists = oracle.top_k_search(q, cands, "jaccard", 3)
    assert ids.tolist() == [0, 1, 2]
    assert abs(dists[0]) < 1e-6 and abs(dists[1] - 1 / 3) < 1e-6 and abs(dists[2] - 1.0) < 1e-6


# ---------------------------------------------------------------- flat_mmap.rs:6021-6106
def test_flat_write_search(oracle):
    data = np.asarray([[1, 0, 0, 0], [0, 1, 0, 0], [0.5, 0.5, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=F)
    ids, dists = oracle.flat_search(data, arr(1, 0, 0, 0), 2, "ip")
    assert len(ids) == 2 and ids[0] == 0 and abs(dists[0] - 1.0) < 1e-6
    ids, _ = oracle.flat_search(data, arr(0, 0, 0, 0), 1, "l2")
    assert ids[0] == 2


def test_flat_jensen_shannon_cache_exact_and_after_append(oracle):
    rows = np.asarray([[1, 2, 3, 4], [4, 3, 2, 1], [0, 0, 0, 0]], dtype=F)
    ids, dists = oracle.flat_search(rows, arr(1, 2, 3, 4), 1, "jensen_shannon")
    assert ids.tolist() == [0] and dists[0] <= 3e-5
    rows = np.vstack([rows, arr(0.5, 1.5, 2.5, 3.5)])
    ids, dists = oracle.flat_search(rows, arr(0.5, 1.5, 2.5, 3.5), 1, "jensen_shannon")
    assert ids.tolist() == [3] and dists[0] <= 3e-5


def test_flat_reopen_l2(oracle):
    ids, _ = oracle.flat_search(np.asarray([[1, 2, 3], [4, 5, 6]], F), arr(1, 2, 3), 1, "l2")
    assert ids[0] == 0


# ---------------------------------------------------------------- flat_mmap.rs:6385-6421
def test_packed_binary_matches_thresholded_f32(oracle):
    dim = 130
    rows = np.zeros((3, dim), dtype=F)
    for index in [0, 1, 64, 129]:
        rows[0, index] = 1.0
        rows[1, index] = 1.0
    rows[1, 5] = 1.0
    for index in [2, 3, 65]:
        rows[2, index] = 1.0
    query = rows[0].copy()
    for metric in ["hamming", "jaccard", "tanimoto", "dice"]:
        p_ids, p_d = oracle.flat_search(rows, query, 3, metric)
        r_ids, r_d = oracle.top_k_search(query, rows, metric, 3)
        assert p_ids.tolist() == r_ids.tolist(), metric
        assert np.all(np.abs(p_d - r_d) < 1e-6), metric
    words = oracle.pack_binary(rows)
    assert words.shape == (3, 3) and words.nbytes == 3 * 3 * 8
    # LSB-first layout: bit i of the row -> word i // 64, bit i % 64
    assert int(words[0, 0]) == 0b11 and int(words[0, 1]) == 1 and int(words[0, 2]) == 1 << 1


# ---------------------------------------------------------------- vector_store.rs:1309-1353
def test_segmented_search_picks_global_best(oracle):
    data = np.arange(400, dtype=F).reshape(100, 4)
    both = np.vstack([data, data])
    ids, _, counts = oracle.store_batch_search(both, arr(0, 1, 2, 3), 1, "l2", segment_rows=[100, 100])
    assert counts[0] == 1 and ids[0, 0] == 0  # tie between row 0 and row 100 -> lower row


def test_segment_merge_ranks_raw_scores(oracle):
    first = np.full((300, 1), 0.1, dtype=F)
    first[0] = 0.6
    second = np.full((300, 1), 0.1, dtype=F)
    second[0] = 0.9
    ids, dists, _ = oracle.store_batch_search(np.vstack([first, second]), arr(1.0), 1, "ip", segment_rows=[300, 300])
    assert ids[0, 0] == 300 and abs(dists[0, 0] - 0.9) < 1e-7


# ---------------------------------------------------------------- cluster.rs:674-720 ordering contract
def test_merge_ordering_ip_desc_l2_asc_with_row_tiebreak(oracle):
    rows = np.asarray([[0.2], [0.9], [0.4], [0.9]], dtype=F)
    ids, dists, _ = oracle.store_batch_search(rows, arr(1.0), 3, "ip", segment_rows=[2, 2])
    assert ids[0].tolist() == [1, 3, 2]
    ids, dists, _ = oracle.store_batch_search(rows, arr(0.0), 3, "l2", segment_rows=[2, 2])
    assert ids[0].tolist() == [0, 2, 1]


# ---------------------------------------------------------------- test_backend.py:107-189
def test_numpy_agreement_on_seed7_matrix(oracle):
    np.random.seed(7)
    vecs = np.random.rand(200, 16).astype(F)
    np.random.seed(1)
    q = np.random.rand(16).astype(F)
    ids, _ = oracle.top_k_search(q, vecs, "ip", 1)
    assert int(ids[0]) == int(np.argmax(vecs @ q))
    ids, _ = oracle.top_k_search(q, vecs, "l2", 1)
    assert int(ids[0]) == int(np.argmin(((vecs - q) ** 2).sum(axis=1)))
    ids, dists = oracle.top_k_search(q, vecs, "ip", 300)
    assert len(ids) == 200
    eye = np.eye(16, dtype=F)
    for i in range(16):
        ids, d = oracle.top_k_search(eye[i], eye, "ip", 1)
        assert int(ids[0]) == i and abs(d[0] - 1.0) < 1e-5


# ---------------------------------------------------------------- test_search.py:30-98 (every FLAT metric finds self)
@pytest.mark.parametrize("metric", ["ip", "l2", "cosine", "l1", "correlation", "hellinger", "wasserstein",
                                    "jensen_shannon", "chebyshev", "canberra", "bray_curtis"])
def test_every_flat_metric_finds_self(oracle, metric):
    rng = np.random.default_rng(20260620)
    rows = rng.random((32, 16), dtype=F) + F(0.05)
    if metric == "ip":
        rows /= np.linalg.norm(rows, axis=1, keepdims=True)
    target = 7
    ids, dists = oracle.flat_search(rows, rows[target], 1, metric)
    assert ids[0] == target
    if metric != "ip":
        assert dists[0] <= 1e-5


# ---------------------------------------------------------------- chunked parallel scan == sequential definition
@pytest.mark.parametrize("metric", ["ip", "l2", "cosine", "hamming", "jensen_shannon"])
@pytest.mark.parametrize("threads", [1, 3, 8])
def test_parallel_scan_matches_bruteforce_order(oracle, metric, threads):
    rng = np.random.default_rng(5)
    n, dim, k = 9000, 24, 17
    rows = rng.random((n, dim), dtype=F)
    rows[4000] = rows[10]  # exact duplicate -> tie resolved by row index
    q = rng.random(dim, dtype=F)
    ids, dists = oracle.flat_search(rows, q, k, metric, n_threads=threads)
    assert len(ids) == k
    sign = -1.0 if metric == "ip" else 1.0
    order = sorted(range(k), key=lambda i: (sign * float(dists[i]), int(ids[i])))
    assert order == list(range(k))
    if metric in ("l2", "cosine"):  # single summation order -> independent of chunking
        all_d = np.asarray([oracle.compute_distance(q, r, metric) for r in rows], dtype=F)
        expect = sorted(range(n), key=lambda i: (float(all_d[i]), i))[:k]
        assert ids.tolist() == expect
        assert np.array_equal(dists, all_d[expect])


# ---- IVF (src/index/ivf.rs:545-679) --------------------------------------------------------------------------------
def _ivf_formula_data(n=800, dim=32):
    i = np.arange(n).reshape(-1, 1)
    j = np.arange(dim).reshape(1, -1)
    return (((i * 131 + j * 17 + 1) % 997).astype(np.float32) / np.float32(997.0) + np.float32(0.01)).astype(np.float32)


def test_ivf_filtered_empty_probe_does_not_leak(oracle):
    data = np.array([[0, 0], [0.1, 0], [10, 10], [10.1, 10]], dtype=np.float32)
    cent, assign = oracle.kmeans_train(data, 2, "l2")
    allow = np.array([0b1100], dtype=np.uint64)           # only rows 2 and 3 (the far cluster) are allowed
    ids, _ = oracle.ivf_search(data, cent, assign, [0.0, 0.0], 2, 1, "l2", allow)
    assert len(ids) > 0 and set(ids.tolist()) <= {2, 3}


def test_ivf_ip_full_probe_is_exact_and_recall_grows(oracle):
    data = _ivf_formula_data()
    cent, assign = oracle.kmeans_train(data, 32, "ip")
    assert cent.shape == (32, 32)
    q = data[0]
    exact = set(np.argsort(-(data.astype(np.float64) @ q.astype(np.float64)), kind="stable")[:10].tolist())
    low, _ = oracle.ivf_search(data, cent, assign, q, 10, 2, "ip")
    high, _ = oracle.ivf_search(data, cent, assign, q, 10, 32, "ip")
    rec_low, rec_high = len(exact & set(low.tolist())) / 10, len(exact & set(high.tolist())) / 10
    assert rec_high >= rec_low and rec_high == 1.0


def test_ivf_hamming_full_probe_matches_flat_distances(oracle):
    n, dim = 256, 32
    i, j = np.arange(n).reshape(-1, 1), np.arange(dim).reshape(1, -1)
    data = (((i * 17 + j * 3) % 2) == 0).astype(np.float32)
    cent, assign = oracle.kmeans_train(data, 16, "l2")    # binary metrics route with L2 (ivf.rs:80-87)
    q = data[0]
    want = np.sort(np.array([oracle.compute_distance(q, data[r], "hamming") for r in range(n)], dtype=np.float32), kind="stable")[:10]
    _, got = oracle.ivf_search(data, cent, assign, q, 10, 16, "hamming")
    assert np.array_equal(got, want)


# ---- the same pins as a committed fixture (tests/golden/) -------------------------------------------------------------
def _golden():
    import json
    from pathlib import Path

    return json.loads((Path(__file__).parent / "golden" / "reference_known_answers.json").read_text())


def test_golden_fixture_compute_distance(oracle):
    for case in _golden()["compute_distance"]:
        got = oracle.compute_distance(np.asarray(case["a"], F), np.asarray(case["b"], F), case["metric"])
        assert abs(got - case["expected"]) <= case["abs_tol"], case


def test_golden_fixture_top_k(oracle):
    for case in _golden()["top_k_search"]:
        ids, dists = oracle.top_k_search(np.asarray(case["query"], F), np.asarray(case["candidates"], F), case["metric"], case["k"])
        assert ids.tolist() == case["ids"], case
        assert np.allclose(dists, case["dists"], atol=1e-6), case


# ---- standalone IVF_FLAT (src/storage/ivf_flat_mmap.rs:674-815) ------------------------------------------------------
_IVF_FLAT_12 = np.array([
    1.0, 0.1, 0.0, 0.0, 0.9, 0.0, 0.1, 0.0, 1.0, 0.0, 0.0, 0.1, 0.8, 0.1, 0.1, 0.0,
    0.0, 1.0, 0.1, 0.0, 0.1, 0.9, 0.0, 0.0, 0.0, 1.0, 0.0, 0.1, 0.1, 0.8, 0.1, 0.0,
    0.0, 0.0, 1.0, 0.1, 0.0, 0.1, 0.9, 0.0, 0.1, 0.0, 1.0, 0.0, 0.0, 0.0, 0.8, 0.1], dtype=np.float32).reshape(12, 4)


def _ivf_flat_lcg_data(n=1000, dim=8):
    """test_ivf_flat_recall's generator (ivf_flat_mmap.rs:786-792)."""
    rng, out = 42, np.zeros(n * dim, dtype=np.float32)
    for i in range(n * dim):
        rng = (rng * 6364136223846793005 + 1) & 0xFFFFFFFFFFFFFFFF
        out[i] = np.float32(np.float32(np.float32(rng >> 33) / np.float32(0xFFFFFFFF)) * np.float32(2.0)) - np.float32(1.0)
    return out.reshape(n, dim)


def test_ivf_flat_build_and_search_pin(oracle):
    # test_ivf_flat_build_and_search: 3 partitions trained with 10 L2 iterations; top hit in the first cluster;
    # nprobe 0 behaves as nprobe 1
    cent, assign = oracle.kmeans_train(_IVF_FLAT_12, 3, "l2", max_iter=10)
    q = np.array([1.0, 0.0, 0.0, 0.0], dtype=np.float32)
    ids, dists = oracle.ivf_flat_search(_IVF_FLAT_12, cent, assign, q, 3, 1, "ip")
    assert len(ids) == 3 and ids[0] <= 3
    ids0, dists0 = oracle.ivf_flat_search(_IVF_FLAT_12, cent, assign, q, 3, 0, "ip")
    assert np.array_equal(ids0, ids) and np.array_equal(dists0, dists)


def test_ivf_flat_reopen_pin(oracle):
    data = np.array([1.0, 0.0, 0.0, 1.0, -1.0, 0.0, 0.0, -1.0], dtype=np.float32).reshape(4, 2)
    cent, assign = oracle.kmeans_train(data, 2, "l2", max_iter=5)
    ids, _ = oracle.ivf_flat_search(data, cent, assign, [1.0, 0.0], 1, 2, "ip")
    assert ids[0] == 0


def test_ivf_flat_full_probe_matches_brute_force(oracle):
    # test_ivf_flat_recall: nprobe = all partitions -> the flat scan's top hit (here: the whole top-5, no ties)
    data = _ivf_flat_lcg_data()
    cent, assign = oracle.kmeans_train(data, 10, "l2", max_iter=10)
    ids, dists = oracle.ivf_flat_search(data, cent, assign, data[0], 5, 10, "ip")
    bf_ids, bf_d = oracle.flat_search(data, data[0], 5, "ip")
    assert ids[0] == bf_ids[0]
    assert np.array_equal(ids, bf_ids.astype(np.uint32))
    assert np.allclose(dists, bf_d, rtol=1e-6)


def test_ivf_flat_ip_routing_shortlist(oracle):
    # inner-product routing on dim >= 64 with >= 64 partitions goes through the 16 highest-variance centroid
    # dimensions (ivf_flat_mmap.rs:312-345, :383-428); other metrics rank every centroid
    rng = np.random.default_rng(5)
    data = rng.random((4000, 64), dtype=np.float32)
    data[:, 7] *= 9.0       # two loud dimensions must be among the routing dimensions
    data[:, 40] *= 7.0
    cent, assign = oracle.kmeans_train(data, 64, "l2", max_iter=5)
    dims = oracle.ivf_flat_routing_dims(cent)
    assert len(dims) == 16 and list(dims) == sorted(dims) and 7 in dims and 40 in dims
    assert len(oracle.ivf_flat_routing_dims(cent[:63])) == 0
    assert len(oracle.ivf_flat_routing_dims(cent[:, :63])) == 0
    q = data[17]
    ids, dists, probes = oracle.ivf_flat_search(data, cent, assign, q, 10, 8, "ip", return_probes=True)
    assert len(probes) == 8 and len(set(probes.tolist())) == 8
    assert np.all(np.diff(dists) <= 0)
    for r, d in zip(ids, dists):
        assert assign[r] in probes
        assert np.float32(oracle.compute_distance(q, data[r], "ip")) == d
    # L2 takes the plain branch: the probed partitions are exactly the 8 nearest centroids
    _, _, probes_l2 = oracle.ivf_flat_search(data, cent, assign, q, 10, 8, "l2", return_probes=True)
    cd = np.array([oracle.compute_distance(q, c, "l2") for c in cent], dtype=np.float32)
    assert set(probes_l2.tolist()) == set(np.argsort(cd, kind="stable")[:8].tolist())


# ---- float16 storage: the scalar f32-query x f16-row kernels (src/distance/simd.rs:805-1092) ------------------------
def test_f16_row_kernels_known_answers(oracle):
    # the f32 known answers (simd.rs:2906-2977) hold on binary16-exact inputs, whichever kernel family computes them
    a, b = [1.0, 2.0, 3.0, 4.0], [4.0, 3.0, 2.0, 1.0]
    assert oracle.compute_distance_f16(a, b, "ip") == 20.0
    assert oracle.compute_distance_f16([1.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0], "l2") == 2.0
    assert oracle.compute_distance_f16(a, a, "cosine") == pytest.approx(0.0, abs=1e-6)
    assert oracle.compute_distance_f16([1.0, 0.0], [0.0, 1.0], "cosine") == 1.0
    assert oracle.compute_distance_f16([0.0, 0.0], [0.0, 1.0], "cosine") == 1.0      # zero norm -> 1 (simd.rs:842-846)
    assert oracle.compute_distance_f16([1.0, 2.0, 3.0], [2.0, 4.0, 1.0], "l1") == 5.0
    assert oracle.compute_distance_f16([1.0, 2.0, 3.0], [2.0, 4.0, 1.0], "chebyshev") == 2.0
    assert oracle.compute_distance_f16([1.0, 0.0, 2.0], [3.0, 0.0, 2.0], "canberra") == 0.5
    assert oracle.compute_distance_f16([1.0, 2.0], [2.0, 1.0], "bray_curtis") == pytest.approx(1.0 / 3.0, rel=1e-6)
    assert oracle.compute_distance_f16([1.0, 0.0], [0.0, 1.0], "jensen_shannon") == pytest.approx(np.sqrt(np.log(2.0)), rel=1e-6)
    assert oracle.compute_distance_f16([1.0, 0.0], [0.0, 1.0], "wasserstein") == 1.0
    assert oracle.compute_distance_f16([1.0, -1.0], [1.0, 1.0], "hellinger") == np.inf


def test_f16_row_kernels_agree_with_the_f32_kernels_to_rounding(oracle):
    rng = np.random.default_rng(11)
    q = rng.random(300, dtype=np.float32)
    row = rng.random(300, dtype=np.float32).astype(np.float16).astype(np.float32)
    for metric in ("ip", "l2", "cosine", "l1", "chebyshev", "canberra", "bray_curtis", "jensen_shannon", "wasserstein",
                   "hellinger", "correlation", "hamming", "jaccard", "dice"):
        assert oracle.compute_distance_f16(q, row, metric) == pytest.approx(oracle.compute_distance(q, row, metric), rel=2e-5, abs=1e-6)
    # and the sequential sum really is a different order from the 8-lane kernels
    assert any(oracle.compute_distance_f16(rng.random(300, dtype=np.float32), row, "ip") != oracle.compute_distance(q, row, "ip")
               for _ in range(4))


def test_f16_collection_reference_cases(oracle):
    # f16_collection_batch_search_reuses_decoded_candidates (src/engine.rs:8043-8076), single-query flavour
    data = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0.5, 0.5, 0, 0]], dtype=np.float32)
    ids, _ = oracle.store_search_f16(data, [1.0, 0.0, 0.0, 0.0], 2, "ip")
    assert ids.tolist() == [0, 2]
    ids, _ = oracle.store_search_f16(data, [0.0, 1.0, 0.0, 0.0], 2, "ip")
    assert ids.tolist() == [1, 2]
