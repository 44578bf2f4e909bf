"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle.

Bar (BASELINE.json north_star): ids and ordering bit-exact; f32 scores bit-exact for the
metrics whose arithmetic is IEEE add/mul/fma/div/sqrt only, and within 1e-5 relative for the
ones with libm transcendentals (Jensen-Shannon tails, Haversine, Hellinger, ...).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

EXACT_BITS = ["ip", "l2", "cosine", "l1", "chebyshev", "canberra", "bray_curtis", "hamming", "jaccard", "tanimoto", "dice"]
TOL_METRICS = ["correlation", "hellinger", "wasserstein", "jensen_shannon"]
REL_TOL = 1e-5


@pytest.fixture(scope="module")
def L():
    import lynsedb_b200

    return lynsedb_b200


def _data(n, dim, seed, positive=True):
    rng = np.random.default_rng(seed)
    x = rng.random((n, dim), dtype=np.float32)
    return x if positive else (x - 0.5).astype(np.float32)


def _check(oracle_out, gpu_out, metric, k):
    o_ids, o_d, o_c = oracle_out
    rows, dists, counts = gpu_out
    assert np.array_equal(o_c, counts)
    for q in range(rows.shape[0]):
        c = int(counts[q])
        assert np.array_equal(o_ids[q, :c].astype(np.uint32), rows[q, :c]), f"ids differ for query {q} ({metric})"
        if metric in EXACT_BITS:
            assert np.array_equal(o_d[q, :c].view(np.uint32), dists[q, :c].view(np.uint32)), \
                f"scores not bit-identical for query {q} ({metric}): {o_d[q, :c]} vs {dists[q, :c]}"
        else:
            np.testing.assert_allclose(dists[q, :c], o_d[q, :c], rtol=REL_TOL, atol=1e-7)
        assert np.all(rows[q, c:] == 0xFFFFFFFF)


# ---- stateless operators -----------------------------------------------------------------------
def test_compute_distance_known_answers(L):
    a, b = np.array([1, 2, 3, 4], np.float32), np.array([4, 3, 2, 1], np.float32)
    assert L.compute_distance(a, b, "ip") == 20.0
    assert L.compute_distance(np.array([1, 0, 0], np.float32), np.array([0, 1, 0], np.float32), "l2") == 2.0
    assert abs(L.compute_distance(a, a, "cosine")) < 1e-6
    assert abs(L.compute_distance(np.array([1, 0], np.float32), np.array([0, 1], np.float32), "cosine") - 1.0) < 1e-6
    assert L.compute_distance(np.array([1, 2, 3], np.float32), np.array([2, 4, 1], np.float32), "l1") == 5.0
    assert L.compute_distance(np.array([1, 2, 3], np.float32), np.array([2, 4, 0], np.float32), "chebyshev") == 3.0
    sh, bj = np.array([121.4737, 31.2304], np.float32), np.array([116.4074, 39.9042], np.float32)
    assert abs(L.compute_distance(sh, bj, "haversine") - 1_067_000) < 10_000
    assert L.compute_distance(sh, sh, "haversine") == 0.0
    with pytest.raises(ValueError, match="Unknown metric"):
        L.compute_distance(a, b, "nope")
    with pytest.raises(ValueError, match="dimensions must match"):
        L.compute_distance(a, b[:3], "ip")
    with pytest.raises(ValueError, match="haversine requires two values"):
        L.compute_distance(a, b, "haversine")


@pytest.mark.parametrize("metric", EXACT_BITS + TOL_METRICS)
@pytest.mark.parametrize("dim", [19, 64])
def test_compute_distance_matches_oracle(L, oracle, metric, dim):
    x = _data(6, dim, 11)
    for i in range(0, 6, 2):
        got = L.compute_distance(x[i], x[i + 1], metric)
        want = oracle.compute_distance(x[i], x[i + 1], metric)
        if metric in EXACT_BITS:
            assert np.float32(got).view(np.uint32) == np.float32(want).view(np.uint32), (metric, got, want)
        else:
            assert abs(got - want) <= REL_TOL * abs(want) + 1e-7


@pytest.mark.parametrize("metric", EXACT_BITS + TOL_METRICS)
def test_top_k_search_matches_oracle(L, oracle, metric):
    cands, q = _data(3000, 19, 5), _data(1, 19, 6)[0]
    ids, d = L.top_k_search(q, cands, metric, 7)
    o_ids, o_d = oracle.top_k_search(q, cands, metric, 7)
    if metric in ("hamming", "jaccard", "tanimoto", "dice"):
        # integer-valued distances tie massively; the reference's quickselect does not order ties by id,
        # so only the distance multiset is defined
        assert np.array_equal(np.sort(d), np.sort(o_d))
        return
    assert np.array_equal(ids, o_ids)
    if metric in EXACT_BITS:
        assert np.array_equal(d.view(np.uint32), o_d.view(np.uint32))
    else:
        np.testing.assert_allclose(d, o_d, rtol=REL_TOL, atol=1e-7)


def test_top_k_search_edges(L):
    cands, q = _data(5, 8, 1), _data(1, 8, 2)[0]
    ids, d = L.top_k_search(q, cands, "l2", 50)
    assert len(ids) == 5 and np.all(np.diff(d) >= 0)
    ids, d = L.top_k_search(q, cands, "ip", 0)
    assert len(ids) == 0 and len(d) == 0
    ids, d = L.top_k_search(q, np.zeros((0, 8), np.float32), "ip", 3)
    assert len(ids) == 0
    with pytest.raises(ValueError, match="Query dimension must match"):
        L.top_k_search(q[:4], cands, "ip", 3)


# ---- flat scan, exact plan -------------------------------------------------------------------------
@pytest.mark.parametrize("metric", EXACT_BITS + TOL_METRICS)
def test_flat_exact_plan_large_segment(L, oracle, metric):
    n, dim, nq, k = 6000, 24, 5, 10  # n >= 4096: chunked path; 6000/3 = 2000 rows per chunk, a multiple of 8
    corpus, queries = _data(n, dim, 21), _data(nq, dim, 22)
    with L.DeviceIndex(dim) as idx:
        idx.set_plan("exact")
        idx.append(corpus)
        got = idx.search(queries, k, metric)
    want = oracle.store_batch_search(corpus, queries, k, metric, n_threads=3)
    _check(want, got, metric, k)


@pytest.mark.parametrize("metric", ["ip", "l2", "cosine", "jensen_shannon", "hamming"])
def test_flat_exact_plan_small_segment(L, oracle, metric):
    n, dim, nq, k = 1000, 37, 4, 12  # n < 4096: sequential path, IP takes the two-accumulator kernel
    corpus, queries = _data(n, dim, 31), _data(nq, dim, 32)
    with L.DeviceIndex(dim) as idx:
        idx.set_plan("exact")
        idx.append(corpus)
        got = idx.search(queries, k, metric)
    want = oracle.store_batch_search(corpus, queries, k, metric, n_threads=4)
    _check(want, got, metric, k)


def test_flat_multi_segment_merge(L, oracle):
    dim, k = 16, 9
    parts = [_data(4800, dim, 41), _data(700, dim, 42), _data(5600, dim, 43)]
    queries = _data(6, dim, 44)
    with L.DeviceIndex(dim) as idx:
        idx.set_plan("exact")
        idx.set_segment_target(1)  # every append opens its own segment
        for p in parts:
            idx.append(p)
        assert idx.segments() == [4800, 700, 5600]
        got_ip = idx.search(queries, k, "ip")
        got_l2 = idx.search(queries, k, "l2")
    corpus = np.concatenate(parts)
    seg = [4800, 700, 5600]
    _check(oracle.store_batch_search(corpus, queries, k, "ip", segment_rows=seg, n_threads=1), got_ip, "ip", k)
    _check(oracle.store_batch_search(corpus, queries, k, "l2", segment_rows=seg, n_threads=1), got_l2, "l2", k)


def test_segment_accounting_follows_the_reference(L):
    with L.DeviceIndex(4) as idx:
        idx.set_segment_target(1024)  # the reference's test-mode target (vector_store.rs:34)
        idx.append(np.zeros((40, 4), np.float32))   # 640 B
        idx.append(np.zeros((20, 4), np.float32))   # 960 B <= 1024: joins
        idx.append(np.zeros((10, 4), np.float32))   # would be 1120 B: new segment
        idx.append(np.zeros((100, 4), np.float32))  # larger than the target on its own: own segment, never split
        assert idx.segments() == [60, 10, 100]
        assert len(idx) == 170


def test_ties_resolve_by_row(L, oracle):
    rng = np.random.default_rng(7)
    base = (rng.random((50, 130)) > 0.5).astype(np.float32)  # dim 130: three words, ragged tail
    corpus = np.tile(base, (120, 1))  # 6000 rows, every row duplicated 120 times
    queries = base[:3]
    for metric in ("hamming", "tanimoto", "dice", "l2"):
        with L.DeviceIndex(130) as idx:
            idx.set_plan("exact")
            idx.append(corpus)
            got = idx.search(queries, 40, metric)
        _check(oracle.store_batch_search(corpus, queries, 40, metric, n_threads=3), got, metric, 40)
        assert np.array_equal(got[0][0, :3], [0, 50, 100])


def test_k_larger_than_n_and_empty(L):
    with L.DeviceIndex(8) as idx:
        q = _data(2, 8, 1)
        rows, dists, counts = idx.search(q, 5, "ip")
        assert np.all(counts == 0) and np.all(rows == 0xFFFFFFFF)
        idx.append(_data(3, 8, 2))
        rows, dists, counts = idx.search(q, 5, "l2")
        assert np.all(counts == 3) and np.all(rows[:, 3:] == 0xFFFFFFFF)
        rows, dists, counts = idx.search(q, 0, "l2")
        assert rows.shape == (2, 0) and np.all(counts == 0)
        with pytest.raises(ValueError, match="Dimension mismatch"):
            idx.search(_data(1, 9, 3), 2, "ip")
        with pytest.raises(ValueError, match="Unknown metric"):
            idx.search(q, 2, "bogus")


def test_filtered_scan(L, oracle):
    n, dim, k = 9000, 20, 8
    corpus, queries = _data(n, dim, 51), _data(4, dim, 52)
    allowed = np.sort(np.random.default_rng(53).choice(n, 4200, replace=False))
    with L.DeviceIndex(dim) as idx:
        idx.append(corpus)
        rows, dists, counts = idx.search(queries, k, "l2", allow_bits=L.make_allow_bits(n, allowed))
    o_ids, o_d, o_c = oracle.store_batch_search(corpus[allowed], queries, k, "l2", n_threads=1)
    assert np.array_equal(counts, o_c)
    assert np.array_equal(rows, allowed[o_ids.astype(np.int64)].astype(np.uint32))
    assert np.array_equal(dists.view(np.uint32), o_d.view(np.uint32))


# ---- packed one-bit rows ---------------------------------------------------------------------------------
@pytest.mark.parametrize("words", [1, 3, 16])
@pytest.mark.parametrize("metric", ["hamming", "tanimoto", "jaccard", "dice"])
def test_packed_index(L, oracle, words, metric):
    rng = np.random.default_rng(61)
    n, nq, k = 7000, 5, 32
    data = rng.integers(0, 2**63, size=(n, words), dtype=np.uint64) * 2 + rng.integers(0, 2, size=(n, words), dtype=np.uint64)
    queries = data[rng.choice(n, nq, replace=False)] ^ rng.integers(0, 2**20, size=(nq, words), dtype=np.uint64)
    with L.DeviceIndex(64 * words, dtype="packed") as idx:
        idx.append(data)
        got = idx.search(queries, k, metric)
    _check(oracle.packed_batch_search(data, queries, k, metric, n_threads=1), got, metric, k)


def test_packed_cache_equals_thresholded_f32(L, oracle):
    # the reference's own pin: packed binary == f32-thresholded (flat_mmap.rs:6385-6421), dim 130
    rng = np.random.default_rng(62)
    corpus = rng.random((5000, 130), dtype=np.float32)
    queries = rng.random((3, 130), dtype=np.float32)
    for metric in ("hamming", "jaccard", "tanimoto", "dice"):
        with L.DeviceIndex(130) as idx:
            idx.append(corpus)
            got = idx.search(queries, 25, metric)
        _check(oracle.store_batch_search(corpus, queries, 25, metric, n_threads=1), got, metric, 25)


# ---- tensor-core plan ------------------------------------------------------------------------------------------
def _bf16_round(x):
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def test_tc_raw_scores_pin_operand_layout(L):
    import ctypes as C
    from lynsedb_b200 import _native as N

    rng = np.random.default_rng(71)
    for (nq, n, dim) in [(5, 200, 64), (130, 333, 200), (128, 1000, 768)]:
        q = (rng.random((nq, dim), dtype=np.float32) - 0.5)
        c = (rng.random((n, dim), dtype=np.float32) - 0.5)
        out = np.zeros((nq, n), dtype=np.float32)
        N.check(N.lib().lb_debug_tc_scores(N.fptr(q), nq, N.fptr(c), n, dim, 0, N.fptr(out)))
        want = _bf16_round(q).astype(np.float64) @ _bf16_round(c).astype(np.float64).T
        err = np.abs(out - want).max()
        assert err < 2e-3, f"tcgen05 scores off by {err} for shape {(nq, n, dim)}"


def _quantise_rows_u8(c):
    """build_shadow_kernel<OPERAND_U8>: one zero point / scale for the corpus, f32 arithmetic."""
    lo, hi = np.float32(c.min()), np.float32(c.max())
    scale = np.float32((hi - lo) / np.float32(255.0)) if hi > lo else np.float32(1.0)
    inv = np.float32(1.0) / scale
    u = np.rint((c - lo).astype(np.float32) * inv)
    return np.clip(u, 0, 255).astype(np.int64)


def _quantise_queries(q):
    """quantise_queries_kernel: u8 (scale max/255) when the batch has no negative element, else s8 (max|.|/127)."""
    signed = bool((q < 0).any())
    levels = np.float32(127.0 if signed else 255.0)
    amax = np.abs(q).max(axis=1).astype(np.float32)
    s = np.where(amax > 0, amax / levels, np.float32(1.0)).astype(np.float32)
    inv = (np.float32(1.0) / s).astype(np.float32)
    u = np.rint(q * inv[:, None])
    return (np.clip(u, -127, 127) if signed else np.clip(u, 0, 255)).astype(np.int64)


@pytest.mark.parametrize("signed_queries", [False, True])
def test_tc_raw_scores_8bit_operands_are_exact_integers(L, signed_queries):
    """kind::i8 operand layouts (A in TMEM, 4 elements per column; B through the same pre-swizzled tiles): the
    accumulators are the integer dot products of the quantised vectors, bit for bit."""
    from lynsedb_b200 import _native as N

    rng = np.random.default_rng(72)
    for (nq, n, dim) in [(5, 200, 64), (130, 333, 200), (128, 1000, 768), (300, 700, 1000)]:
        q = rng.random((nq, dim), dtype=np.float32)
        if signed_queries:
            q -= np.float32(0.5)
        c = rng.random((n, dim), dtype=np.float32) - np.float32(0.25)
        out = np.zeros((nq, n), dtype=np.float32)
        N.check(N.lib().lb_debug_tc_scores(N.fptr(q), nq, N.fptr(c), n, dim, 1, N.fptr(out)))
        want = _quantise_queries(q) @ _quantise_rows_u8(c).T
        # the dump converts the s32 accumulators to f32 (round to nearest): exact below 2^24, the same rounding above
        assert np.array_equal(out, want.astype(np.float32)), f"8-bit accumulators differ for shape {(nq, n, dim)}"


@pytest.mark.parametrize("metric,n,dim,nq,k", [
    ("ip", 20000, 768, 130, 10),
    ("ip", 50001, 128, 64, 10),
    ("l2", 30000, 128, 40, 100),
    ("cosine", 30000, 96, 40, 10),
    ("l2", 12345, 765, 5, 5),
])
def test_tc_plan_matches_oracle(L, oracle, metric, n, dim, nq, k):
    corpus, queries = _data(n, dim, 81), _data(nq, dim, 82)
    with L.DeviceIndex(dim) as idx:
        idx.append(corpus)
        got = idx.search(queries, k, metric)
        st = idx.last_stats()
    assert st["plan_used"] == 1, st
    threads = 1
    if metric == "ip" and n % 8:
        # the n % 8 tail rows of the single chunk take the two-accumulator kernel in the reference; keep them out
        want = oracle.store_batch_search(corpus, queries, k, metric, n_threads=threads)
        rows, dists, counts = got
        assert np.array_equal(want[0].astype(np.uint32), rows)
        np.testing.assert_allclose(dists, want[1], rtol=REL_TOL)
        return
    _check(oracle.store_batch_search(corpus, queries, k, metric, n_threads=threads), got, metric, k)


@pytest.mark.parametrize("dim,nq,k", [(768, 130, 10), (766, 40, 50), (767, 300, 100)])
def test_tc_l2_rows_without_room_for_norm_columns(L, oracle, dim, nq, k):
    """L2 over rows of 766..768 dims: |c|^2 cannot ride in the padded operand row, the epilogue subtracts it as a side
    value (CM_F32_BIAS, cp.async ring); one- and two-CTA kernels, shortlist and hit modes."""
    corpus, queries = _data(40000, dim, 181), _data(nq, dim, 182)
    with L.DeviceIndex(dim) as idx:
        idx.append(corpus)
        got = idx.search(queries, k, "l2")
        st = idx.last_stats()
    assert st["plan_used"] == 1, st
    _check(oracle.store_batch_search(corpus, queries, k, "l2", n_threads=oracle.host_threads()), got, "l2", k)


def test_tc_l2_side_value_mode_on_narrow_rows(L, oracle, monkeypatch):
    monkeypatch.setenv("LYNSE_B200_TC_L2_BIAS", "1")
    corpus, queries = _data(120000, 128, 183), _data(300, 128, 184)
    with L.DeviceIndex(128) as idx:
        idx.append(corpus)
        got = idx.search(queries, 100, "l2")
        st = idx.last_stats()
    assert st["plan_used"] == 1, st
    _check(oracle.store_batch_search(corpus, queries, 100, "l2", n_threads=oracle.host_threads()), got, "l2", 100)


def test_tc_plan_signed_data_and_small_segments(L, oracle):
    dim, k = 64, 10
    parts = [_data(8000, dim, 91, positive=False), _data(1000, dim, 92, positive=False)]
    queries = _data(9, dim, 93, positive=False)
    with L.DeviceIndex(dim) as idx:
        idx.set_segment_target(1)
        for p in parts:
            idx.append(p)
        got = idx.search(queries, k, "ip")
        assert idx.last_stats()["plan_used"] == 1
    _check(oracle.store_batch_search(np.concatenate(parts), queries, k, "ip", segment_rows=[8000, 1000], n_threads=1),
           got, "ip", k)


def test_tc_uncertified_queries_fall_back_to_the_exact_scan(L, oracle):
    # near-duplicate rows: gaps far below the bf16 error bound, so the shortlist cannot be certified
    rng = np.random.default_rng(95)
    base = rng.random((1, 128), dtype=np.float32)
    corpus = np.repeat(base, 8000, axis=0) + rng.random((8000, 128), dtype=np.float32) * 1e-4
    queries = rng.random((5, 128), dtype=np.float32)
    with L.DeviceIndex(128) as idx:
        idx.append(corpus)
        got = idx.search(queries, 10, "ip")
        st = idx.last_stats()
    assert st["plan_used"] == 1 and st["n_fallback"] == 5, st
    _check(oracle.store_batch_search(corpus, queries, 10, "ip", n_threads=1), got, "ip", 10)


# one query tile -> one-CTA kernel, more -> CTA-pair kernel; dims chosen so a tile is 1..12 K blocks, with and
# without a partial last pipeline stage; row counts that leave a ragged last 64-row tile
@pytest.mark.parametrize("nq", [5, 128, 129, 300, 1000])
@pytest.mark.parametrize("n,dim", [(33333, 320), (4100, 64), (20011, 200), (9000, 768), (70001, 130)])
def test_tc_kernels_agree_with_oracle(L, oracle, n, dim, nq):
    k = 10
    corpus, queries = _data(n, dim, 97), _data(nq, dim, 98)
    with L.DeviceIndex(dim) as idx:
        idx.append(corpus)
        got = idx.search(queries, k, "ip")
        assert idx.last_stats()["plan_used"] == 1
    want = oracle.store_batch_search(corpus, queries, k, "ip", n_threads=1)
    assert np.array_equal(want[0].astype(np.uint32), got[0])
    np.testing.assert_allclose(got[1], want[1], rtol=REL_TOL)


def test_tc_shadow_follows_appends(L, oracle):
    # the tiled shadow is a derived structure: it must track rows appended after a search (partial tiles included)
    dim, k = 96, 10
    parts = [_data(5000, dim, 201), _data(37, dim, 202), _data(9000, dim, 203)]
    queries = _data(200, dim, 204)
    with L.DeviceIndex(dim) as idx:
        seen = []
        for p in parts:
            idx.append(p)
            seen.append(p)
            for metric in ("ip", "l2", "cosine"):
                got = idx.search(queries, k, metric)
                assert idx.last_stats()["plan_used"] == 1
                want = oracle.store_batch_search(np.concatenate(seen), queries, k, metric,
                                                 segment_rows=[len(np.concatenate(seen))], n_threads=1)
                assert np.array_equal(want[0].astype(np.uint32), got[0]), metric
                np.testing.assert_allclose(got[1], want[1], rtol=REL_TOL)


def test_tc_and_exact_plans_agree(L):
    n, dim, nq, k = 40000, 256, 33, 20
    corpus, queries = _data(n, dim, 101), _data(nq, dim, 102)
    with L.DeviceIndex(dim) as idx:
        idx.append(corpus)
        a = idx.search(queries, k, "ip")
        idx.set_plan("exact")
        b = idx.search(queries, k, "ip")
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))


# ---- synthetic generator -------------------------------------------------------------------------------------------
def test_device_synthetic_rows_match_host(L):
    from lynsedb_b200 import synthetic

    with L.DeviceIndex(24) as idx:
        idx.append_synthetic(1000, seed=42, row_offset=123456789)
        got = idx.read_rows(990, 10)
    want = synthetic.rows_f32(42, np.arange(123456789 + 990, 123456789 + 1000), 24)
    assert np.array_equal(got, want)


# ---- reference-style FlatIndex ---------------------------------------------------------------------------------------
def test_flat_index_surface(L, oracle, tmp_path):
    data = _data(300, 16, 111)
    idx = L.FlatIndex(str(tmp_path / "vectors.bin"), 16)
    idx.write(data)
    assert len(idx) == 300 and idx.dim == 16
    ids, d = idx.search(data[17], k=5, metric="l2")
    assert ids[0] == 17 and d[0] == 0.0
    res = idx.batch_search(data[:4], k=3, metric="ip")
    o_ids, o_d, _ = oracle.store_batch_search(data, data[:4], 3, "ip", n_threads=1)
    for i, (ids, d) in enumerate(res):
        assert np.array_equal(ids, o_ids[i].astype(np.uint32))
    reopened = L.FlatIndex(str(tmp_path / "vectors.bin"), 16)
    assert len(reopened) == 300
    with pytest.raises(ValueError, match="Unknown metric"):
        idx.search(data[0], 3, "zzz")


# ---- committed golden fixture through the C ABI ---------------------------------------------------------------------
def test_golden_fixture_through_the_c_abi(L):
    import json
    from pathlib import Path

    golden = json.loads((Path(__file__).parent / "golden" / "reference_known_answers.json").read_text())
    for case in golden["compute_distance"]:
        got = L.compute_distance(np.asarray(case["a"], np.float32), np.asarray(case["b"], np.float32), case["metric"])
        assert abs(got - case["expected"]) <= case["abs_tol"], case
    for case in golden["top_k_search"]:
        ids, dists = L.top_k_search(np.asarray(case["query"], np.float32), np.asarray(case["candidates"], np.float32), case["metric"], case["k"])
        assert ids.tolist() == case["ids"], case
        assert np.allclose(dists, case["dists"], atol=1e-6), case


# ---- mid-size corpora: the lockstep slots, the seeded floors of the large-k path and the 128-row tiles ---------------
def _oracle_ids_scores(oracle, corpus, queries, k, metric):
    ids, d, c = oracle.store_batch_search(corpus, queries, k, metric, n_threads=oracle.host_threads())
    assert np.all(c == k)
    return ids.astype(np.uint32), d


def test_tc_lockstep_mid_size_matches_oracle(L, oracle):
    n, dim, nq, k = 300_000, 96, 600, 10          # five query tiles -> three query groups streaming in lockstep
    corpus, queries = _data(n, dim, 301), _data(nq, dim, 302)
    with L.DeviceIndex(dim) as idx:
        for lo in range(0, n, 100_000):
            idx.append(corpus[lo:lo + 100_000])
        rows, dists, counts = idx.search(queries, k, "ip")
        st = idx.last_stats()
    assert st["plan_used"] == 1 and st["n_fallback"] == 0 and np.all(counts == k)
    want_ids, want_d = _oracle_ids_scores(oracle, corpus, queries, k, "ip")
    assert np.array_equal(rows, want_ids)
    assert np.array_equal(dists.view(np.uint32), want_d.view(np.uint32))


@pytest.mark.parametrize("metric", ["l2", "ip"])
def test_tc_large_k_seeded_floors_match_oracle(L, oracle, metric):
    # k > 12 on a corpus long enough for the pre-pass (>= 9216 tiles of 128 rows): seeded floors + 128-row tiles
    n, dim, nq, k = 1_250_000, 64, 200, 50
    corpus, queries = _data(n, dim, 311), _data(nq, dim, 312)
    with L.DeviceIndex(dim) as idx:
        for lo in range(0, n, 250_000):
            idx.append(corpus[lo:lo + 250_000])
        rows, dists, counts = idx.search(queries, k, metric)
        st = idx.last_stats()
    assert st["plan_used"] == 1 and np.all(counts == k), st
    want_ids, want_d = _oracle_ids_scores(oracle, corpus, queries, k, metric)
    assert np.array_equal(rows, want_ids)
    assert np.array_equal(dists.view(np.uint32), want_d.view(np.uint32))


def test_packed_two_million_rows_matches_oracle(L, oracle):
    from lynsedb_b200 import synthetic

    n, nq, k = 2_000_000, 40, 32
    q = synthetic.rows_packed(43, np.arange(nq), 16)
    with L.DeviceIndex(1024, "packed") as idx:
        idx.append_synthetic(n, 42, 0)
        for metric in ("hamming", "tanimoto"):
            rows, dists, counts = idx.search(q, k, metric)
            data = synthetic.rows_packed(42, np.arange(n), 16)
            o_ids, o_d, o_c = oracle.packed_batch_search(data, q, k, metric, n_threads=oracle.host_threads())
            assert np.array_equal(rows, o_ids.astype(np.uint32)) and np.array_equal(dists.view(np.uint32), o_d.view(np.uint32))


def test_concurrent_searches_on_one_index_are_serialised(L):
    # ctypes releases the GIL; the reference holds a read lock per collection (src/python/mod.rs:1187, :1399).  Every
    # entry point takes the index mutex, and the per-call scoring rule (flat / pairwise / f16 rows) lives under it.
    import threading

    rng = np.random.default_rng(123)
    data = rng.random((20000, 64), dtype=np.float32).astype(np.float16).astype(np.float32)
    queries = rng.random((8, 64), dtype=np.float32)
    idx = L.DeviceIndex(64)
    idx.append(data)
    jobs = [("ip", {}), ("ip", {"pairwise": True}), ("ip", {"f16_rows": True}), ("l2", {}), ("cosine", {"f16_rows": True}),
            ("l1", {}), ("hamming", {}), ("jensen_shannon", {})]
    want = [idx.search(queries, 10, m, **kw) for m, kw in jobs]
    got = [[None] * len(jobs) for _ in range(4)]
    errors = []

    def worker(t):
        try:
            for rep in range(3):
                for j in range(len(jobs)):
                    jj = (j + t) % len(jobs)
                    m, kw = jobs[jj]
                    got[t][jj] = idx.search(queries, 10, m, **kw)
        except Exception as e:  # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors
    for t in range(4):
        for j in range(len(jobs)):
            for a, b in zip(want[j], got[t][j]):
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (t, jobs[j])
    idx.close()


@pytest.mark.parametrize("metric", ["ip", "l2", "cosine"])
def test_tc_plan_adversarial_ties(L, oracle, metric):
    # every row duplicated 40 times: the coarse pass sees 40-way exact ties at every rank, and the result must still be
    # the (score, row) order of VectorStore::merge_results (src/storage/vector_store.rs:953-970)
    rng = np.random.default_rng(17)
    base = rng.random((256, 64), dtype=np.float32)
    corpus = np.ascontiguousarray(np.tile(base, (40, 1)))          # 10240 rows
    queries = np.ascontiguousarray(base[:160] + np.float32(0.001))  # two query tiles -> the CTA-pair kernel too
    with L.DeviceIndex(64) as idx:
        idx.append(corpus)
        got = idx.search(queries, 50, metric)
        plan = idx.last_stats()["plan_used"]
    assert plan == 1
    _check(oracle.store_batch_search(corpus, queries, 50, metric), got, metric, 50)


@pytest.mark.parametrize("metric", ["hamming", "tanimoto", "dice"])
def test_packed_1024_bit_adversarial_ties(L, oracle, metric):
    # the 1024-bit TMA kernel on a corpus with 8192-way duplicates (adversarial-ties set)
    rng = np.random.default_rng(23)
    base = rng.integers(0, 2 ** 63, size=(8, 16), dtype=np.uint64)
    corpus = np.ascontiguousarray(np.tile(base, (8192, 1)))         # 65536 rows
    queries = np.ascontiguousarray(base[:5] ^ np.uint64(1))
    with L.DeviceIndex(1024, "packed") as idx:
        idx.append(corpus)
        rows, dists, counts = idx.search(queries, 32, metric)
    o_ids, o_d, o_c = oracle.packed_batch_search(corpus, queries, 32, metric)
    assert np.array_equal(counts, o_c)
    assert np.array_equal(rows, o_ids.astype(np.uint32))
    assert np.array_equal(dists.view(np.uint32), o_d.view(np.uint32))
    assert np.array_equal(rows[0], np.arange(32, dtype=np.uint32) * 8)   # the 32 lowest copies of row 0


def test_k_limits_and_ragged_dimensions(L, oracle):
    corpus, queries = _data(5000, 7, 91), _data(3, 7, 92)
    with L.DeviceIndex(7) as idx:
        idx.append(corpus)
        got = idx.search(queries, 2048, "l2")                        # MAX_K
        _check(oracle.store_batch_search(corpus, queries, 2048, "l2"), got, "l2", 2048)
        with pytest.raises(RuntimeError, match="k above 2048"):
            idx.search(queries, 2049, "l2")
    for dim in (1, 3, 9, 130, 771):                                  # 771: wider than the tensor path takes
        c, q = _data(4500, dim, dim), _data(2, dim, dim + 1)
        with L.DeviceIndex(dim) as idx:
            idx.append(c)
            for metric in ("ip", "cosine", "l1"):
                _check(oracle.store_batch_search(c, q, 5, metric), idx.search(q, 5, metric), metric, 5)


def test_few_queries_on_a_small_corpus_take_the_exact_scan(L, oracle):
    # plan selection only: up to 4 queries over < 256 MiB of rows is a latency case (two launches instead of three)
    corpus, queries = _data(30000, 64, 301), _data(5, 64, 302)
    with L.DeviceIndex(64) as idx:
        idx.append(corpus)
        for nq, plan in ((1, 0), (4, 0), (5, 1)):
            got = idx.search(queries[:nq], 10, "ip")
            assert idx.last_stats()["plan_used"] == plan
            _check(oracle.store_batch_search(corpus, queries[:nq], 10, "ip"), got, "ip", 10)


@pytest.mark.parametrize("metric", ["jensen_shannon", "ip", "wasserstein", "hamming"])
def test_pairwise_search_scores_every_pair_with_compute_distance(L, oracle, metric):
    # search_range / pending_search semantics (src/engine.rs:6410-6483, :3310-3360): compute_distance_f32 per row — for
    # Jensen-Shannon that is the direct kernel, not the cached entropy form of the FLAT scan
    corpus, queries = _data(3000, 24, 401), _data(3, 24, 402)
    with L.DeviceIndex(24) as idx:
        idx.append(corpus)
        rows, dists, counts = idx.search(queries, 7, metric, pairwise=True)
    for qi in range(3):
        score = np.array([oracle.compute_distance(queries[qi], r, metric) for r in corpus], dtype=np.float32)
        order = np.lexsort((np.arange(len(corpus)), -score if metric == "ip" else score))[:7]
        assert counts[qi] == 7
        assert np.array_equal(rows[qi], order.astype(np.uint32)), metric
        if metric in EXACT_BITS:
            assert np.array_equal(dists[qi].view(np.uint32), score[order].view(np.uint32))
        else:
            np.testing.assert_allclose(dists[qi], score[order], rtol=REL_TOL, atol=1e-7)


@pytest.mark.parametrize("metric,n,dim,nq,k,density", [
    ("ip", 40000, 96, 40, 10, 0.5),       # one query tile: coarse_single_kernel
    ("l2", 60000, 64, 300, 10, 0.3),      # CTA pairs
    ("cosine", 50000, 128, 200, 10, 0.9),
    ("l2", 300000, 64, 200, 50, 0.5),     # seeded floors, two epilogue sets
    ("ip", 30000, 200, 150, 10, 0.05),    # sparse filter: 1500 allowed rows
])
def test_tc_plan_with_a_row_filter_matches_oracle(L, oracle, metric, n, dim, nq, k, density):
    # search(where=...) / filter_ids on a batch: the filter masks the hit bits of the tensor-core pass, so the result is
    # the exact top-k of the allowed rows (VectorStore::search_filtered, src/storage/vector_store.rs:1006-1039)
    rng = np.random.default_rng(n + nq)
    corpus, queries = _data(n, dim, 501), _data(nq, dim, 502)
    allowed = np.sort(rng.choice(n, int(n * density), replace=False))
    with L.DeviceIndex(dim) as idx:
        idx.append(corpus)
        rows, dists, counts = idx.search(queries, k, metric, allow_bits=L.make_allow_bits(n, allowed))
        st = idx.last_stats()
    assert st["plan_used"] == 1, st
    o_ids, o_d, o_c = oracle.store_batch_search(np.ascontiguousarray(corpus[allowed]), queries, k, metric, n_threads=1)
    assert np.array_equal(counts, o_c)
    assert np.array_equal(rows, allowed[o_ids.astype(np.int64)].astype(np.uint32))
    if metric == "ip":
        # which rows take the batch-8 or the single-row IP kernel depends on the segment they sit in; the filtered
        # oracle run sees a different segmentation, so compare to rounding (ids above are exact)
        np.testing.assert_allclose(dists, o_d, rtol=REL_TOL)
    else:
        assert np.array_equal(dists.view(np.uint32), o_d.view(np.uint32))
