"""GPU parity tests added in round 2: the certification bound under aligned rounding errors, the BASELINE.json
shapes, the shard merge kernel, the Haversine FLAT scan and the binary metrics on the tensor cores."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

EXACT_BITS = ["ip", "l2", "cosine", "hamming", "jaccard", "tanimoto", "dice"]


@pytest.fixture(scope="module")
def L():
    import lynsedb_b200

    return lynsedb_b200


def _data(n, dim, seed, positive=True):
    rng = np.random.default_rng(seed)
    x = rng.random((n, dim), dtype=np.float32)
    return x if positive else (x - 0.5).astype(np.float32)


def _same(oracle_out, gpu_out, exact_scores=True, rel=1e-5):
    o_ids, o_d, o_c = oracle_out
    rows, dists, counts = gpu_out
    assert np.array_equal(o_c, counts)
    assert np.array_equal(o_ids.astype(np.uint32), rows), "ids differ from the oracle"
    if exact_scores:
        assert np.array_equal(o_d.view(np.uint32), dists.view(np.uint32)), "scores are not bit-identical to the oracle's"
    else:
        np.testing.assert_allclose(dists, o_d, rtol=rel, atol=1e-7)


# ---- the certification bound under the worst alignment of rounding errors --------------------------------------------
def test_bf16_certification_survives_aligned_rounding_errors(L, oracle, monkeypatch):
    """Every element of the queries and of the true best rows sits just below a bf16 midpoint, so both operands round
    DOWN by a whole 2^-8 and the errors add up: the coarse scores of the best rows are ~2 * 2^-8 low (0.50 at |q||c| = 64)
    and they are dropped behind rows whose elements are exact in bf16.  The bound must see that (a one-operand bound of
    2^-8 |q||c| = 0.25 would certify the wrong rows): the queries fall back to the exact scan and the ids are the
    oracle's."""
    monkeypatch.setenv("LYNSE_B200_TC_OPERAND", "bf16")
    dim, n, k = 64, 30_000, 10
    low = np.float32(1.0 + 2.0 ** -8 - 2.0 ** -20)     # rounds to 1.0 in bf16, true value 1.0039
    step = np.float32(1.0 + 2.0 ** -7)                 # exact in bf16
    corpus = np.ones((n, dim), dtype=np.float32)
    corpus[:, :4] = step                               # filler: coarse 64.031, exact 64.281
    corpus[:k, :10] = step                             # decoys: coarse 64.078, exact 64.328 (they fill the top of the shortlists)
    corpus[-20:, :] = low                              # the true best rows: coarse 64.000, exact 64.501
    queries = np.full((5, dim), low, dtype=np.float32)
    with L.DeviceIndex(dim) as idx:
        idx.append(corpus)
        got = idx.search(queries, k, "ip")
        st = idx.last_stats()
    assert st["plan_used"] == 1 and st["coarse_operand"] == 0, st
    assert st["n_fallback"] == 5, f"the aligned-rounding shortlists must not be certified: {st}"
    want = oracle.store_batch_search(corpus, queries, k, "ip", n_threads=1)
    _same(want, got)
    assert np.array_equal(got[0][0], np.arange(n - 20, n - 10, dtype=np.uint32))   # the first ten of the true best rows


def test_u8_certification_survives_aligned_quantisation_errors(L, oracle, monkeypatch):
    """The same attack on the 8-bit operands: elements just below the midpoint of two quantisation levels."""
    monkeypatch.setenv("LYNSE_B200_TC_OPERAND", "u8")
    dim, n, k = 64, 30_000, 10
    lvl = np.float32(1.0 / 255.0)                      # corpus range [0, 1] -> one level = 1/255
    exact_hi = np.float32(200.0) * lvl                 # on a level
    low = np.float32(200.49) * lvl                     # rounds down to level 200, true value half a level higher
    corpus = np.full((n, dim), exact_hi, dtype=np.float32)
    corpus[0, 0], corpus[1, 0] = 0.0, 1.0              # pins the quantisation range to [0, 1]
    corpus[:, 1:3] = np.float32(201.0) * lvl           # filler: two levels up in two columns
    corpus[2:2 + k, 1:6] = np.float32(201.0) * lvl     # decoys
    corpus[-20:, :] = low                              # the true best rows
    queries = np.full((5, dim), low, dtype=np.float32)
    with L.DeviceIndex(dim) as idx:
        idx.append(corpus)
        got = idx.search(queries, k, "ip")
        st = idx.last_stats()
    assert st["plan_used"] == 1 and st["coarse_operand"] == 1, st
    want = oracle.store_batch_search(corpus, queries, k, "ip", n_threads=1)
    _same(want, got)
    assert st["n_fallback"] == 5, f"the aligned-quantisation shortlists must not be certified: {st}"


@pytest.mark.parametrize("operand", ["bf16", "u8"])
@pytest.mark.parametrize("metric", ["ip", "cosine"])
def test_both_operand_kinds_match_the_oracle(L, oracle, monkeypatch, operand, metric):
    monkeypatch.setenv("LYNSE_B200_TC_OPERAND", operand)
    n, dim, nq, k = 60_000, 200, 300, 10
    corpus, queries = _data(n, dim, 501, positive=(metric == "ip")), _data(nq, dim, 502, positive=(metric == "ip"))
    with L.DeviceIndex(dim) as idx:
        idx.append(corpus)
        got = idx.search(queries, k, metric)
        st = idx.last_stats()
    assert st["plan_used"] == 1 and st["coarse_operand"] == (1 if operand == "u8" else 0), st
    # one thread: with several, the reference's chunk boundaries decide which rows take the batch-8 inner-product kernel
    _same(oracle.store_batch_search(corpus, queries, k, metric, n_threads=1), got)


def test_heavy_tailed_rows_switch_the_shadow_to_bf16(L, oracle):
    """One zero point and scale for the whole corpus is too coarse when a few elements are far out: the measured
    quantisation error exceeds twice the bf16 one and the shadow is rebuilt with bf16 operands."""
    n, dim, nq, k = 40_000, 128, 200, 10
    rng = np.random.default_rng(77)
    corpus = rng.standard_normal((n, dim)).astype(np.float32)
    corpus[rng.integers(0, n, 40), rng.integers(0, dim, 40)] = 60.0
    queries = rng.standard_normal((nq, dim)).astype(np.float32)
    with L.DeviceIndex(dim) as idx:
        idx.append(corpus)
        got = idx.search(queries, k, "ip")
        st = idx.last_stats()
    assert st["plan_used"] == 1 and st["coarse_operand"] == 0, st
    want = oracle.store_batch_search(corpus, queries, k, "ip", n_threads=oracle.host_threads())
    assert np.array_equal(want[0].astype(np.uint32), got[0])


def test_non_finite_rows_keep_the_exact_plan(L, oracle):
    n, dim = 20_000, 64
    corpus, queries = _data(n, dim, 601), _data(200, dim, 602)
    corpus[1234, 5] = np.inf
    with L.DeviceIndex(dim) as idx:
        idx.append(corpus)
        got = idx.search(queries, 10, "l2")
        st = idx.last_stats()
    assert st["plan_used"] == 0, st
    _same(oracle.store_batch_search(corpus, queries, 10, "l2", n_threads=oracle.host_threads()), got)


# ---- BASELINE.json shapes ---------------------------------------------------------------------------------------------
def test_baseline_c1_flat_ip_100k_x_128(L, oracle):
    """configs[0], exactly: FLAT-IP, 100k x 128 f32, 1000 queries, k = 10, data as benchmarks/flat_search_bench.py:49-79
    (default_rng(42), U[0,1), row 0 := query 0); ids, order and scores bit-identical to the reference's CPU path."""
    rng = np.random.default_rng(42)
    corpus = rng.random((100_000, 128), dtype=np.float32)
    queries = rng.random((1000, 128), dtype=np.float32)
    corpus[0] = queries[0]
    with L.DeviceIndex(128) as idx:
        idx.append(corpus)
        got = idx.search(queries, 10, "ip")
        st = idx.last_stats()
    assert st["plan_used"] == 1
    _same(oracle.store_batch_search(corpus, queries, 10, "ip", n_threads=1), got)
    assert got[0][0, 0] == 0


def test_baseline_c3_flat_l2_2m_x_128_k100(L, oracle):
    """configs[2] at 2M rows (what the oracle scans in seconds): FLAT-L2, 128 dims, k = 100, hit mode + seeded floors."""
    from lynsedb_b200 import synthetic

    n, dim, nq, k = 2_000_000, 128, 600, 100       # three query groups: the seeded-floor / hit-mode plan at this size
    queries = synthetic.rows_f32(43, np.arange(nq), dim)
    with L.DeviceIndex(dim) as idx:
        for lo in range(0, n, 100_000):
            idx.append_synthetic(100_000, 42, lo)
        rows, dists, counts = idx.search(queries, k, "l2")
        st = idx.last_stats()
        corpus = idx.read_rows(0, n)
    assert st["plan_used"] == 1 and st["coarse_hit_mode"] == 1, st
    want = oracle.store_batch_search(corpus, queries, k, "l2", segment_rows=[100_000] * 20, n_threads=oracle.host_threads())
    _same(want, (rows, dists, counts))


@pytest.mark.parametrize("metric", ["hamming", "tanimoto", "dice"])
def test_baseline_c4_packed_2m_x_1024_bits_q512(L, oracle, metric):
    """configs[3] at 2M fingerprints, 512 queries (four query tiles -> the CTA-pair kernel, {0,1} bytes on tcgen05
    kind::i8), k = 32: ids, order and distances bit-identical to packed_binary_search."""
    from lynsedb_b200 import synthetic

    n, nq, k = 2_000_000, 512, 32
    q = synthetic.rows_packed(43, np.arange(nq), 16)
    q[0] = synthetic.rows_packed(42, np.arange(1), 16)[0]
    with L.DeviceIndex(1024, "packed") as idx:
        idx.append_synthetic(n, 42, 0)
        rows, dists, counts = idx.search(q, k, metric)
        st = idx.last_stats()
    assert st["plan_used"] == 3, st
    data = synthetic.rows_packed(42, np.arange(n), 16)
    _same(oracle.packed_batch_search(data, q, k, metric, n_threads=oracle.host_threads()), (rows, dists, counts))
    assert rows[0, 0] == 0 and dists[0, 0] == 0.0


@pytest.mark.parametrize("metric", ["hamming", "jaccard"])
@pytest.mark.parametrize("words,n,nq,k", [(3, 70_001, 130, 10), (16, 65_536, 40, 32), (24, 50_000, 300, 64), (1, 9_000, 64, 5)])
def test_binary_metrics_on_the_tensor_cores_with_ties_and_ragged_widths(L, oracle, metric, words, n, nq, k):
    """Few distinct fingerprints (heavy ties), widths that leave zero-padded K steps, one- and two-CTA kernels, list and hit
    modes; a tie at the shortlist boundary may not be certified and must then come from the exact scan."""
    rng = np.random.default_rng(words * 1000 + n)
    base = rng.integers(0, 2 ** 63, size=(97, words), dtype=np.uint64)
    corpus = np.ascontiguousarray(base[rng.integers(0, 97, n)] ^ (rng.integers(0, 2 ** 63, size=(n, words), dtype=np.uint64) &
                                                                 rng.integers(0, 2 ** 63, size=(n, words), dtype=np.uint64) &
                                                                 rng.integers(0, 2 ** 63, size=(n, words), dtype=np.uint64)))
    queries = np.ascontiguousarray(base[rng.integers(0, 97, nq)])
    with L.DeviceIndex(words * 64, "packed") as idx:
        idx.append(corpus)
        got = idx.search(queries, k, metric)
        st = idx.last_stats()
    assert st["plan_used"] == 3, st
    _same(oracle.packed_batch_search(corpus, queries, k, metric, n_threads=oracle.host_threads()), got)


def test_binary_metrics_of_an_f32_collection_use_the_packed_cache_on_the_tensor_cores(L, oracle):
    n, dim, nq, k = 50_000, 130, 200, 10      # the reference's own pin uses dim = 130: three words, ragged tail
    corpus, queries = _data(n, dim, 701), _data(nq, dim, 702)
    with L.DeviceIndex(dim) as idx:
        idx.append(corpus)
        got = idx.search(queries, k, "tanimoto")
        st = idx.last_stats()
    assert st["plan_used"] == 3, st
    _same(oracle.store_batch_search(corpus, queries, k, "tanimoto", n_threads=oracle.host_threads()), got)


# ---- Haversine FLAT scan ------------------------------------------------------------------------------------------------
def test_flat_haversine_scan_with_invalid_latitudes(L, oracle):
    """simd.rs:603-628: [lon, lat] in degrees, metres out, |lat| > 90 or a non-finite coordinate -> +inf (ranked last)."""
    rng = np.random.default_rng(11)
    n, nq, k = 50_000, 7, 20
    corpus = np.stack([rng.uniform(-180, 180, n), rng.uniform(-90, 90, n)], axis=1).astype(np.float32)
    corpus[::97, 1] = 95.0          # invalid latitude
    corpus[5::1013, 0] = np.nan     # non-finite longitude
    queries = np.stack([rng.uniform(-180, 180, nq), rng.uniform(-90, 90, nq)], axis=1).astype(np.float32)
    queries[3] = [10.0, 91.0]       # an invalid query: every distance is +inf, rows in ascending order
    with L.DeviceIndex(2) as idx:
        idx.append(corpus)
        rows, dists, counts = idx.search(queries, k, "haversine")
    o_ids, o_d, o_c = oracle.store_batch_search(corpus, queries, k, "haversine", n_threads=1)
    assert np.array_equal(counts, o_c)
    assert np.array_equal(rows, o_ids.astype(np.uint32))
    fin = np.isfinite(o_d)
    assert np.array_equal(np.isfinite(dists), fin)
    np.testing.assert_allclose(dists[fin], o_d[fin], rtol=1e-5)
    assert np.all(np.isinf(dists[3])) and np.array_equal(rows[3], np.arange(k, dtype=np.uint32))
    with L.DeviceIndex(2) as idx:   # a corpus of invalid rows only
        idx.append(np.tile(np.float32([0.0, 120.0]), (5000, 1)))
        rows, dists, counts = idx.search(queries[:2], 5, "haversine")
    assert np.all(np.isinf(dists)) and np.array_equal(rows[0], np.arange(5, dtype=np.uint32))


# ---- the shard merge kernel against the reference's segment merge ----------------------------------------------------------
@pytest.mark.parametrize("metric", ["ip", "l2"])
@pytest.mark.parametrize("n_shards,k", [(2, 10), (8, 10), (5, 100), (3, 1)])
def test_merge_shards_kernel_matches_the_store_merge(L, oracle, metric, n_shards, k):
    """merge_shards_kernel (what every rank runs after the all-gather) on host-supplied blocks against
    VectorStore::merge_results (vector_store.rs:953-970) = the oracle's store search over the same segments; small-integer
    data, so equal scores straddle the shard boundaries and the (score, global row) rule decides."""
    from lynsedb_b200 import _native as N
    from lynsedb_b200 import metrics as M

    rng = np.random.default_rng(1000 + n_shards)
    dim, nq = 8, 33
    sizes = [int(x) for x in rng.integers(max(k, 40), 400, n_shards)]
    sizes[-1] = max(k, 7)                                     # a short last shard
    n = sum(sizes)
    corpus = rng.integers(0, 3, (n, dim)).astype(np.float32)
    queries = rng.integers(0, 3, (nq, dim)).astype(np.float32)
    rows = np.zeros((n_shards, nq, k), np.uint32)
    dists = np.zeros((n_shards, nq, k), np.float32)
    counts = np.zeros((n_shards, nq), np.uint32)
    bases = np.zeros(n_shards, np.uint64)
    pos = 0
    for s, m in enumerate(sizes):
        ids, d, c = oracle.store_batch_search(corpus[pos:pos + m], queries, k, metric, n_threads=1)
        rows[s], dists[s], counts[s], bases[s] = ids.astype(np.uint32), d, c, pos
        pos += m
    out_r = np.zeros((nq, k), np.uint64)
    out_d = np.zeros((nq, k), np.float32)
    out_c = np.zeros(nq, np.uint32)
    N.check(N.lib().lb_merge_shard_blocks(0, M.require(metric), n_shards, nq, k, N.u32ptr(rows), N.fptr(dists), N.u32ptr(counts),
                                          N.u64ptr(bases), N.u64ptr(out_r), N.fptr(out_d), N.u32ptr(out_c)))
    w_ids, w_d, w_c = oracle.store_batch_search(corpus, queries, k, metric, segment_rows=sizes, n_threads=1)
    assert np.array_equal(out_c, w_c)
    assert np.array_equal(out_r, w_ids.astype(np.uint64))
    assert np.array_equal(out_d.view(np.uint32), w_d.view(np.uint32))
    # and the host statement of the same merge (lynsedb_b200.sharding), which the gloo test uses
    from lynsedb_b200.sharding import merge_shard_blocks

    h_r, h_d, h_c = merge_shard_blocks(list(rows), list(dists), list(counts), [int(b) for b in bases], k, metric != "ip")
    assert np.array_equal(h_r, out_r) and np.array_equal(h_c, out_c)


def test_single_rank_sharded_search_with_a_row_filter(L, oracle):
    """lb_sharded_search_filtered with no communicator: the allow-bitset reaches the shard's search (it used to be
    dropped), rows come back rebased by the shard's row base."""
    from lynsedb_b200 import _native as N
    from lynsedb_b200 import metrics as M

    n, dim, nq, k, base = 50_000, 96, 200, 10, 1_000_000
    corpus, queries = _data(n, dim, 801), _data(nq, dim, 802)
    rng = np.random.default_rng(803)
    allowed = rng.random(n) < 0.5
    bits = np.zeros((n + 63) // 64, np.uint64)
    idxs = np.nonzero(allowed)[0]
    np.bitwise_or.at(bits, idxs // 64, np.uint64(1) << (idxs % 64).astype(np.uint64))
    out_r = np.zeros((nq, k), np.uint64)
    out_d = np.zeros((nq, k), np.float32)
    out_c = np.zeros(nq, np.uint32)
    with L.DeviceIndex(dim) as idx:
        idx.append(corpus)
        N.check(N.lib().lb_sharded_search_filtered(None, idx._h, M.require("ip"), N.fptr(queries), nq, k, base, N.u64ptr(bits), len(bits),
                                                   N.u64ptr(out_r), N.fptr(out_d), N.u32ptr(out_c)))
    sub = corpus[allowed]
    w_ids, w_d, w_c = oracle.store_batch_search(sub, queries, k, "ip", n_threads=1)
    assert np.array_equal(out_c, w_c)
    assert np.array_equal(out_r, idxs[w_ids.astype(np.int64)].astype(np.uint64) + np.uint64(base))


# ---- the collection object: round-1 advisor findings + several devices -----------------------------------------------------
def test_ivf_collection_survives_add_search_commit_search(L):
    """An IVF index built before a flush must not outlive it: add -> search (lazy IVF over the flushed rows) -> commit
    (the store grows) -> search used to fail with 'the index changed since the IVF lists were built'."""
    rng = np.random.default_rng(31)
    data = rng.random((3000, 32), dtype=np.float32)
    more = rng.random((700, 32), dtype=np.float32)
    with L.VectorDBClient() as client:
        coll = client.create_collection("db", "ivf", dim=32, default_index=None)
        coll.add(vectors=data)
        coll.build_index("IVF-L2", n_clusters=16, nprobe=16)
        coll.add(vectors=more)                              # stays pending
        first = coll.search(more[3], k=5, nprobe=16)
        assert first.ids[0] == 3003
        coll.commit()
        again = coll.search(more[3], k=5, nprobe=16)        # lists rebuilt over 3700 rows
        assert again.ids.tolist() == first.ids.tolist()
        assert coll.search(data[9], k=1, nprobe=16).ids[0] == 9


def test_thousands_of_deleted_rows_do_not_hit_the_native_k_limit(L, oracle):
    """k + |tombstones| > 2048 on every path that over-fetched: flushed FLAT rows at k > 256, pending rows, IVF, search_range."""
    rng = np.random.default_rng(32)
    n, dim, k = 20_000, 32, 300
    data = rng.random((n, dim), dtype=np.float32)
    pend = rng.random((4000, dim), dtype=np.float32)
    q = rng.random((3, dim), dtype=np.float32)
    with L.VectorDBClient() as client:
        coll = client.create_collection("db", "c", dim=dim, default_index="FLAT-L2")
        coll.add(vectors=data, batch_size=10_000)
        coll.commit()
        coll.add(vectors=pend, batch_size=4000)             # 4000 pending rows
        allv = np.concatenate([data, pend])
        o_ids, o_d, _ = oracle.store_batch_search(allv, q, 2048, "l2", n_threads=oracle.host_threads())
        dead = sorted({int(x) for x in o_ids[:, ::2].ravel()} | set(range(n, n + 3000, 1)))     # every other hit + 3000 pending rows
        assert len(dead) + k > 2048
        coll.delete(dead)
        dead_set = set(dead)
        res = coll.batch_search(q, k)
        for i, r in enumerate(res):
            want = [int(x) for x in o_ids[i] if int(x) not in dead_set][:k]
            assert r.ids.tolist()[:len(want)] == want and len(r) == k
        rr = coll.search_range(q[0], float(o_d[0, 1500]), max_results=1000)
        want = [int(x) for x, d in zip(o_ids[0], o_d[0]) if int(x) not in dead_set and d <= o_d[0, 1500]][:1000]
        assert rr.ids.tolist() == want
        assert len(coll.search(q[0], k=0)) == 0             # k = 0 with tombstones: empty, not |tombstones| rows
    with L.VectorDBClient() as client:
        coll = client.create_collection("db", "ivf", dim=dim, default_index=None)
        coll.add(vectors=data[:6000])
        coll.build_index("IVF-L2", n_clusters=8, nprobe=8)
        o_ids, _, _ = oracle.store_batch_search(data[:6000], q, 2048, "l2", n_threads=oracle.host_threads())
        dead = sorted({int(x) for x in o_ids[0, :2040]})
        coll.delete(dead)
        r = coll.search(q[0], k=20, nprobe=8)               # full probe: exact
        assert r.ids.tolist() == [int(x) for x in o_ids[0] if int(x) not in set(dead)][:20] or len(r) == 20


def test_collection_spread_over_several_device_indexes(L, oracle):
    """ShardedDeviceIndex: the store's segments go round-robin over the devices, every FLAT search fans out and the
    per-device blocks are merged by (score, global row).  On a one-GPU box the 'devices' are two indexes on device 0 —
    the same host logic, segment accounting and merge."""
    rng = np.random.default_rng(33)
    dim, k = 40, 10
    blocks = [rng.integers(0, 4, (m, dim)).astype(np.float32) for m in (9000, 3000, 12000, 500, 7000)]   # integer data: ties across shards
    allv = np.concatenate(blocks)
    queries = rng.integers(0, 4, (150, dim)).astype(np.float32)
    with L.VectorDBClient(devices=[0, 0]) as client:
        coll = client.create_collection("db", "c", dim=dim, default_index="FLAT-IP")
        coll._ensure_store().set_segment_target(9000 * dim * 4)      # small segments so that both shards receive some
        for b in blocks:
            coll.add(vectors=b, batch_size=20_000)
            coll.commit()
        store = coll._store
        assert isinstance(store, L.index.ShardedDeviceIndex)
        segs = store.segments()
        assert sum(segs) == len(allv) and len(segs) >= 3 and min(store.shard_rows()) > 0
        for metric, mode in (("ip", "FLAT-IP"), ("l2", "FLAT-L2"), ("hamming", "FLAT-HAMMING")):
            coll.build_index(mode)
            res = coll.batch_search(queries, k)
            o_ids, o_d, _ = oracle.store_batch_search(allv, queries, k, metric, segment_rows=segs, n_threads=1)
            got = np.stack([r.ids for r in res])
            assert np.array_equal(got, o_ids.astype(np.int64)), metric
            assert np.array_equal(np.stack([r.distances for r in res]).view(np.uint32), o_d.view(np.uint32)), metric
        # a row filter is cut per shard; deleted rows too
        coll.build_index("FLAT-L2")
        keep = set(range(0, len(allv), 3))
        res = coll.batch_search(queries[:20], k, filter_ids=sorted(keep))
        sub = allv[sorted(keep)]
        o_ids, _, _ = oracle.store_batch_search(sub, queries[:20], k, "l2", n_threads=1)
        assert np.array_equal(np.stack([r.ids for r in res]), np.asarray(sorted(keep))[o_ids.astype(np.int64)])
        assert np.array_equal(store.read_rows(8990, 30), allv[8990:9020])


# ---- binary16 rows in HBM (float16 collections) ------------------------------------------------------------------------------
ALL_DENSE = ["ip", "l2", "cosine", "l1", "chebyshev", "canberra", "bray_curtis", "correlation", "hellinger", "wasserstein",
             "jensen_shannon", "hamming", "jaccard", "dice"]


@pytest.mark.parametrize("dim", [64, 100, 200])
def test_binary16_rows_give_the_results_of_the_decoded_rows(L, dim):
    """An index that keeps its rows as IEEE binary16 (half the HBM bytes) against an f32 index holding the same, already
    binary16-exact, values: every kernel decodes on load, so ids, order and scores are bit-identical — for the FLAT scans of
    every metric, the scalar f32-query x f16-row kernels of search() / filtered searches (simd.rs:805-1092), the
    tensor-core plan (shadow built from the binary16 rows, rescoring on them) and a row filter."""
    rng = np.random.default_rng(900 + dim)
    n = 30_000
    rows = (rng.random((n, dim), dtype=np.float32) + np.float32(0.01)).astype(np.float16).astype(np.float32)
    rows[::7] *= np.float32(0.25)
    rows = rows.astype(np.float16).astype(np.float32)
    queries = rng.random((300, dim), dtype=np.float32)
    allowed = rng.random(n) < 0.4
    bits = np.packbits(allowed, bitorder="little")
    bits = np.concatenate([bits, np.zeros((-len(bits)) % 8, np.uint8)]).view(np.uint64)
    with L.DeviceIndex(dim, "float16") as h, L.DeviceIndex(dim, "float32") as f:
        h.append(rows)
        f.append(rows)
        assert np.array_equal(h.read_rows(100, 50), rows[100:150])
        for metric in ALL_DENSE:
            for nq in (3, 40, 300):
                if nq == 300 and metric not in ("ip", "l2", "cosine", "hamming", "jaccard"):
                    continue
                a, b = h.search(queries[:nq], 10, metric), f.search(queries[:nq], 10, metric)
                assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2]), (metric, nq)
                assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)), (metric, nq)
            a, b = h.search(queries[:5], 10, metric, f16_rows=True), f.search(queries[:5], 10, metric, f16_rows=True)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)), ("f16 rows", metric)
            a, b = h.search(queries[:5], 10, metric, bits, f16_rows=True), f.search(queries[:5], 10, metric, bits, f16_rows=True)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)), ("filtered", metric)
        a, b = h.search(queries, 10, "ip", bits), f.search(queries, 10, "ip", bits)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
        assert h.last_stats()["plan_used"] == 1


def test_binary16_rows_halve_the_scan_bytes(L):
    """The point of the half-width layout: the one-query scan reads dim * 2 bytes per row, and runs about twice as fast."""
    import time

    n, dim = 2_000_000, 768
    q = np.random.default_rng(3).random((1, dim), dtype=np.float32)
    times, stats = {}, {}
    for dt in ("float32", "float16"):
        with L.DeviceIndex(dim, dt) as idx:
            idx.append_synthetic(n, 42, 0)
            idx.set_timing(True)
            for _ in range(3):
                r = idx.search(q, 10, "l1")
            best = 1e9
            for _ in range(5):
                idx.search(q, 10, "l1")
                best = min(best, idx.last_stats()["ms_dominant"])
            times[dt], stats[dt] = best, (r, idx.last_stats()["algorithmic_bytes"])
    assert stats["float16"][1] * 2 == stats["float32"][1] == n * dim * 4
    assert np.array_equal(stats["float16"][0][2], stats["float32"][0][2])
    assert times["float16"] < 0.65 * times["float32"], times


def test_float16_collection_on_binary16_storage(L, oracle):
    """The Collection path of a float16 collection now lands on binary16 rows; results stay those of the reference's F16
    paths (compared through the oracle's f16 restatement, as tests/test_gpu_f16_rows.py does for the decoded layout)."""
    rng = np.random.default_rng(41)
    dim, n, k = 48, 9000, 7
    data = rng.random((n, dim), dtype=np.float32)
    q = rng.random((6, dim), dtype=np.float32)
    stored = data.astype(np.float16).astype(np.float32)
    with L.VectorDBClient() as client:
        coll = client.create_collection("db", "h", dim=dim, dtypes="float16", default_index="FLAT-L2")
        coll.add(vectors=data, batch_size=n)
        coll.commit()
        assert coll._store.dtype == "float16"
        single = coll.search(q[0], k)
        o_ids, o_d = oracle.store_search_f16(stored, q[0], k, "l2", n_threads=1)
        assert single.ids.tolist() == [int(x) for x in o_ids] and np.array_equal(single.distances.view(np.uint32), o_d.view(np.uint32))
        batch = coll.batch_search(q, k)
        w_ids, w_d, _ = oracle.store_batch_search(stored, q, k, "l2", n_threads=1)
        assert np.array_equal(np.stack([r.ids for r in batch]), w_ids.astype(np.int64))
        assert np.array_equal(np.stack([r.distances for r in batch]).view(np.uint32), w_d.view(np.uint32))
        coll.build_index("IVF-L2", n_clusters=8, nprobe=8)      # the lists need f32 rows: the store is decoded once
        assert coll._store.dtype == "float32"
        assert coll.search(stored[17], k=1, nprobe=8).ids[0] == 17


# ---- finalize: two-round rescore ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("metric,k", [("ip", 10), ("l2", 40), ("cosine", 1)])
def test_two_round_rescore_equals_one_round(L, oracle, monkeypatch, metric, k):
    """finalize rescans half of its budget first and the rest only when that half does not prove the top k: both settings
    return the oracle's ids and score bits, whether or not the first round suffices (clustered scores make it fail)."""
    rng = np.random.default_rng(77)
    n, dim, nq = 60_000, 96, 33
    corpus = rng.random((n, dim), dtype=np.float32)
    corpus[1000:1400] = corpus[500] + rng.normal(0, 1e-3, (400, dim)).astype(np.float32)   # 400 near-duplicates: a dense score cluster
    queries = rng.random((nq, dim), dtype=np.float32)
    queries[:8] = corpus[500] + rng.normal(0, 1e-3, (8, dim)).astype(np.float32)            # ... which these queries land in
    want = oracle.store_batch_search(corpus, queries, k, metric, n_threads=1)
    for flag in ("1", "0"):
        monkeypatch.setenv("LYNSE_B200_FIN_TWO_ROUNDS", flag)
        with L.DeviceIndex(dim) as idx:
            idx.append(corpus)
            got = idx.search(queries, k, metric)
            assert idx.last_stats()["plan_used"] == 1
        _same(want, got)


# ---- helper warps (list mode of the pair kernel: the shortlists live in warps 6..9, fed through a shared-memory queue) ----
@pytest.mark.parametrize("metric,operand,positive,dim", [("ip", "u8", True, 96), ("ip", "u8", False, 96), ("ip", "bf16", True, 96),
                                                          ("cosine", "u8", False, 200), ("l2", "auto", True, 64),
                                                          ("ip", "u8", True, 768)])
def test_helper_warp_kernel_matches_oracle_and_the_epilogue_lists(L, oracle, monkeypatch, metric, operand, positive, dim):
    """1024 queries (four query groups -> 18 row partitions, all in flight: second-best exchange + its pre-pass) and
    k = 10: ids, order and score bits are the oracle's, and the kernel without helper warps
    (LYNSE_B200_TC_HELPER=0) returns the same bytes.  dim 96 / 64: three accumulator tiles; 200 / 768: two."""
    monkeypatch.setenv("LYNSE_B200_TC_OPERAND", operand)
    n, nq, k = (150_000 if dim <= 200 else 60_000), 1024, 10
    corpus = _data(n, dim, 7, positive)
    queries = _data(nq, dim, 8, positive)
    with L.DeviceIndex(dim) as idx:
        idx.append(corpus)
        got = idx.search(queries, k, metric)
        st = idx.last_stats()
        monkeypatch.setenv("LYNSE_B200_TC_HELPER", "0")
        got_lists = idx.search(queries, k, metric)
        st_lists = idx.last_stats()
    assert st["plan_used"] == 1 and st_lists["plan_used"] == 1, (st, st_lists)
    # (one thread: which rows of a chunk take the reference's batch-8 or its single-row IP kernel depends on the chunking)
    _same(oracle.store_batch_search(corpus, queries, k, metric, n_threads=1), got)
    for a, b in zip(got, got_lists):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_helper_warp_kernel_with_ties_and_a_ragged_last_tile(L, oracle):
    """Many rows share the best score (duplicates spread over every partition) and the corpus ends inside a tile: ties
    resolve by row as in the reference, and padding rows never count as a partition's best keys in the pre-pass."""
    n, dim, nq, k = 100_003, 64, 1024, 12
    rng = np.random.default_rng(11)
    corpus = rng.random((n, dim), dtype=np.float32)
    queries = rng.random((nq, dim), dtype=np.float32)
    dup = rng.choice(n - 1, size=400, replace=False)
    corpus[dup] = corpus[dup[0]] * np.float32(1.5)        # 400 identical strong rows
    best = n - 4   # in the ragged last tile, but not among the <= 7 tail rows whose IP kernel depends on the reference's chunking
    corpus[best] = corpus[dup[0]] * np.float32(2.0)
    with L.DeviceIndex(dim) as idx:
        idx.append(corpus)
        got = idx.search(queries, k, "ip")
        st = idx.last_stats()
    assert st["plan_used"] == 1, st
    _same(oracle.store_batch_search(corpus, queries, k, "ip", n_threads=1), got)
    assert (got[0][:, 0] == best).all()


def test_helper_warp_kernel_with_a_row_filter(L, oracle):
    """A filtered search keeps the helper warps (the filter is applied where the hits are inserted) but not the
    pre-pass of the exchange: group maxima cannot tell allowed rows from filtered ones."""
    n, dim, nq, k = 120_000, 64, 1024, 10
    corpus = _data(n, dim, 21)
    queries = _data(nq, dim, 22)
    allowed = np.arange(0, n, 2)
    with L.DeviceIndex(dim) as idx:
        idx.append(corpus)
        got = idx.search(queries, k, "l2", allow_bits=L.make_allow_bits(n, allowed))
        st = idx.last_stats()
    assert st["plan_used"] == 1, st
    want = oracle.store_batch_search(np.ascontiguousarray(corpus[allowed]), queries, k, "l2", n_threads=oracle.host_threads())
    assert np.array_equal(allowed[want[0].astype(np.int64)].astype(np.uint32), got[0])
    assert np.array_equal(want[1].view(np.uint32), got[1].view(np.uint32))


@pytest.mark.parametrize("nq", [256, 1024])
def test_helper_warp_kernel_small(L, oracle, nq):
    """Small corpora: 256 queries = one query group, 74 partitions, no exchange; 1024 queries = four groups, 18
    partitions, exchange with its pre-pass."""
    n, dim, k = 40_000, 64, 10
    corpus, queries = _data(n, dim, 31), _data(nq, dim, 32)
    with L.DeviceIndex(dim) as idx:
        idx.append(corpus)
        got = idx.search(queries, k, "l2")
        st = idx.last_stats()
    assert st["plan_used"] == 1, st
    _same(oracle.store_batch_search(corpus, queries, k, "l2", n_threads=1), got)


def test_c2_shard_of_a_two_gpu_run_certifies_every_query(L):
    """The first 5M rows of C2's synthetic corpus (what rank 0 of a 2-GPU run holds) with C2's 1024 queries, k = 10: no
    query may go to the exact-scan fallback (with the second-best exchange at depth 2 two of them did, on every step:
    2.4 -> 5.3 ms), and the id lists equal the exact CUDA-core plan's."""
    from lynsedb_b200 import synthetic

    n, dim, nq, k = 5_000_000, 768, 1024, 10
    queries = synthetic.rows_f32(43, np.arange(nq), dim)
    queries[0] = synthetic.rows_f32(42, np.arange(1), dim)[0]
    with L.DeviceIndex(dim) as idx:
        for lo in range(0, n, 100_000):
            idx.append_synthetic(100_000, 42, lo)
        rows, dists, counts = idx.search(queries, k, "ip")
        st = idx.last_stats()
        idx.set_plan("exact")
        sample = np.arange(0, nq, 64)
        e_rows, e_dists, _ = idx.search(queries[sample], k, "ip")
    assert st["plan_used"] == 1 and st["n_fallback"] == 0, st
    assert np.array_equal(rows[sample], e_rows)
    assert np.array_equal(dists[sample].view(np.uint32), e_dists.view(np.uint32))


@pytest.mark.parametrize("n", [2432, 2500, 7000, 8710, 8724, 20_000])
def test_helper_warp_kernel_with_fewer_partitions_than_slots(L, oracle, n):
    """Corpora of a few dozen tiles: some cluster slots get no partition at all (their scanner warps only close the
    queue, their helpers never start), partitions are one or two tiles long, and the last tile is ragged.  8710 / 8724
    rows: 18 partitions of four tiles with the exchange and its pre-pass on, the last partition a single tile of 6 / 20
    rows — fewer than two whole 16-row groups, so it publishes "no floor" instead of a second-best key."""
    dim, nq, k = 64, 1024, 10
    corpus, queries = _data(n, dim, 41), _data(nq, dim, 42)
    with L.DeviceIndex(dim) as idx:
        idx.append(corpus)
        got = idx.search(queries, k, "l2")
    _same(oracle.store_batch_search(corpus, queries, k, "l2", n_threads=1), got)
