"""GPU parity tests of the row-tile exact scan (lb_scan3.cuh): batches of a dozen queries and more over contiguous f32
rows of up to 512 dims, every 8-lane f32 metric, against the CPU oracle — ids, order and score bits.

The kernel splits the eight AVX lanes of a pair over eight threads, so every shape below is chosen to hit one of its
seams: dims with and without a scalar tail (dim % 8 == 4), each CTA shape (<= 128, <= 256, <= 512 dims), partial query
tiles (nq not a multiple of 16 or 8), a last row block that is not full, several small segments (inner product takes the
two-accumulator kernel there), a row filter, and k larger than a row block."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

METRICS = ["ip", "l2", "cosine", "l1", "chebyshev", "canberra", "bray_curtis"]


@pytest.fixture(scope="module")
def L():
    import lynsedb_b200

    return lynsedb_b200


def _data(n, dim, seed, signed=False):
    rng = np.random.default_rng(seed)
    x = rng.random((n, dim), dtype=np.float32)
    return (x - 0.5).astype(np.float32) if signed else x


def _same(want, got):
    o_ids, o_d, o_c = want
    rows, dists, counts = got
    assert np.array_equal(o_c, counts)
    assert np.array_equal(o_ids.astype(np.uint32), rows), "ids differ from the oracle"
    assert np.array_equal(o_d.view(np.uint32), dists.view(np.uint32)), "scores are not bit-identical to the oracle's"


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("dim,n,nq,k", [(96, 9_001, 40, 10), (100, 8_200, 21, 7), (256, 12_345, 64, 10), (260, 6_000, 13, 3),
                                        (512, 5_000, 33, 20), (8, 70_000, 50, 10)])
def test_tile_scan_matches_the_oracle(L, oracle, metric, dim, n, nq, k):
    corpus, queries = _data(n, dim, 1000 + dim, signed=metric in ("ip", "cosine")), _data(nq, dim, 2000 + dim, signed=metric in ("ip", "cosine"))
    if metric == "canberra":
        corpus[::7, 3] = 0.0     # zero denominators: the term is skipped
        queries[:, 3] = 0.0
    with L.DeviceIndex(dim) as idx:
        idx.set_plan("exact")
        idx.append(corpus)
        got = idx.search(queries, k, metric)
    _same(oracle.store_batch_search(corpus, queries, k, metric, n_threads=3), got)


@pytest.mark.parametrize("metric", ["ip", "l1"])
def test_tile_scan_equals_the_streaming_scan(L, metric, monkeypatch):
    """The same batch with the row-tile kernel switched off goes through lb_scan2.cuh: identical output, bit for bit."""
    n, dim, nq, k = 30_000, 128, 48, 10
    corpus, queries = _data(n, dim, 5), _data(nq, dim, 6)
    outs = []
    for flag in ("1", "0"):
        monkeypatch.setenv("LYNSE_B200_SCAN_TILE", flag)
        with L.DeviceIndex(dim) as idx:
            idx.set_plan("exact")
            idx.append(corpus)
            outs.append(idx.search(queries, k, metric))
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_tile_scan_inner_product_over_small_segments(L, oracle):
    """Segments under 4096 rows score inner products with the two-accumulator kernel (flat_mmap.rs:4845-4869)."""
    dim, k, nq = 72, 9, 30
    parts = [_data(4800, dim, 41), _data(700, dim, 42), _data(5600, dim, 43), _data(33, dim, 44)]
    queries = _data(nq, dim, 45)
    seg = [len(p) for p in parts]
    with L.DeviceIndex(dim) as idx:
        idx.set_plan("exact")
        idx.set_segment_target(1)  # every append opens its own segment
        for p in parts:
            idx.append(p)
        assert idx.segments() == seg
        got_ip = idx.search(queries, k, "ip")
        got_l2 = idx.search(queries, k, "l2")
    corpus = np.concatenate(parts)
    _same(oracle.store_batch_search(corpus, queries, k, "ip", segment_rows=seg, n_threads=1), got_ip)
    _same(oracle.store_batch_search(corpus, queries, k, "l2", segment_rows=seg, n_threads=1), got_l2)


@pytest.mark.parametrize("metric", ["l1", "chebyshev", "l2"])
def test_tile_scan_with_a_row_filter(L, oracle, metric):
    n, dim, k, nq = 20_000, 64, 8, 25
    corpus, queries = _data(n, dim, 51), _data(nq, dim, 52)
    allowed = np.sort(np.random.default_rng(53).choice(n, 9_000, replace=False))
    with L.DeviceIndex(dim) as idx:
        idx.set_plan("exact")
        idx.append(corpus)
        rows, dists, counts = idx.search(queries, k, metric, allow_bits=L.make_allow_bits(n, allowed))
    o_ids, o_d, o_c = oracle.store_batch_search(corpus[allowed], queries, k, metric, n_threads=1)
    assert np.array_equal(counts, o_c)
    assert np.array_equal(rows, allowed[o_ids.astype(np.int64)].astype(np.uint32))
    assert np.array_equal(dists.view(np.uint32), o_d.view(np.uint32))


def test_tile_scan_large_k_and_ties(L, oracle):
    """k above a row block (300 > 128 rows), and duplicated rows: ties resolve by ascending row."""
    n, dim, k, nq = 10_000, 32, 300, 16
    corpus = _data(n, dim, 61)
    corpus[5000:5200] = corpus[100:300]          # exact duplicates
    queries = np.concatenate([corpus[100:108], _data(nq - 8, dim, 62)])
    with L.DeviceIndex(dim) as idx:
        idx.set_plan("exact")
        idx.append(corpus)
        got = idx.search(queries, k, "l1")
    _same(oracle.store_batch_search(corpus, queries, k, "l1", n_threads=1), got)
