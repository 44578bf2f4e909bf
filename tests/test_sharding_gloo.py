"""N > 1 host-side path on CPU: two gloo ranks each own a contiguous row shard, produce their local top-k block,
all-gather the blocks and merge them — the plan and merge rule of lb_sharded_search (one ncclAllGather + merge by
(score, global row)).  The per-shard search itself is a numpy stand-in here; the GPU one is covered by -m gpu tests."""
import os
import socket
import sys

import numpy as np
import pytest

from lynsedb_b200.sharding import merge_shard_blocks, shard_range


def test_shard_ranges_cover_the_corpus():
    for n, w in [(10, 1), (10, 3), (7, 8), (80_000_000, 8), (0, 2)]:
        spans = [shard_range(n, w, r) for r in range(w)]
        assert sum(c for _, c in spans) == n
        pos = 0
        for base, cnt in spans:
            assert base == pos or cnt == 0
            pos += cnt


def _local_topk(corpus, queries, k, ascending):
    nq = queries.shape[0]
    rows = np.full((nq, k), 0xFFFFFFFF, np.uint32)
    dists = np.full((nq, k), np.nan, np.float32)
    counts = np.zeros(nq, np.uint32)
    for q in range(nq):
        s = ((corpus - queries[q]) ** 2).sum(1).astype(np.float32) if ascending else (corpus @ queries[q]).astype(np.float32)
        order = np.lexsort((np.arange(len(s)), s if ascending else -s))[:k]
        counts[q] = len(order)
        rows[q, :len(order)] = order
        dists[q, :len(order)] = s[order]
    return rows, dists, counts


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(123)                      # every rank sees the same corpus and queries
        n, dim, nq, k = 1001, 8, 6, 5
        corpus = rng.integers(0, 4, (n, dim)).astype(np.float32)   # small integers: plenty of exact ties across shards
        queries = rng.integers(0, 4, (nq, dim)).astype(np.float32)
        ok = True
        for ascending in (True, False):
            base, cnt = shard_range(n, world, rank)
            rows, dists, counts = _local_topk(corpus[base:base + cnt], queries, k, ascending)
            block = torch.from_numpy(np.concatenate([rows.astype(np.float64).ravel(), dists.astype(np.float64).ravel(),
                                                     counts.astype(np.float64), [float(base)]]))
            gathered = [torch.empty_like(block) for _ in range(world)]
            dist.all_gather(gathered, block)                  # the one collective of the path
            rs, ds, cs, bs = [], [], [], []
            for g in gathered:
                g = g.numpy()
                rs.append(g[:nq * k].reshape(nq, k).astype(np.uint32))
                ds.append(g[nq * k:2 * nq * k].reshape(nq, k).astype(np.float32))
                cs.append(g[2 * nq * k:2 * nq * k + nq].astype(np.uint32))
                bs.append(int(g[-1]))
            mr, md, mc = merge_shard_blocks(rs, ds, cs, bs, k, ascending)
            wr, wd, wc = _local_topk(corpus, queries, k, ascending)      # single-shard answer
            ok &= bool(np.array_equal(mr, wr.astype(np.uint64)) and np.array_equal(md, wd) and np.array_equal(mc, wc))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_allgather_and_merge_equals_single_shard():
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)]
