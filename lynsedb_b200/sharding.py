"""Row sharding across GPUs and the merge of per-shard results.

The reference fans a query out over shards and merges the per-shard top-k blocks (src/cluster.rs:173-218,
:327-393); inside one node its segments are merged by ``VectorStore::merge_results``
(src/storage/vector_store.rs:953-970): best score first, ties by ascending GLOBAL row.  Here one process drives one
GPU; rows are cut in contiguous ranges so ``global = base + local`` and the order survives.  On the GPU path the
blocks travel through one ``ncclAllGather`` and are merged by ``merge_shards_kernel`` (``lb_sharded_search``); the
functions below are the host-side statement of the same plan and merge (used by bench.py for the plan, by callers
that gather on the host, and by the world_size-2 gloo test).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def shard_range(n_rows: int, world_size: int, rank: int) -> Tuple[int, int]:
    """``(base, count)`` of the contiguous row range of ``rank``: ``[rank * ceil(n / world), ...)`` clipped to n."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad world_size / rank")
    per = (n_rows + world_size - 1) // world_size
    base = min(rank * per, n_rows)
    return base, max(0, min(per, n_rows - base))


def merge_shard_blocks(rows: Sequence[np.ndarray], dists: Sequence[np.ndarray], counts: Sequence[np.ndarray],
                       bases: Sequence[int], k: int, ascending: bool) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Merge per-shard ``[nq][k]`` blocks (local rows, scores, counts) into global ``(rows u64, dists f32, counts u32)``.

    Order per query: score best-first, then global row ascending — ``VectorStore::merge_results``."""
    nq = rows[0].shape[0]
    out_r = np.full((nq, k), np.iinfo(np.uint64).max, dtype=np.uint64)
    out_d = np.full((nq, k), np.nan, dtype=np.float32)
    out_c = np.zeros(nq, dtype=np.uint32)
    for q in range(nq):
        gr: List[np.ndarray] = []
        gd: List[np.ndarray] = []
        for r, d, c, b in zip(rows, dists, counts, bases):
            n = int(c[q])
            gr.append(r[q, :n].astype(np.uint64) + np.uint64(b))
            gd.append(d[q, :n].astype(np.float32))
        allr, alld = np.concatenate(gr), np.concatenate(gd)
        key = alld.astype(np.float64) if ascending else -alld.astype(np.float64)
        order = np.lexsort((allr, key))[:k]
        out_c[q] = len(order)
        out_r[q, :len(order)] = allr[order]
        out_d[q, :len(order)] = alld[order]
    return out_r, out_d, out_c
