#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2h_c2_${N}gpu.json 2> gpurun_out/r2h_c2_${N}gpu.err
echo "== c2 x$N rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2h_c2_${N}gpu.json') if l.startswith('{')][-1])
print(' value %.0f e2e %.0f ms/step %.3f e2e_ms %.3f kernel %.3f fb %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['fallback_queries']))
print(' verified', d['verified'])
print(' c5', {k: d['c5_weak'][k] for k in ('value','e2e_value','ms_per_step','kernel_ms','fallback_queries','verified')})
PY
# a collection spread over the GPUs of one process
LYNSE_B200_DEVICES=$(python -c "print(','.join(str(i) for i in range($N)))") timeout 600 python - <<PY
import numpy as np, time, sys
sys.path.insert(0, '.')
import lynsedb_b200 as L
import oracle
rng = np.random.default_rng(5)
dim, k = 96, 10
data = rng.random((400_000, dim), dtype=np.float32)
q = rng.random((256, dim), dtype=np.float32)
with L.VectorDBClient() as client:
    coll = client.create_collection("db", "c", dim=dim, default_index="FLAT-IP")
    coll._ensure_store().set_segment_target(40_000 * dim * 4)
    for lo in range(0, len(data), 40_000):
        coll.add(vectors=data[lo:lo + 40_000], batch_size=40_000)
    coll.commit()
    store = coll._store
    print(' devices', store.devices, 'shard rows', store.shard_rows(), 'segments', len(store.segments()))
    res = coll.batch_search(q, k)
    o_ids, o_d, _ = oracle.store_batch_search(data, q, k, "ip", segment_rows=store.segments(), n_threads=1)
    print(' sharded collection ids equal oracle:', np.array_equal(np.stack([r.ids for r in res]), o_ids.astype(np.int64)),
          'scores bit-equal:', np.array_equal(np.stack([r.distances for r in res]).view(np.uint32), o_d.view(np.uint32)))
    t0 = time.perf_counter()
    for _ in range(10): coll.batch_search(q, k)
    print(' batch_search 256 q: %.2f ms' % ((time.perf_counter() - t0) * 100))
PY
