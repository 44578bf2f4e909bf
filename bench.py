#!/usr/bin/env python
"""bench.py — FLAT-IP batched search throughput on B200 (BASELINE.json metric).

One "step" = one pass of the hot path over one batch of synthetic queries: the 1024 queries are scored
against the whole corpus and the exact top-10 per query is produced.

  python bench.py --gpus 1 --steps 20 --warmup 5
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus 8 --steps 20 --warmup 5
  python bench.py --impl reference ...      # the reference algorithm's CPU path (C++ oracle) on the host cores

The corpus (default: BASELINE configs[1], 10M x 768 f32) is generated on the device and row-sharded across the
ranks (strong scaling: the corpus is fixed, each rank holds rows/N).  `value` times the path with queries and
results resident in HBM; `e2e` times the same path through the host-buffer C-ABI call (pinned host queries in,
host results out, copies inside the timed region).  torch is used for the multi-process rendezvous only.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (metric, rows, dim, nq, k, description)
    "c2": ("ip", 10_000_000, 768, 1024, 10, "FLAT-IP 10M x 768 f32, batch-1024, k=10 (BASELINE configs[1])"),
    "c1": ("ip", 100_000, 128, 1000, 10, "FLAT-IP 100k x 128 f32, 1k queries, k=10 (BASELINE configs[0])"),
    "c3": ("l2", 10_000_000, 128, 1024, 100, "FLAT-L2 10M x 128 f32, batch-1024, k=100 (BASELINE configs[2])"),
    "c4": ("hamming", 50_000_000, 1024, 4096, 32, "packed Hamming 50M x 1024-bit, batch-4096, k=32 (BASELINE configs[3])"),
    "c4t": ("tanimoto", 50_000_000, 1024, 4096, 32, "packed Tanimoto 50M x 1024-bit, batch-4096, k=32 (BASELINE configs[3])"),
    # weak scaling: 10M rows per GPU, i.e. the full 80M rows at --gpus 8
    "c5": ("ip", 80_000_000, 768, 1024, 10, "FLAT-IP 80M x 768 f32 row-sharded across 8 GPUs (10M rows per GPU), batch-1024, k=10 (BASELINE configs[4])"),
}
PACKED = {"c4", "c4t"}
WEAK = {"c5"}
SEED_CORPUS, SEED_QUERIES = 42, 43
APPEND_ROWS = 100_000  # ingestion batch, as benchmarks/flat_search_bench.py feeds the reference


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    p.add_argument("--rows", type=int, default=None, help="override corpus rows (development only)")
    p.add_argument("--nq", type=int, default=None)
    p.add_argument("--cpu-sample-rows", type=int, default=2_000_000)
    p.add_argument("--cpu-sample-queries", type=int, default=96, help="queries per step of the --impl reference arm")
    p.add_argument("--cpu-baseline-queries", type=int, default=256,
                   help="queries of the cpu_baseline leg of the own arm (about 10 s of CPU work on 16 threads at C2)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--plan", default="auto", choices=["auto", "exact"])
    p.add_argument("--verify-queries", type=int, default=64, help="queries whose id lists are compared with the exact plan")
    p.add_argument("--no-api-e2e", action="store_true", help="skip the Collection.batch_search wall-clock line")
    p.add_argument("--no-c5", action="store_true", help="multi-GPU default run: skip the extra 10M-rows-per-GPU (configs[4]) measurement")
    return p.parse_args()


def measured_peaks():
    """Roofline denominators: MEASURED_PEAKS.json (driver-written) when present, else the profiling guide's fallback.
    The file's key names are matched loosely: an HBM copy bandwidth, a burst and a sustained bf16 GEMM throughput."""
    fallback = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}
    path = ROOT / "MEASURED_PEAKS.json"
    if not path.exists():
        return fallback
    try:
        flat = {}

        def walk(prefix, node):
            if isinstance(node, dict):
                for key, val in node.items():
                    walk(f"{prefix}.{key}".lower(), val)
            elif isinstance(node, (int, float)) and not isinstance(node, bool):
                flat[prefix] = float(node)

        walk("", json.loads(path.read_text()))
        hbm = [v for k, v in flat.items() if "hbm" in k and ("gb" in k or "tb" in k or "bandwidth" in k or "bw" in k)]
        sus = [v for k, v in flat.items() if "bf16" in k and "sustain" in k]
        burst = [v for k, v in flat.items() if "bf16" in k and "sustain" not in k and ("tflop" in k or "tf" in k or "flops" in k)]
        if not hbm or not (sus or burst):
            return fallback
        h = hbm[0] * (1000.0 if hbm[0] < 100 else 1.0)       # TB/s -> GB/s
        b = (burst or sus)[0]
        su = (sus or burst)[0]
        scale = (lambda x: x * 1000.0 if x < 20 else x)       # PFLOP/s -> TFLOP/s
        return {"hbm_gbs": h, "bf16_tflops": scale(b), "bf16_tflops_sustained": scale(su), "source": "measured"}
    except Exception:
        return fallback


def ncu_traffic(args, world, nq):
    """dram__bytes_read.sum + dram__bytes_write.sum (bytes) of the dominant kernel, per launch, from the committed
    `ncu --set full` capture of this same command (profiles/r2_coarse_pair*_ncu.json); None when the run is not a
    captured configuration.  (C4's step is two launches of 2048 queries: the figure is per launch, like `achieved`.)"""
    name = {"c2": "r2_coarse_pair_ncu.json", "c3": "r2_coarse_pair_c3_ncu.json", "c4": "r2_coarse_pair_c4_ncu.json"}.get(args.workload)
    if name is None or args.rows or args.nq or world != 1 or args.plan != "auto":
        return None
    try:
        d = json.loads((ROOT / "profiles" / name).read_text())["kernels"][0]
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        rd, wr = d["dram__bytes_read.sum"], d["dram__bytes_write.sum"]
        return rd["value"] * mult[rd["unit"]] + wr["value"] * mult[wr["unit"]]
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        for line in self.lines:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        loaded = [c for c, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
        return {"sm_mhz": statistics.median(loaded), "sm_max_mhz": max(mx), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


def make_queries(metric: str, nq: int, dim: int, packed: bool = False) -> np.ndarray:
    from lynsedb_b200 import synthetic

    if packed:
        q = synthetic.rows_packed(SEED_QUERIES, np.arange(nq), dim // 64)
        q[0] = synthetic.rows_packed(SEED_CORPUS, np.arange(1), dim // 64)[0]
        return np.ascontiguousarray(q, dtype=np.uint64)
    q = synthetic.rows_f32(SEED_QUERIES, np.arange(nq), dim)
    q[0] = synthetic.rows_f32(SEED_CORPUS, np.arange(1), dim)[0]  # row 0 := query 0 (flat_search_bench.py:76-79)
    return np.ascontiguousarray(q, dtype=np.float32)


# ------------------------------------------------------------------------------------------ reference arm
def cpu_reference_qps(metric, rows_total, dim, k, sample_rows, sample_queries, corpus_sample, queries, repeats=1):
    """The reference algorithm's CPU path (oracle/lynse_oracle.cpp: AVX2+FMA scan, rayon-style chunks, one full
    scan per query as Collection::batch_search does for f32 FLAT, src/engine.rs:5484-5497), all host threads,
    on a row subsample; QPS is extrapolated linearly in the row count."""
    import oracle

    threads = oracle.host_threads()
    nq = min(sample_queries, len(queries))
    seg = []
    left = len(corpus_sample)
    while left > 0:
        seg.append(min(APPEND_ROWS, left))
        left -= seg[-1]
    oracle.store_batch_search(corpus_sample[: min(len(corpus_sample), 20000)], queries[:2], k, metric, n_threads=threads)  # warm
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        oracle.store_batch_search(corpus_sample, queries[:nq], k, metric, segment_rows=seg, n_threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    qps_sample = nq / best
    qps_full = qps_sample * (len(corpus_sample) / rows_total)
    return {"value": qps_full, "unit": "queries/s", "cores": threads, "kind": "port",
            "sample": f"{nq} queries x {len(corpus_sample)} of {rows_total} rows x {dim} dims on {threads} threads, "
                      f"{best:.2f} s; QPS scaled by rows (scan cost is linear in rows)"}


def cpu_reference_packed_qps(metric, rows_total, dim, k, sample_rows, sample_queries, queries):
    """packed_binary_search (flat_mmap.rs:1345-1409) of the oracle on the first rows of the synthetic fingerprints."""
    import oracle
    from lynsedb_b200 import synthetic

    threads = oracle.host_threads()
    words = dim // 64
    data = np.empty((sample_rows, words), dtype=np.uint64)
    for lo in range(0, sample_rows, 500_000):
        hi = min(lo + 500_000, sample_rows)
        data[lo:hi] = synthetic.rows_packed(SEED_CORPUS, np.arange(lo, hi), words)
    nq = min(sample_queries, len(queries))
    oracle.packed_batch_search(data[:20000], queries[:2], k, metric, n_threads=threads)
    t0 = time.perf_counter()
    oracle.packed_batch_search(data, queries[:nq], k, metric, n_threads=threads)
    dt = time.perf_counter() - t0
    return {"value": (nq / dt) * (sample_rows / rows_total), "unit": "queries/s", "cores": threads, "kind": "port",
            "sample": f"{nq} queries x {sample_rows} of {rows_total} fingerprints of {dim} bits on {threads} threads (hardware popcnt), "
                      f"{dt:.2f} s; QPS scaled by rows"}


def run_reference(args, metric, rows, dim, nq, k, desc):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from lynsedb_b200 import synthetic

    sample_rows = min(args.cpu_sample_rows, rows)
    if args.workload in PACKED:
        queries = make_queries(metric, nq, dim, True)
        t0 = time.perf_counter()
        cb = None
        for _ in range(max(args.steps, 1)):
            cb = cpu_reference_packed_qps(metric, rows, dim, k, sample_rows, args.cpu_sample_queries, queries)
        dt = time.perf_counter() - t0
        line = {"impl": "reference", "metric": "queries/sec", "value": cb["value"], "unit": "queries/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt / max(args.steps, 1), "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": {"workload": desc, "rows": rows, "dim": dim, "nq": nq, "k": k, "metric": metric}, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return
    # host copy of the first rows of the same synthetic corpus
    corpus = np.empty((sample_rows, dim), dtype=np.float32)
    step = 50_000
    for lo in range(0, sample_rows, step):
        hi = min(lo + step, sample_rows)
        corpus[lo:hi] = synthetic.rows_f32(SEED_CORPUS, np.arange(lo, hi), dim)
    queries = make_queries(metric, nq, dim)
    sq = max(1, min(args.cpu_sample_queries, nq))
    vals = []
    import oracle

    threads = oracle.host_threads()
    seg = []
    left = sample_rows
    while left > 0:
        seg.append(min(APPEND_ROWS, left))
        left -= seg[-1]
    for _ in range(max(args.warmup, 0)):
        oracle.store_batch_search(corpus, queries[:1], k, metric, segment_rows=seg, n_threads=threads)
    t_total = 0.0
    for s in range(args.steps):
        lo = (s * sq) % max(nq - sq + 1, 1)
        t0 = time.perf_counter()
        oracle.store_batch_search(corpus, queries[lo:lo + sq], k, metric, segment_rows=seg, n_threads=threads)
        t_total += time.perf_counter() - t0
    qps_sample = (sq * args.steps) / t_total
    value = qps_sample * (sample_rows / rows)
    line = {
        "impl": "reference", "metric": "queries/sec", "value": value, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * t_total / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "rows": rows, "dim": dim, "nq": nq, "k": k, "metric": metric},
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": threads, "kind": "port",
                         "sample": f"each step = {sq} queries x {sample_rows} of {rows} rows (sequential scans, "
                                   f"{threads} threads); QPS scaled by rows"},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ our arm
class Rig:
    """One rank's corpus shard + pinned / device buffers + the timed step functions."""

    def __init__(self, args, lib, N, M, comm, rank, world, local_rank, metric, rows, dim, nq, k, packed, dist):
        from lynsedb_b200.index import DeviceIndex
        from lynsedb_b200.sharding import shard_range

        self.lib, self.N, self.comm, self.rank, self.world, self.local_rank, self.dist = lib, N, comm, rank, world, local_rank, dist
        self.metric, self.rows, self.dim, self.nq, self.k, self.packed = metric, rows, dim, nq, k, packed
        self.base, self.n_local = shard_range(rows, world, rank)
        self.idx = DeviceIndex(dim, "packed" if packed else "float32", device=local_rank)
        self.idx.reserve(self.n_local)
        done = 0
        while done < self.n_local:
            m = min(APPEND_ROWS, self.n_local - done)
            self.idx.append_synthetic(m, SEED_CORPUS, self.base + done)
            done += m
        self.idx.set_plan(args.plan)
        self.m_id = M.require(metric)
        self.idx.prepare(self.m_id)
        self.idx.set_timing(True)
        self.queries = make_queries(metric, nq, dim, packed)
        self.qbytes = self.queries.nbytes
        # pinned host staging for the e2e path
        self.hq, self.hrows, self.hdists, self.hcounts = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        N.check(lib.lb_host_malloc(self.qbytes, C.byref(self.hq)))
        N.check(lib.lb_host_malloc(nq * k * 8, C.byref(self.hrows)))
        N.check(lib.lb_host_malloc(nq * k * 4, C.byref(self.hdists)))
        N.check(lib.lb_host_malloc(nq * 4, C.byref(self.hcounts)))
        C.memmove(self.hq, self.queries.ctypes.data, self.qbytes)
        # device-resident buffers for the `value` path
        self.dq, self.drows, self.ddists, self.dcounts = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        N.check(lib.lb_device_malloc(local_rank, self.qbytes, C.byref(self.dq)))
        N.check(lib.lb_device_malloc(local_rank, nq * k * 8, C.byref(self.drows)))
        N.check(lib.lb_device_malloc(local_rank, nq * k * 4, C.byref(self.ddists)))
        N.check(lib.lb_device_malloc(local_rank, nq * 4, C.byref(self.dcounts)))
        N.check(lib.lb_memcpy_h2d(local_rank, self.dq, self.hq, self.qbytes))

    def close(self):
        for p in (self.hq, self.hrows, self.hdists, self.hcounts):
            self.lib.lb_host_free(p)
        for p in (self.dq, self.drows, self.ddists, self.dcounts):
            self.lib.lb_device_free(self.local_rank, p)
        self.idx.close()

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def max_over_ranks(self, x: float) -> float:
        if self.dist is None:
            return x
        import torch

        t = torch.tensor([x], dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x: float) -> float:
        if self.dist is None:
            return x
        import torch

        t = torch.tensor([x], dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def step_device(self):
        self.N.check(self.lib.lb_sharded_search_device(self.comm, self.idx._h, self.m_id, self.dq, self.nq, self.k, self.base,
                                                       self.drows, self.ddists, self.dcounts))

    def step_host(self, nq=None):
        nq = self.nq if nq is None else nq
        if self.packed:
            self.N.check(self.lib.lb_sharded_search_packed(self.comm, self.idx._h, self.m_id, C.cast(self.hq, C.POINTER(C.c_uint64)), nq,
                                                           self.k, self.base, C.cast(self.hrows, C.POINTER(C.c_uint64)),
                                                           C.cast(self.hdists, C.POINTER(C.c_float)), C.cast(self.hcounts, C.POINTER(C.c_uint32))))
            return
        self.N.check(self.lib.lb_sharded_search(self.comm, self.idx._h, self.m_id, C.cast(self.hq, C.POINTER(C.c_float)), nq, self.k,
                                                self.base, C.cast(self.hrows, C.POINTER(C.c_uint64)), C.cast(self.hdists, C.POINTER(C.c_float)),
                                                C.cast(self.hcounts, C.POINTER(C.c_uint32))))

    def host_results(self, nq=None):
        nq = self.nq if nq is None else nq
        grow = np.ctypeslib.as_array(C.cast(self.hrows, C.POINTER(C.c_uint64)), shape=(nq * self.k,)).reshape(nq, self.k).copy()
        gd = np.ctypeslib.as_array(C.cast(self.hdists, C.POINTER(C.c_float)), shape=(nq * self.k,)).reshape(nq, self.k).copy()
        return grow, gd

    def timed(self, fn, steps, slot):
        lib, N, idx = self.lib, self.N, self.idx
        self.barrier()
        N.check(lib.lb_device_synchronize(self.local_rank))
        N.check(lib.lb_index_event_record(idx._h, slot))
        dom, launches, fallbacks = [], 0, 0
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
            st = idx.last_stats()
            dom.append(st["ms_dominant"])
            launches += st["kernels_launched"]
            fallbacks += st["n_fallback"]
        N.check(lib.lb_index_event_record(idx._h, slot + 1))
        N.check(lib.lb_device_synchronize(self.local_rank))
        wall_ms = (time.perf_counter() - t0) * 1000.0
        ms = C.c_float(0)
        N.check(lib.lb_index_event_elapsed_ms(idx._h, slot, slot + 1, C.byref(ms)))
        self.barrier()
        # fallbacks of EVERY rank: one rank's exact-scan re-run holds all of them up at the gather
        return self.max_over_ranks(float(ms.value)), wall_ms, dom, launches, int(self.sum_over_ranks(float(fallbacks))), idx.last_stats()

    def measure(self, steps, warmup):
        for _ in range(max(warmup, 3)):
            self.step_device()
        ms_dev, wall_dev, dom, launches, fallbacks, st = self.timed(self.step_device, steps, 0)
        for _ in range(2):
            self.step_host()
        ms_e2e, wall_e2e, _, _, fb2, _ = self.timed(self.step_host, steps, 2)
        return {"ms_dev": ms_dev, "wall_dev": wall_dev, "dom": dom, "launches": launches, "fallbacks": fallbacks + fb2, "st": st,
                "ms_e2e": ms_e2e, "wall_e2e": wall_e2e}


def verify_results(rig, args, oracle_mod):
    """The last e2e result against (1) the oracle's exact scores of the returned rows, (2) the id lists of the exact
    CUDA-core plan on the same index for a sample of queries (that plan is oracle-verified bit for bit by the GPU
    tests), (3) for corpora the oracle scans in seconds, the oracle's id lists over the full corpus."""
    from lynsedb_b200 import synthetic

    metric, dim, nq, k, packed = rig.metric, rig.dim, rig.nq, rig.k, rig.packed
    grow, gd = rig.host_results()
    out = {}
    # (2) every rank runs the exact plan on the sampled queries (a collective when sharded); rank 0 compares
    ns = min(args.verify_queries, nq)
    sample = np.unique(np.linspace(0, nq - 1, ns).astype(np.int64))
    ns = len(sample)
    sub = np.ascontiguousarray(rig.queries[sample])
    C.memmove(rig.hq, sub.ctypes.data, sub.nbytes)
    rig.idx.set_plan("exact")
    rig.step_host(ns)
    erow, ed = rig.host_results(ns)
    rig.idx.set_plan(args.plan)
    C.memmove(rig.hq, rig.queries.ctypes.data, rig.qbytes)
    if rig.rank != 0:
        return None
    ids_same = bool(np.array_equal(erow, grow[sample]))
    scores_same = bool(np.array_equal(ed.view(np.uint32), gd[sample].view(np.uint32)))
    out["ids_exact_vs_exact_plan"] = ids_same
    out["scores_bit_exact_vs_exact_plan"] = scores_same
    out["exact_plan_queries_compared"] = int(ns)
    # (1) exact scores of the returned rows, order, self-hit
    ok_scores, ok_sorted = True, True
    checked = sorted({0, min(1, nq - 1), nq // 2, nq - 1})
    for qi in checked:
        if packed:
            rws = synthetic.rows_packed(SEED_CORPUS, grow[qi], dim // 64)
            for j in range(k):
                ok_scores &= bool(np.float32(oracle_mod.packed_distance(rig.queries[qi], rws[j], metric)) == gd[qi, j])
            ok_sorted &= bool(np.all(np.diff(gd[qi]) >= 0)) and bool(np.all((np.diff(gd[qi]) > 0) | (np.diff(grow[qi].astype(np.int64)) > 0)))
            continue
        rws = synthetic.rows_f32(SEED_CORPUS, grow[qi], dim)
        for j in range(k):
            if metric == "ip":
                want = oracle_mod.inner_product_batch8_order(rig.queries[qi], rws[j])
            else:
                want = oracle_mod.compute_distance(rig.queries[qi], rws[j], metric)
            ok_scores &= bool(np.float32(want) == gd[qi, j])
        d = gd[qi] if metric != "ip" else -gd[qi]
        ok_sorted &= bool(np.all(np.diff(d) >= 0))
    out.update({"scores_bit_exact_vs_oracle": ok_scores, "sorted": ok_sorted, "queries_checked": len(checked),
                "self_hit_query0_row": int(grow[0, 0])})
    # (3) the whole result against the oracle when the corpus is small enough for the CPU (C1)
    if not packed and rig.world == 1 and rig.rows * dim <= 64_000_000:
        corpus = rig.idx.read_rows(0, rig.rows)
        seg, left = [], rig.rows
        while left > 0:
            seg.append(min(APPEND_ROWS, left))
            left -= seg[-1]
        o_ids, o_d, _ = oracle_mod.store_batch_search(corpus, rig.queries, k, metric, segment_rows=seg, n_threads=oracle_mod.host_threads())
        out["ids_exact_vs_oracle_full_corpus"] = bool(np.array_equal(o_ids.astype(np.uint64), grow))
        out["scores_bit_exact_vs_oracle_full_corpus"] = bool(np.array_equal(o_d.view(np.uint32), gd.view(np.uint32)))
    return out


def roofline_of(rig, res, peaks, args, info):
    st, nq, dim = res["st"], rig.nq, rig.dim
    dom_ms = statistics.mean(res["dom"]) if res["dom"] else float("nan")
    flops_per_launch = 2.0 * nq * rig.n_local * dim  # algorithmic: 2*Q*N*D for this rank's shard
    if st["plan_used"] in (1, 3) and dom_ms > 0:
        ach = flops_per_launch / (dom_ms * 1e-3) / 1e12
        eight = st["coarse_operand"] == 1
        # 8-bit operands: tcgen05 kind::i8 is K = 32 per instruction at the issue cadence of kind::f16's K = 16
        # (profiles/r2_mma_issue_probe.txt), so its ceiling is twice the measured bf16 one
        peak = peaks["bf16_tflops_sustained"] * (2.0 if eight else 1.0)
        return {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TOP/s" if eight else "TFLOP/s", "frac": ach / peak,
                "traffic": ncu_traffic(args, rig.world, nq), "traffic_source": "profiles/r2_coarse_pair[_c3|_c4]_ncu.json (ncu --set full, one launch of this command)",
                "algorithmic_bytes": float(st["algorithmic_bytes"]),
                "kernel": "lb::tc::coarse_pair_kernel" if nq > 128 else "lb::tc::coarse_single_kernel", "kernel_ms": dom_ms,
                "operand": "u8 x u8 -> s32 (tcgen05 kind::i8)" if eight else "bf16 x bf16 -> f32 (tcgen05 kind::f16)",
                "peak_source": peaks["source"] + (" sustained bf16 x 2 (8-bit operands run two K-steps per bf16 K-step)" if eight
                                                  else " (sustained bf16: the kernel is timed inside a long step)"),
                "hbm_gbs_of_kernel": st["algorithmic_bytes"] / (dom_ms * 1e-3) / 1e9,
                "hbm_frac_of_kernel": st["algorithmic_bytes"] / (dom_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}
    if dom_ms > 0:
        bytes_per_launch = float(st["algorithmic_bytes"])
        ach = bytes_per_launch / (dom_ms * 1e-3) / 1e9
        kname = "lb::scan_packed16_kernel" if st["plan_used"] == 2 else "lb::scan_stream_kernel"
        roofline = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                    "traffic": None, "kernel": kname, "kernel_ms": dom_ms, "peak_source": peaks["source"]}
        if st["plan_used"] == 2 and nq > 8:
            # at this batch size the packed scan is bound by the POPC pipe, not by HBM
            pairs = float(nq) * rig.n_local
            popc_peak = cuda_core_peak("popc_b32_per_clk_per_sm", 16.0) * info["sm_count"] * 1.965e9 / (dim // 32)
            roofline.update({"bound": "popc", "achieved": pairs / (dom_ms * 1e-3), "peak": popc_peak, "unit": "pairs/s",
                             "frac": pairs / (dom_ms * 1e-3) / popc_peak,
                             "peak_source": "profiles/r2_cuda_core_peaks.json popc.b32 rate x SMs x 1.965 GHz / (dim/32 popc per pair)"})
        return roofline
    return None


def cuda_core_peak(key, default):
    try:
        return float(json.loads((ROOT / "profiles" / "r2_cuda_core_peaks.json").read_text())[key])
    except Exception:
        return default


def main():
    args = parse_args()
    metric, rows, dim, nq, k, desc = WORKLOADS[args.workload]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    weak = args.workload in WEAK
    if weak:
        rows = rows * world // 8 if args.rows is None else rows  # 10M rows per GPU: the full 80M at 8 GPUs
    if args.rows:
        rows = args.rows
        desc += f" [rows overridden to {rows}]"
    if args.nq:
        nq = args.nq
        desc += f" [nq overridden to {nq}]"
    if args.impl == "reference":
        run_reference(args, metric, rows, dim, nq, k, desc)
        return

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1 (one process per GPU)")
    from lynsedb_b200 import _native as N
    from lynsedb_b200 import metrics as M

    lib = N.lib()
    dist = None
    comm = C.c_void_p()
    if world > 1:
        import torch.distributed as dist  # rendezvous + timing reduction only; the data path is native NCCL

        dist.init_process_group(backend="gloo", init_method="env://")
        ident = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            N.check(lib.lb_nccl_unique_id(ident.ctypes.data_as(C.POINTER(C.c_uint8))))
        import torch

        t = torch.from_numpy(ident)
        dist.broadcast(t, src=0)
        N.check(lib.lb_comm_create(C.byref(comm), local_rank, world, rank, ident.ctypes.data_as(C.POINTER(C.c_uint8))))

    packed = args.workload in PACKED
    rig = Rig(args, lib, N, M, comm, rank, world, local_rank, metric, rows, dim, nq, k, packed, dist)
    info = N.device_info(local_rank)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    res = rig.measure(args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None

    import oracle

    verified = verify_results(rig, args, oracle)
    peaks = measured_peaks()
    value = nq * args.steps / (res["ms_dev"] / 1000.0)
    e2e_value = nq * args.steps / (res["ms_e2e"] / 1000.0)
    st = res["st"]
    roofline = roofline_of(rig, res, peaks, args, info)
    kernel_mhz = float(st.get("coarse_sm_mhz") or 0.0)
    if rank == 0 and roofline and roofline.get("bound") == "tensor" and kernel_mhz > 0:
        # what the tensor pipe can do at the SM clock this run actually held (the B200 power-caps dense 8-bit / bf16 MMAs):
        # 4096 bf16 MACs (8192 8-bit MACs) per clock per SM, i.e. the 64-cycle M128 x N128 MMA of profiles/r2_mma_issue_probe.txt
        macs = 8192.0 if st["coarse_operand"] == 1 else 4096.0
        pipe = 2.0 * macs * info["sm_count"] * kernel_mhz * 1e6 / 1e12
        roofline["kernel_sm_mhz"] = kernel_mhz   # the kernel's own cycle counter over its own wall time (nvidia-smi samples too coarsely)
        roofline["pipe_peak_at_observed_clock"] = pipe
        roofline["frac_of_pipe_at_observed_clock"] = roofline["achieved"] / pipe

    # BASELINE configs[4] beside the headline when the default workload runs on several GPUs: 10M rows per GPU
    # (80M x 768 at 8 GPUs), weak scaling, same queries, same path
    extra = None
    if world > 1 and args.workload == "c2" and not args.rows and not args.no_c5:
        rig.close()
        rig = Rig(args, lib, N, M, comm, rank, world, local_rank, metric, 10_000_000 * world, dim, nq, k, packed, dist)
        r5 = rig.measure(max(3, args.steps // 2), 3)
        steps5 = max(3, args.steps // 2)
        v5 = verify_results(rig, args, oracle)
        extra = {"workload": f"FLAT-IP {10 * world}M x 768 f32 row-sharded across {world} GPUs (10M rows per GPU), batch-1024, k=10 (BASELINE configs[4] at 8 GPUs)",
                 "scaling": "weak", "rows": 10_000_000 * world, "value": nq * steps5 / (r5["ms_dev"] / 1000.0), "unit": "queries/s",
                 "ms_per_step": r5["ms_dev"] / steps5, "e2e_value": nq * steps5 / (r5["ms_e2e"] / 1000.0), "e2e_ms_per_step": r5["ms_e2e"] / steps5,
                 "kernel_ms": statistics.mean(r5["dom"]) if r5["dom"] else None, "fallback_queries": int(r5["fallbacks"]), "verified": v5}

    # the same step through the Python object model a LynseDB user calls: Collection.batch_search over the same index
    api_e2e = None
    if rank == 0 and world == 1 and not packed and not args.no_api_e2e:
        from lynsedb_b200.client import Collection

        coll = Collection("bench", dim, default_index=None)
        coll.attach_store(rig.idx)
        coll.build_index({"ip": "FLAT-IP", "l2": "FLAT-L2", "cosine": "FLAT-COS"}.get(metric, "FLAT-IP"))
        for _ in range(2):
            coll.batch_search(rig.queries, k)
        t0 = time.perf_counter()
        n_api = max(3, args.steps // 2)
        for _ in range(n_api):
            views = coll.batch_search(rig.queries, k)
        api_ms = (time.perf_counter() - t0) * 1000.0 / n_api
        cabi_ms = res["ms_e2e"] / args.steps
        # the same batch through DeviceIndex.search alone (ctypes call, fresh pageable result arrays): what is left of
        # api_ms above it is pure Python (tombstone filter, row -> id, result views)
        store = coll._store
        for _ in range(2):
            store.search(rig.queries, k, metric)
        t0 = time.perf_counter()
        for _ in range(n_api):
            store.search(rig.queries, k, metric)
        native_ms = (time.perf_counter() - t0) * 1000.0 / n_api
        rig.step_host()
        same = bool(np.array_equal(np.stack([v.ids for v in views]).astype(np.uint64), rig.host_results()[0]))
        api_e2e = {"value": nq / (api_ms / 1000.0), "unit": "queries/s", "ms_per_step": api_ms,
                   "call": f"Collection.batch_search({nq} x {dim} float32 ndarray, k={k}) -> list[ResultView] (pageable host arrays)",
                   "above_c_abi_ms": api_ms - cabi_ms, "above_c_abi_frac_of_step": (api_ms - cabi_ms) / api_ms,
                   "device_index_search_ms": native_ms,
                   "python_layer_ms": api_ms - native_ms, "python_layer_frac_of_step": (api_ms - native_ms) / api_ms,
                   "ids_equal_c_abi_result": same}
        coll._store = None   # the rig owns the index

    if rank == 0:
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            sample_rows = min(args.cpu_sample_rows, rig.n_local)
            if packed:
                cpu_baseline = cpu_reference_packed_qps(metric, rows, dim, k, sample_rows, args.cpu_baseline_queries, rig.queries)
            else:
                corpus_sample = rig.idx.read_rows(0, sample_rows)
                cpu_baseline = cpu_reference_qps(metric, rows, dim, k, sample_rows, args.cpu_baseline_queries, corpus_sample, rig.queries)
        eight = st["coarse_operand"] == 1
        if st["plan_used"] == 3:
            dtype = "u8 {0,1} contraction on tcgen05 kind::i8 (s32 accumulate) + exact u64 popcount rescore"
        elif st["plan_used"] == 1:
            dtype = ("u8 coarse contraction (s32 accumulate)" if eight else "bf16 coarse contraction (f32 accumulate)") + " + f32 exact-order rescore"
        else:
            dtype = "u64 popcount" if packed else "f32"
        line = {
            "metric": "queries/sec", "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": res["ms_dev"] / args.steps, "higher_is_better": True,
            "scaling": "weak" if weak else "strong", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": {"workload": desc, "rows": rows, "rows_per_gpu": rig.n_local if extra is None else (rows + world - 1) // world, "dim": dim,
                       "nq": nq, "k": k, "metric": metric, "sharding": f"contiguous row shards x{world}", "plan": args.plan,
                       "l2_policy": "inputs larger than L2 (shadow %.1f GB per GPU vs 126 MB L2)" % (st["algorithmic_bytes"] / 1e9),
                       "device": info["name"], "sm_count": info["sm_count"]},
            "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": int(rig.qbytes),
                    "d2h_bytes_per_step": int(nq * k * 12 + nq * 4), "ms_per_step": res["ms_e2e"] / args.steps},
            "gpu_launches": int(res["launches"]),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "clocks": clocks,
            "fallback_queries": int(res["fallbacks"]),
            "partitions": int(st["n_partitions"]),
            "wall_ms_per_step": res["wall_dev"] / args.steps,
            "verified": verified,
            "api_e2e": api_e2e,
        }
        if extra is not None:
            line["c5_weak"] = extra
        print(json.dumps(line), flush=True)
    rig.close()
    if comm.value:
        lib.lb_comm_destroy(comm)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
