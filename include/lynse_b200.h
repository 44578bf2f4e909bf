/* lynse_b200.h — C ABI of liblynse_b200.so, the B200-native replacement for the
 * native side of LynseDB's batched distance + top-k path.
 *
 * The reference has no C FFI: its seam is the pyo3 extension module
 * `lynse._core` (reference src/python/mod.rs:25-42) consumed by
 * python/lynse/_backend.py:28.  Each entry point below names the pyo3
 * function / Rust routine it replaces (paths relative to the reference tree).
 * INTEGRATION.md shows the ctypes stub a maintainer would drop into
 * python/lynse/_backend.py to bind them.
 *
 * Conventions
 *   - plain pointers and sizes only; the caller owns every host buffer;
 *   - every function returns an lb_status (0 = ok) and leaves a thread-local
 *     message retrievable with lb_last_error();
 *   - status -> Python exception mirrors src/error.rs:54-72
 *       LB_INVALID_ARGUMENT, LB_DIMENSION_MISMATCH -> ValueError
 *       LB_IO                                     -> IOError
 *       everything else                           -> RuntimeError
 *   - rows are u32 positions inside one index (one GPU shard); the host shim
 *     rebases them to u64 global rows (src/storage/vector_store.rs:988-998);
 *   - result blocks are [nq][k]; entries past counts[q] hold row 0xFFFFFFFF.
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     point returns LB_CUDA.
 */
#ifndef LYNSE_B200_H
#define LYNSE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum lb_status {
    LB_OK = 0,
    LB_INVALID_ARGUMENT = 1,
    LB_DIMENSION_MISMATCH = 2,
    LB_IO = 3,
    LB_CUDA = 4,
    LB_NCCL = 5,
    LB_UNSUPPORTED = 6,
    LB_INTERNAL = 7
} lb_status;

/* DistanceMetric, same order as src/distance/mod.rs:19-36 */
typedef enum lb_metric {
    LB_IP = 0,
    LB_L2 = 1,
    LB_COSINE = 2,
    LB_HAMMING = 3,
    LB_JACCARD = 4,
    LB_MANHATTAN = 5,
    LB_HAVERSINE = 6,
    LB_CORRELATION = 7,
    LB_HELLINGER = 8,
    LB_WASSERSTEIN = 9,
    LB_DICE = 10,
    LB_TANIMOTO = 11,
    LB_JENSEN_SHANNON = 12,
    LB_CHEBYSHEV = 13,
    LB_CANBERRA = 14,
    LB_BRAY_CURTIS = 15,
    LB_METRIC_COUNT = 16
} lb_metric;

/* Storage dtype of an index. LB_F32 = VectorDtype::F32 (src/storage/dtype.rs).
 * LB_PACKED_U64 holds pre-packed one-bit rows (the reference's lazily built
 * BinaryData cache, src/storage/flat_mmap.rs:126-160, ingested directly so a
 * 50M x 1024-bit corpus need not exist as f32).  LB_F16 = VectorDtype::F16 (src/storage/dtype.rs:60-67): rows are kept as
 * IEEE binary16 in HBM — half the bytes of every scan — and decoded (exactly) on load by every kernel; queries stay f32. */
typedef enum lb_dtype { LB_F32 = 0, LB_PACKED_U64 = 1, LB_F16 = 2 } lb_dtype;

/* Search plan selector for lb_index_set_plan (diagnostics / tests). */
typedef enum lb_plan {
    LB_PLAN_AUTO = 0,   /* tensor-core coarse pass + exact rescore where available */
    LB_PLAN_EXACT = 1   /* CUDA-core exact scan only */
} lb_plan;

typedef struct lb_index lb_index;
typedef struct lb_comm lb_comm;
typedef struct lb_ivf lb_ivf;

/* ---- library ----------------------------------------------------------- */
const char* lb_last_error(void);
const char* lb_version(void);
int lb_device_count(int* out);
/* name[cap], total/free bytes, SM count of a device */
int lb_device_info(int device, char* name, int cap, uint64_t* total_bytes, uint64_t* free_bytes, int* sm_count);

/* ---- stateless operators ---------------------------------------------- */
/* replaces py_compute_distance (src/python/mod.rs:2161-2185) ->
 * distance::compute_distance_f32 (src/distance/mod.rs:193-213). */
int lb_compute_distance(const float* a, const float* b, uint32_t dim, int metric, float* out);

/* replaces py_top_k_search (src/python/mod.rs:2189-2223) ->
 * distance::top_k_search (src/distance/mod.rs:373-422).  candidates is a host
 * [n][dim] row-major matrix.  Writes min(k,n) entries best-first. */
int lb_top_k_search(const float* query, const float* candidates, uint64_t n, uint32_t dim, uint32_t k, int metric,
                    uint32_t* ids, float* dists, uint32_t* out_count);

/* ---- flat index (one GPU shard) ---------------------------------------- */
/* replaces FlatMmap::open + VectorStore segment bookkeeping
 * (src/storage/flat_mmap.rs, src/storage/vector_store.rs:379-445). */
int lb_index_create(lb_index** out, uint32_t dim, int dtype, int device);
void lb_index_destroy(lb_index* idx);
int lb_index_reserve(lb_index* idx, uint64_t n_rows);
/* segment_target_bytes: DEFAULT_SEGMENT_TARGET_BYTES / LYNSE_SEGMENT_TARGET_BYTES
 * (src/storage/vector_store.rs:32, :225-229); 0 = default 256 MiB. */
int lb_index_set_segment_target(lb_index* idx, uint64_t segment_target_bytes);
/* The next append opens a new segment even if it would fit the last one: a host that spreads the segments of one
 * VectorStore over several indexes (one per GPU) keeps the reference's segment boundaries this way (the inner-product
 * kernel choice depends on them, src/storage/flat_mmap.rs:4845-4869). */
int lb_index_new_segment(lb_index* idx);
/* replaces VectorStore::append / FlatMmap::write: host row-major rows -> HBM.
 * An append is never split across segments. */
int lb_index_append_f32(lb_index* idx, const float* rows, uint64_t n);
/* LB_PACKED_U64 indexes: rows of ceil(dim/64) u64 words, bit i of a row at
 * word i/64 bit i%64 (src/distance/simd.rs:750-757). */
int lb_index_append_packed(lb_index* idx, const uint64_t* words, uint64_t n);
/* LB_F16 indexes: raw binary16 rows (the reference's on-disk f16 segments, src/storage/vector_store.rs:24-60);
 * lb_index_append_f32 on an LB_F16 index narrows f32 values with round-to-nearest-even. */
int lb_index_append_f16(lb_index* idx, const uint16_t* rows, uint64_t n);
/* synthetic corpus generated on the device (bench / large parity tests):
 * f32: value(row, col) = u24(hash(seed, row*dim+col)) * 2^-24  in [0,1)
 * packed: word(row, w) = hash64(seed, row*words+w).  `row_offset` is the global
 * row of this shard's first new row, so shards of one corpus agree. */
int lb_index_append_synthetic(lb_index* idx, uint64_t n, uint64_t seed, uint64_t row_offset);
uint64_t lb_index_len(const lb_index* idx);
uint32_t lb_index_dim(const lb_index* idx);
/* segment row counts, oldest first; returns number of segments */
int lb_index_segments(const lb_index* idx, uint64_t* rows_out, int cap, int* n_segments);
/* copy rows [first, first+n) back to the host (tests) */
int lb_index_read_rows_f32(lb_index* idx, uint64_t first, uint64_t n, float* out);

/* Build the per-metric side structures now instead of on first search:
 * bf16 shadow + norms (IP / L2 / cosine tensor-core path), packed bits
 * (binary metrics; FlatMmap::ensure_binary, flat_mmap.rs:388-401), row stats
 * (Jensen-Shannon; ensure_jensen_shannon, flat_mmap.rs:949-983). */
int lb_index_prepare(lb_index* idx, int metric);
int lb_index_set_plan(lb_index* idx, int plan);

/* replaces Collection::search / batch_search down to FlatMmap::search
 * (src/engine.rs:4697-4833, :5352-5498; src/storage/vector_store.rs:972-1039;
 *  src/storage/flat_mmap.rs:824-923).  Host buffers; H2D/D2H inside.
 * queries: [nq][dim] f32 (for LB_PACKED_U64 indexes pass query words through
 * lb_index_search_packed).  allow_bits: optional row filter, bit r = row r
 * allowed, LSB-first u64 words (src/storage/bitset.rs). */
int lb_index_search(lb_index* idx, int metric, const float* queries, uint32_t nq, uint32_t k,
                    const uint64_t* allow_bits, uint64_t allow_words, uint32_t* out_rows, float* out_dists,
                    uint32_t* out_counts);
/* The same search with every pair scored by compute_distance_f32 (src/distance/mod.rs:193-213: the single-row
 * kernels, which for IP means two accumulators) on the exact scan — what Collection::search_range
 * (src/engine.rs:6410-6483) and the pending-rows search (:3310-3360) do. */
int lb_index_search_pairwise(lb_index* idx, int metric, const float* queries, uint32_t nq, uint32_t k,
                             const uint64_t* allow_bits, uint64_t allow_words, uint32_t* out_rows, float* out_dists,
                             uint32_t* out_counts);
/* The same search for a collection whose rows are binary16 values (VectorDtype::F16, held decoded in `idx`):
 * every pair is scored by compute_distance_f16 (src/distance/mod.rs:217-237 -> the scalar f32-query x f16-row
 * kernels, src/distance/simd.rs:805-1092) as FlatMmap::search / search_filtered do on F16 storage
 * (src/storage/flat_mmap.rs:905-907 -> exact_flat_search_f16 :1259-1281; :511-520 -> search_filtered_f16
 * :5329-5437).  The binary metrics use the packed cache, as there. */
int lb_index_search_f16_rows(lb_index* idx, int metric, const float* queries, uint32_t nq, uint32_t k,
                             const uint64_t* allow_bits, uint64_t allow_words, uint32_t* out_rows, float* out_dists,
                             uint32_t* out_counts);
int lb_index_search_packed(lb_index* idx, int metric, const uint64_t* query_words, uint32_t nq, uint32_t k,
                           uint32_t* out_rows, float* out_dists, uint32_t* out_counts);

/* Same search with queries and results resident in HBM (device pointers),
 * enqueued on the index stream; returns after the work is complete. */
int lb_index_search_device(lb_index* idx, int metric, const void* d_queries, uint32_t nq, uint32_t k,
                           uint32_t* d_out_rows, float* d_out_dists, uint32_t* d_out_counts);

/* Statistics of the most recent search on this index. */
typedef struct lb_search_stats {
    uint32_t plan_used;        /* 0 exact scan, 1 tensor coarse + rescore, 2 packed scan, 3 tensor coarse over {0,1} bytes + exact counts */
    uint32_t n_fallback;       /* queries whose shortlist was not certified and were re-run exactly */
    uint32_t n_partitions;     /* row partitions of the dominant kernel */
    uint32_t kernels_launched; /* kernels launched by the search call */
    float ms_dominant;         /* CUDA-event duration of the dominant kernel (when timing is on) */
    float ms_total;            /* CUDA-event duration of all device work of the call */
    uint64_t algorithmic_bytes;/* bytes the dominant kernel must read once (shadow or corpus) */
    uint64_t algorithmic_flops;/* 2*nq*n*dim for the tensor path, else 0 */
    uint32_t coarse_operand;   /* tensor plans: 0 = bf16 operands (f32 accumulators), 1 = 8-bit operands (s32 accumulators) */
    uint32_t coarse_hit_mode;  /* tensor plans: 1 = rows above a seeded floor were appended to per-query hit buffers (large k) */
    float coarse_sm_mhz;       /* tensor plans: SM clock the coarse kernel actually ran at (its own cycle counter over its own wall time) */
    uint32_t reserved;
} lb_search_stats;
int lb_index_set_timing(lb_index* idx, int enabled);
int lb_index_last_stats(const lb_index* idx, lb_search_stats* out);

/* ---- IVF over one index ------------------------------------------------- */
/* replaces IVFIndex::build (src/index/ivf.rs:132-179) = kmeans::train_for_metric (src/index/kmeans.rs:74-139:
 * farthest-point seeding on a sample drawn with the seed-42 LCG, at most max_iter Lloyd steps — the reference
 * passes 20 — assignment under the routing metric, which is L2 for the binary metrics, ivf.rs:80-87) +
 * inverted_lists_from_assignments (kmeans.rs:317-345).  The rows stay in `idx`; rebuild after appends. */
int lb_ivf_train(lb_index* idx, int metric, uint32_t n_clusters, uint32_t max_iter, lb_ivf** out);
/* the same index from given centroids [n_centroids][dim] and per-row assignments [len(idx)] */
int lb_ivf_create(lb_index* idx, int metric, const float* centroids, uint32_t n_centroids, const uint32_t* assignments,
                  lb_ivf** out);
void lb_ivf_destroy(lb_ivf* ivf);
int lb_ivf_info(const lb_ivf* ivf, uint32_t* n_centroids, uint64_t* n_rows);
int lb_ivf_centroids(const lb_ivf* ivf, float* out);         /* [n_centroids][dim] */
int lb_ivf_assignments(const lb_ivf* ivf, uint32_t* out);    /* [n_rows] */
/* replaces IVFIndex::search (src/index/ivf.rs:181-348) for a batch of queries; nprobe == 0 means 1.
 * allow_bits: optional subset filter (SearchParams::subset), same layout as lb_index_search. */
int lb_ivf_search(lb_ivf* ivf, const float* queries, uint32_t nq, uint32_t k, uint32_t nprobe, const uint64_t* allow_bits,
                  uint64_t allow_words, uint32_t* out_rows, float* out_dists, uint32_t* out_counts);
/* replaces IvfFlatMmap::search (src/storage/ivf_flat_mmap.rs:225-300; pyo3 `_core.IvfFlatIndex.search`,
 * src/python/mod.rs:2128-2155) over an index trained with metric L2 (IvfFlatMmap::build -> kmeans::train_l2,
 * ivf_flat_mmap.rs:96-101): partitions are chosen by find_nearest_centroids (:383-446) under the SEARCH metric,
 * with the routing-dimension shortlist for inner product on dim >= 64 and >= 64 partitions; every row of the
 * probed partitions is scored with compute_distance_f32; no corpus fallback.  Rows are build-data positions. */
int lb_ivf_flat_search(lb_ivf* ivf, const float* queries, uint32_t nq, uint32_t k, uint32_t nprobe, int metric,
                       uint32_t* out_rows, float* out_dists, uint32_t* out_counts);

/* ---- device / pinned memory helpers (bench + tests; no torch) ---------- */
int lb_device_malloc(int device, uint64_t bytes, void** out);
int lb_device_free(int device, void* p);
int lb_host_malloc(uint64_t bytes, void** out);   /* pinned */
int lb_host_free(void* p);
int lb_memcpy_h2d(int device, void* dst, const void* src, uint64_t bytes);
int lb_memcpy_d2h(int device, void* dst, const void* src, uint64_t bytes);
int lb_device_synchronize(int device);
/* write `bytes` of device memory (L2 flush helper) */
int lb_device_memset(int device, void* dst, int value, uint64_t bytes);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink ------------------- */
/* replaces the reference's shard fan-out + merge (src/cluster.rs:173-218,
 * :327-393): every rank searches its row shard, one ncclAllGather moves the
 * per-shard [nq][k] blocks, the host merges by (score, global row). */
int lb_nccl_unique_id(uint8_t* id128);
int lb_comm_create(lb_comm** out, int device, int world_size, int rank, const uint8_t* id128);
void lb_comm_destroy(lb_comm* c);
/* gathers `bytes` from every rank's d_send into d_recv[rank*bytes ...] */
int lb_comm_allgather(lb_comm* c, const void* d_send, void* d_recv, uint64_t bytes);
int lb_comm_barrier(lb_comm* c);

/* The multi-GPU hot path in one call (per rank): search this rank's shard, pack the [nq][k]
 * block (u32 local rows, f32 scores, counts, the shard's global row base), ncclAllGather the
 * blocks of all ranks over NVLink on the index stream, merge them on the GPU by (score, global
 * row) — VectorStore::merge_results order (src/storage/vector_store.rs:953-970) — and return
 * global u64 rows.  comm == NULL means a single shard.  Requires k <= rows of every shard. */
int lb_sharded_search(lb_comm* comm, lb_index* idx, int metric, const float* queries, uint32_t nq, uint32_t k,
                      uint64_t row_base, uint64_t* out_rows, float* out_dists, uint32_t* out_counts);
/* the same with a row filter over this rank's shard (allow_bits: bit r = local row r allowed, as lb_index_search):
 * Collection.search(where=...) / tombstoned rows on a sharded collection (src/storage/vector_store.rs:1006-1039). */
int lb_sharded_search_filtered(lb_comm* comm, lb_index* idx, int metric, const float* queries, uint32_t nq, uint32_t k,
                               uint64_t row_base, const uint64_t* allow_bits, uint64_t allow_words, uint64_t* out_rows,
                               float* out_dists, uint32_t* out_counts);
int lb_sharded_search_packed(lb_comm* comm, lb_index* idx, int metric, const uint64_t* query_words, uint32_t nq, uint32_t k,
                             uint64_t row_base, uint64_t* out_rows, float* out_dists, uint32_t* out_counts);
/* The merge step of the sharded search by itself (tests): n_shards blocks of [nq][k] u32 local rows / f32 scores,
 * [nq] counts and one row base per shard, laid out shard-major, merged by (score, global row) on `device` —
 * VectorStore::merge_results (src/storage/vector_store.rs:953-970). */
int lb_merge_shard_blocks(int device, int metric, int n_shards, uint32_t nq, uint32_t k, const uint32_t* rows, const float* dists,
                          const uint32_t* counts, const uint64_t* row_bases, uint64_t* out_rows, float* out_dists,
                          uint32_t* out_counts);
int lb_sharded_search_device(lb_comm* comm, lb_index* idx, int metric, const void* d_queries, uint32_t nq, uint32_t k,
                             uint64_t row_base, uint64_t* d_out_rows, float* d_out_dists, uint32_t* d_out_counts);

/* CUDA events on the index stream (bench timing: device time, not wall clock). slot in [0,8). */
int lb_index_event_record(lb_index* idx, int slot);
int lb_index_event_elapsed_ms(lb_index* idx, int slot_a, int slot_b, float* ms);

/* ---- diagnostics -------------------------------------------------------- */
/* Runs the tensor-core coarse kernel (inner product) on a small problem and dumps the raw coarse keys [nq][n]
 * (the parity tests use it to pin the tcgen05 operand layouts).  operand 0: bf16 operands, keys = f32 accumulators;
 * operand 1: 8-bit operands (rows: u8 with one zero point / scale for the corpus; queries: u8, or s8 when the batch
 * holds a negative element), keys = the integer accumulators sum(qhat * chat), returned as f32. */
int lb_debug_tc_scores(const float* queries, uint32_t nq, const float* rows, uint32_t n, uint32_t dim, int operand, float* out);
#ifdef __cplusplus
}
#endif
#endif /* LYNSE_B200_H */
