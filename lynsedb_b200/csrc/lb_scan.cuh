// lb_scan.cuh — exact CUDA-core scan kernels with fused per-partition top-k.
//
// Replaces the reference's fused scan loops (src/storage/flat_mmap.rs):
//   fused_topk_ip_parallel :4845-4869 / ip_scan_chunk_topk :2179-2256
//   fused_topk_parallel    :4876-4982 / fused_topk_seq     :4985-5044
//   packed_binary_search   :1345-1409
//   jensen_shannon_cached_parallel :985-1111
//   fused_topk_parallel_filtered   :5439-5554 (bitset) / direct_access_topk :5223-5274 (row list)
//   topk_insert :2132-2166, merge_topk_results :5183-5214, VectorStore::merge_results (vector_store.rs:953-970)
//
// Shape: the corpus is cut into P contiguous row partitions (the analogue of
// the reference's rayon chunks and 256 MiB segments), one resident CTA each.
// A CTA walks its partition in blocks of 256 rows (one row per thread); for a
// tile of up to 16 queries every thread evaluates its row against each query
// with the order-exact functions of lb_metrics.cuh and drops a 64-bit ranking
// key into shared memory; then each warp takes queries of the tile and offers
// the 256 keys to that (partition, query)'s running top-k list, gated by the
// list's current worst key exactly like the reference's strict threshold gate.
// Because keys order by (score, row) the result is the reference's top-k no
// matter how rows are partitioned.  A second kernel merges the P lists per
// query with a shared-memory bitonic sort.
#pragma once
#include "lb_metrics.cuh"

#include <type_traits>

namespace lb {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_TQ = 16;  // queries per tile

struct ScanArgs {
    const float* corpus;        // [n][dim] f32 (null for packed scan and for binary16 rows)
    const __half* corpus_h;     // [n][dim] binary16 rows of a float16 index (kernels instantiated with RT = __half)
    const uint64_t* words;      // [n][n_words] packed rows (packed scan)
    uint32_t n_rows;            // rows scanned (or length of row_ids)
    int dim;
    int n_words;
    const float* queries;       // [nq][dim]  (Jensen-Shannon cached: mass-normalised queries)
    const uint64_t* qwords;     // [nq][n_words] (packed scan)
    int nq;
    int k;
    int metric;
    const uint64_t* allow_bits; // optional row filter, bit r = row r allowed
    const uint32_t* row_ids;    // optional explicit row list (ascending) instead of 0..n_rows
    const uint32_t* small_seg;  // optional [2*n_small] (start,end) ranges of rows living in segments < 4096 rows
    int n_small;
    int ip_single;              // 1: IP uses the two-accumulator kernel for every row (compute_distance semantics)
    int f16_rows;               // 1: every pair goes through compute_distance_f16's scalar kernels (float16 collections)
    const float* row_stats;     // Jensen-Shannon cached: [n][2] (inv_mass, entropy) or null
    const float* query_stats;   // Jensen-Shannon cached: [nq][2]
    const double* row_mass;     // Wasserstein streaming scan: [n] f64 row sums (NaN = invalid value in the row) or null
    uint64_t* lists;            // [P][nq][k] keys, unsorted
    uint32_t* counts;           // [P][nq]
    uint64_t* thr;              // [P][nq] current worst kept key (valid when count == k)
    uint32_t rows_per_part;     // multiple of SCAN_THREADS
    const float* query_tiles;   // row-tile scan (lb_scan3.cuh): the query tiles in the kernel's shared-memory layout
    int smem_lists;             // 1: a batch of one query tile keeps its top-k lists in shared memory (room reserved by the host)
};

__device__ __forceinline__ uint64_t warp_max_u64(uint64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        uint64_t other = __shfl_xor_sync(0xffffffffu, v, o);
        v = other > v ? other : v;
    }
    return v;
}

// One warp offers 256 keys (8 per lane) of one query to the (partition, query) list in global memory.
__device__ __forceinline__ void warp_offer_keys(const uint64_t* __restrict__ skeys /*[256] smem*/, uint64_t* list,
                                                uint32_t* count_p, uint64_t* thr_p, int k, int lane) {
    uint32_t cnt = *count_p;
    uint64_t thr = cnt < (uint32_t)k ? KEY_NONE : *thr_p;
    bool dirty = false;
    for (int i = 0; i < SCAN_THREADS / 32; ++i) {
        uint64_t key = skeys[i * 32 + lane];
        unsigned m = __ballot_sync(0xffffffffu, key < thr);
        while (m) {
            int src = __ffs(m) - 1;
            m &= m - 1;
            uint64_t kk = __shfl_sync(0xffffffffu, key, src);
            if (kk >= thr) continue;  // threshold tightened by an earlier insert of this group
            dirty = true;
            if (cnt < (uint32_t)k) {
                if (lane == 0) list[cnt] = kk;
                ++cnt;
                __syncwarp();
                if (cnt == (uint32_t)k) {
                    uint64_t mx = 0;
                    for (int idx = lane; idx < k; idx += 32) {
                        uint64_t v = __ldcg(list + idx);
                        mx = v > mx ? v : mx;
                    }
                    thr = warp_max_u64(mx);
                }
            } else {
                // replace the current worst (keys are unique, so exactly one slot equals thr)
                for (int idx = lane; idx < k; idx += 32)
                    if (__ldcg(list + idx) == thr) list[idx] = kk;
                __syncwarp();
                uint64_t mx = 0;
                for (int idx = lane; idx < k; idx += 32) {
                    uint64_t v = __ldcg(list + idx);
                    mx = v > mx ? v : mx;
                }
                thr = warp_max_u64(mx);
            }
        }
    }
    if (dirty && lane == 0) {
        *count_p = cnt;
        *thr_p = thr;
    }
}

template <class RT>
__device__ __forceinline__ const RT* corpus_rows(const ScanArgs& a) {
    if constexpr (std::is_same<RT, float>::value) return a.corpus;
    else return a.corpus_h;
}

__device__ __forceinline__ bool row_allowed(const uint64_t* __restrict__ allow_bits, uint32_t row) {
    return allow_bits == nullptr || ((__ldg(allow_bits + (row >> 6)) >> (row & 63)) & 1ull);
}
__device__ __forceinline__ bool in_small_segment(const uint32_t* __restrict__ small_seg, int n_small, uint32_t row) {
    for (int i = 0; i < n_small; ++i)
        if (row >= small_seg[2 * i] && row < small_seg[2 * i + 1]) return true;
    return false;
}

// Ranking value of one (query, row) pair for the FLAT scan (flat_mmap.rs:1173-1230 dispatch).
template <bool ASC, class CP>
__device__ __forceinline__ float flat_pair_value(const ScanArgs& a, const float* __restrict__ q /*smem*/, int qi,
                                                 CP c, uint32_t row, bool vec, bool small) {
    if (a.f16_rows) return compute_distance_f16order<false>(a.metric, q, c, a.dim, vec);  // flat_mmap.rs:1259-1281
    if (!ASC) {
        // IP: rows of segments >= 4096 rows take the batch-8 kernel, smaller segments the single-row kernel
        // (flat_mmap.rs:4845-4869; the <=7 tail rows of a rayon chunk are thread-count dependent)
        return (a.ip_single || small) ? ip_single_order<false>(q, c, a.dim, vec) : ip_batch8_order<false>(q, c, a.dim, vec);
    }
    if (a.metric == LB_JENSEN_SHANNON && a.row_stats != nullptr) {
        float q_inv = a.query_stats[2 * qi], q_ent = a.query_stats[2 * qi + 1];
        float r_inv = __ldg(a.row_stats + 2 * (size_t)row), r_ent = __ldg(a.row_stats + 2 * (size_t)row + 1);
        if (q_inv == 0.0f) {  // zero-mass query (flat_mmap.rs:938-972)
            if (r_inv != r_inv || !isfinite(r_ent)) return INFINITY;
            return r_inv == 0.0f ? 0.0f : kLn2;  // ranked on the squared distance
        }
        return jensen_shannon_precomputed_divergence<false>(q, c, a.dim, vec, q_ent, r_inv, r_ent);
    }
    return compute_distance<false>(a.metric, q, c, a.dim, vec);
}

template <bool ASC, class RT = float>
__global__ void __launch_bounds__(SCAN_THREADS) scan_exact_kernel(ScanArgs a) {
    extern __shared__ __align__(16) unsigned char smem_scan[];
    uint64_t* skeys = reinterpret_cast<uint64_t*>(smem_scan);                           // [SCAN_TQ][256]
    float* sq = reinterpret_cast<float*>(smem_scan + SCAN_TQ * SCAN_THREADS * 8);      // [SCAN_TQ][dim_pad]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int part = blockIdx.x;
    const int dim = a.dim;
    const int dim_pad = (dim + 3) & ~3;
    const bool vec = (dim & 3) == 0;
    const uint64_t part_begin = (uint64_t)part * a.rows_per_part;
    uint64_t part_end = part_begin + a.rows_per_part;
    if (part_end > a.n_rows) part_end = a.n_rows;

    for (uint64_t blk = part_begin; blk < part_end; blk += SCAN_THREADS) {
        uint64_t slot = blk + tid;
        bool valid = slot < part_end;
        uint32_t row = 0;
        if (valid) row = a.row_ids ? __ldg(a.row_ids + slot) : (uint32_t)slot;
        if (valid && !row_allowed(a.allow_bits, row)) valid = false;
        const RT* c = corpus_rows<RT>(a) + (size_t)row * dim;
        bool small = valid && a.n_small > 0 && in_small_segment(a.small_seg, a.n_small, row);
        for (int q0 = 0; q0 < a.nq; q0 += SCAN_TQ) {
            int tq = min(SCAN_TQ, a.nq - q0);
            __syncthreads();  // previous tile's keys and queries are no longer read
            for (int i = tid; i < tq * dim; i += SCAN_THREADS) {
                int qq = i / dim, d = i - qq * dim;
                sq[qq * dim_pad + d] = __ldg(a.queries + (size_t)(q0 + qq) * dim + d);
            }
            __syncthreads();
            for (int j = 0; j < tq; ++j) {
                uint64_t key = KEY_NONE;
                if (valid) {
                    float v = flat_pair_value<ASC>(a, sq + j * dim_pad, q0 + j, c, row, vec, small);
                    key = make_key<ASC>(v, row);
                }
                skeys[j * SCAN_THREADS + tid] = key;
            }
            __syncthreads();
            for (int j = warp; j < tq; j += SCAN_THREADS / 32) {
                size_t lq = (size_t)part * a.nq + (q0 + j);
                warp_offer_keys(skeys + j * SCAN_THREADS, a.lists + lq * a.k, a.counts + lq, a.thr + lq, a.k, lane);
            }
        }
    }
}

// ---- packed one-bit rows: Hamming / Jaccard / Tanimoto / Dice on u64 words -----------------------------
// (flat_mmap.rs:1345-1409 + simd.rs:765-801).  Row words stay in registers across the query tile.
template <int W>  // W = words per row held in registers (0 = generic, re-read)
__global__ void __launch_bounds__(SCAN_THREADS) scan_packed_kernel(ScanArgs a) {
    extern __shared__ __align__(16) unsigned char smem_packed[];
    uint64_t* skeys = reinterpret_cast<uint64_t*>(smem_packed);                           // [SCAN_TQ][256]
    uint64_t* sq = reinterpret_cast<uint64_t*>(smem_packed + SCAN_TQ * SCAN_THREADS * 8); // [SCAN_TQ][n_words]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int part = blockIdx.x;
    const int nw = W > 0 ? W : a.n_words;
    const uint64_t part_begin = (uint64_t)part * a.rows_per_part;
    uint64_t part_end = part_begin + a.rows_per_part;
    if (part_end > a.n_rows) part_end = a.n_rows;
    const int metric = a.metric;

    for (uint64_t blk = part_begin; blk < part_end; blk += SCAN_THREADS) {
        uint64_t slot = blk + tid;
        bool valid = slot < part_end;
        uint32_t row = 0;
        if (valid) row = a.row_ids ? __ldg(a.row_ids + slot) : (uint32_t)slot;
        if (valid && !row_allowed(a.allow_bits, row)) valid = false;
        const uint64_t* rw = a.words + (size_t)row * nw;
        uint64_t r[W > 0 ? W : 1];
        uint32_t row_pop = 0;
        if (W > 0 && valid) {
            if (W >= 2) {
#pragma unroll
                for (int w = 0; w < W; w += 2) {
                    ulonglong2 t = __ldg(reinterpret_cast<const ulonglong2*>(rw + w));
                    r[w] = t.x;
                    r[w + 1] = t.y;
                }
            } else {
                r[0] = __ldg(rw);
            }
#pragma unroll
            for (int w = 0; w < W; ++w) row_pop += __popcll(r[w]);
        }
        for (int q0 = 0; q0 < a.nq; q0 += SCAN_TQ) {
            int tq = min(SCAN_TQ, a.nq - q0);
            __syncthreads();
            for (int i = tid; i < tq * nw; i += SCAN_THREADS) sq[i] = __ldg(a.qwords + (size_t)q0 * nw + i);
            __syncthreads();
            for (int j = 0; j < tq; ++j) {
                uint64_t key = KEY_NONE;
                if (valid) {
                    const uint64_t* qw = sq + j * nw;
                    uint32_t x = 0, y = 0;
                    if (metric == LB_HAMMING) {
                        if (W > 0) {
#pragma unroll
                            for (int w = 0; w < W; ++w) x += __popcll(r[w] ^ qw[w]);
                        } else {
                            for (int w = 0; w < nw; ++w) x += __popcll(__ldg(rw + w) ^ qw[w]);
                        }
                    } else if (metric == LB_DICE) {
                        uint32_t qpop = 0;
                        if (W > 0) {
#pragma unroll
                            for (int w = 0; w < W; ++w) {
                                x += __popcll(r[w] & qw[w]);
                                qpop += __popcll(qw[w]);
                            }
                            y = qpop + row_pop;
                        } else {
                            for (int w = 0; w < nw; ++w) {
                                uint64_t rv = __ldg(rw + w);
                                x += __popcll(rv & qw[w]);
                                y += __popcll(rv) + __popcll(qw[w]);
                            }
                        }
                    } else {
                        if (W > 0) {
#pragma unroll
                            for (int w = 0; w < W; ++w) {
                                x += __popcll(r[w] & qw[w]);
                                y += __popcll(r[w] | qw[w]);
                            }
                        } else {
                            for (int w = 0; w < nw; ++w) {
                                uint64_t rv = __ldg(rw + w);
                                x += __popcll(rv & qw[w]);
                                y += __popcll(rv | qw[w]);
                            }
                        }
                    }
                    key = make_key<true>(packed_finish(metric, x, y), row);
                }
                skeys[j * SCAN_THREADS + tid] = key;
            }
            __syncthreads();
            for (int j = warp; j < tq; j += SCAN_THREADS / 32) {
                size_t lq = (size_t)part * a.nq + (q0 + j);
                warp_offer_keys(skeys + j * SCAN_THREADS, a.lists + lq * a.k, a.counts + lq, a.thr + lq, a.k, lane);
            }
        }
    }
}

// ---- block-wide bitonic sort of M (power of two) u64 keys in shared memory, ascending ---------------------
// The stages of stride <= 32 of every bitonic step run inside one warp on a 64-key chunk held in registers (two keys
// per lane, partners reached with shuffles): only the strides >= 64 go through shared memory with block-wide barriers.
// 512 keys: 9 barriers instead of 45 (the sort was a third of finalize_kernel's time, and the whole latency of the merge
// kernels of a single-query search).
__device__ __forceinline__ void warp_bitonic_strides(uint64_t& x0, uint64_t& x1, int base, int lane, int size, int first_stride) {
    // x0 = key base + lane, x1 = key base + 32 + lane; the compare-exchanges of strides first_stride (<= 32) .. 1 of step `size`
    for (int stride = first_stride; stride > 0; stride >>= 1) {
        if (stride == 32) {
            const bool up = ((base + lane) & size) == 0;   // size >= 64 here: both keys of the pair see the same direction
            const uint64_t lo = x0 < x1 ? x0 : x1, hi = x0 < x1 ? x1 : x0;
            x0 = up ? lo : hi;
            x1 = up ? hi : lo;
        } else {
            const uint64_t y0 = __shfl_xor_sync(0xffffffffu, x0, stride), y1 = __shfl_xor_sync(0xffffffffu, x1, stride);
            const bool lower = (lane & stride) == 0;
            const bool up0 = ((base + lane) & size) == 0, up1 = ((base + 32 + lane) & size) == 0;
            const uint64_t mn0 = x0 < y0 ? x0 : y0, mx0 = x0 < y0 ? y0 : x0;
            const uint64_t mn1 = x1 < y1 ? x1 : y1, mx1 = x1 < y1 ? y1 : x1;
            x0 = (lower == up0) ? mn0 : mx0;
            x1 = (lower == up1) ? mn1 : mx1;
        }
    }
}
__device__ inline void bitonic_sort_u64(uint64_t* s, int M) {
    if (M < 64 || (blockDim.x & 31u) != 0u) {
        for (int size = 2; size <= M; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                __syncthreads();
                // one thread per compare-exchange pair (i, i + stride): pair t maps to the t-th index whose `stride` bit is clear
                for (int t = threadIdx.x; t < (M >> 1); t += blockDim.x) {
                    const int i = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
                    const int p = i | stride;
                    const bool up = (i & size) == 0;
                    const uint64_t x = s[i], y = s[p];
                    if ((x > y) == up) {
                        s[i] = y;
                        s[p] = x;
                    }
                }
            }
        }
        __syncthreads();
        return;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    __syncthreads();
    // steps 2 .. 64: every 64-key chunk is sorted by one warp, in the direction step 64 asks of it
    for (int base = warp * 64; base < M; base += nwarps * 64) {
        uint64_t x0 = s[base + lane], x1 = s[base + 32 + lane];
        for (int size = 2; size <= 64; size <<= 1) warp_bitonic_strides(x0, x1, base, lane, size, size >> 1);
        s[base + lane] = x0;
        s[base + 32 + lane] = x1;
    }
    for (int size = 128; size <= M; size <<= 1) {
        for (int stride = size >> 1; stride >= 64; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (M >> 1); t += blockDim.x) {
                const int i = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
                const int p = i | stride;
                const bool up = (i & size) == 0;
                const uint64_t x = s[i], y = s[p];
                if ((x > y) == up) {
                    s[i] = y;
                    s[p] = x;
                }
            }
        }
        __syncthreads();
        for (int base = warp * 64; base < M; base += nwarps * 64) {
            uint64_t x0 = s[base + lane], x1 = s[base + 32 + lane];
            warp_bitonic_strides(x0, x1, base, lane, size, 32);
            s[base + lane] = x0;
            s[base + 32 + lane] = x1;
        }
    }
    __syncthreads();
}

// ---- merge of the P per-partition lists (merge_topk_results + VectorStore::merge_results) ------------------
struct MergeArgs {
    const uint64_t* lists;   // [P][nq][k]
    const uint32_t* counts;  // [P][nq]
    int P, nq, k;
    int M;                   // sort width (power of two, > kp2)
    int kp2;                 // next power of two >= k
    int asc;
    int sqrt_scores;         // Jensen-Shannon cached path ranks on the squared distance (flat_mmap.rs:1106-1110)
    const uint32_t* qmap;    // optional: result slot of local query q
    uint32_t* out_rows;      // [*][k]
    float* out_dists;        // [*][k]
    uint32_t* out_counts;    // [*]
};

static __global__ void __launch_bounds__(1024) merge_lists_kernel(MergeArgs a) {
    extern __shared__ __align__(16) unsigned char smem_merge[];
    uint64_t* s = reinterpret_cast<uint64_t*>(smem_merge);
    const int q = blockIdx.x;
    const int k = a.k;
    for (int i = threadIdx.x; i < a.M; i += blockDim.x) s[i] = KEY_NONE;
    __syncthreads();
    const int per_round = a.M - a.kp2;
    const long total = (long)a.P * k;
    for (long base = 0; base < total; base += per_round) {
        for (int i = threadIdx.x; i < per_round; i += blockDim.x) {
            long slot = base + i;
            uint64_t key = KEY_NONE;
            if (slot < total) {
                int p = (int)(slot / k), j = (int)(slot - (long)p * k);
                size_t lq = (size_t)p * a.nq + q;
                if ((uint32_t)j < a.counts[lq]) key = a.lists[lq * k + j];
            }
            s[a.kp2 + i] = key;
        }
        bitonic_sort_u64(s, a.M);
    }
    const int slot = a.qmap ? (int)a.qmap[q] : q;
    __shared__ uint32_t n_valid;
    if (threadIdx.x == 0) n_valid = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        uint64_t key = s[i];
        uint32_t row = ROW_NONE;
        float score = __int_as_float(0x7fc00000);
        if (key != KEY_NONE) {
            row = key_row(key);
            score = a.asc ? key_score<true>(key) : key_score<false>(key);
            if (a.sqrt_scores) score = sqrtf(score);
            atomicAdd(&n_valid, 1u);
        }
        a.out_rows[(size_t)slot * k + i] = row;
        a.out_dists[(size_t)slot * k + i] = score;
    }
    __syncthreads();
    if (threadIdx.x == 0) a.out_counts[slot] = n_valid;
}

// The same merge for small k (k <= 32, P*k <= 4096 keys): no sort.  Every thread holds up to four keys; each warp
// pulls its k smallest out in order (k warp-wide minimum reductions), then warp 0 merges the 32 sorted runs the same
// way.  A single-query search over a small corpus spends most of its device time in the merge: this one is a few
// microseconds where the 2048-wide bitonic sort takes 66 block-wide barriers.
__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v) {
    const uint32_t hi = (uint32_t)(v >> 32), lo = (uint32_t)v;
    const uint32_t mhi = __reduce_min_sync(0xffffffffu, hi);
    const uint32_t mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xFFFFFFFFu);
    return ((uint64_t)mhi << 32) | mlo;
}

constexpr int MERGE_SMALL_MAX_K = 32;
constexpr int MERGE_SMALL_MAX_KEYS = 4096;

static __global__ void __launch_bounds__(1024) merge_lists_small_kernel(MergeArgs a) {
    __shared__ uint64_t runs[32][MERGE_SMALL_MAX_K];
    const int q = blockIdx.x, k = a.k, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int total = a.P * k;
    uint64_t mine[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int slot = tid + j * 1024;
        uint64_t key = KEY_NONE;
        if (slot < total) {
            const int p = slot / k, jj = slot - p * k;
            const size_t lq = (size_t)p * a.nq + q;
            if ((uint32_t)jj < a.counts[lq]) key = a.lists[lq * k + jj];
        }
        mine[j] = key;
    }
    for (int r = 0; r < k; ++r) {
        uint64_t m = mine[0];
#pragma unroll
        for (int j = 1; j < 4; ++j) m = mine[j] < m ? mine[j] : m;
        const uint64_t w = warp_min_u64(m);
        if (w != KEY_NONE) {  // keys are unique: exactly one holder
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (mine[j] == w) mine[j] = KEY_NONE;
        }
        if (lane == 0) runs[warp][r] = w;
    }
    __syncthreads();
    if (warp != 0) return;
    const int slot = a.qmap ? (int)a.qmap[q] : q;
    int head = 0;
    uint32_t n_valid = 0;
    for (int r = 0; r < k; ++r) {
        const uint64_t front = head < k ? runs[lane][head] : KEY_NONE;
        const uint64_t w = warp_min_u64(front);
        if (w != KEY_NONE && front == w) ++head;
        if (lane == 0) {
            uint32_t row = ROW_NONE;
            float score = __int_as_float(0x7fc00000);
            if (w != KEY_NONE) {
                row = key_row(w);
                score = a.asc ? key_score<true>(w) : key_score<false>(w);
                if (a.sqrt_scores) score = sqrtf(score);
                ++n_valid;
            }
            a.out_rows[(size_t)slot * k + r] = row;
            a.out_dists[(size_t)slot * k + r] = score;
        }
    }
    if (lane == 0) a.out_counts[slot] = n_valid;
}

// ---- side-structure builders ---------------------------------------------------------------------------------
// pack_binary_f32 (simd.rs:750-757, flat_mmap.rs:1283-1290): bit = x > 0.5, word i/64 bit i%64.  One warp per row.
template <class RT>
__global__ void pack_binary_kernel(const RT* __restrict__ rows, uint64_t n, int dim, int n_words, float threshold,
                                   uint64_t* __restrict__ out) {
    uint64_t row = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= n) return;
    const RT* r = rows + row * dim;
    for (int w = 0; w < n_words; ++w) {
        int i0 = w * 64 + lane, i1 = i0 + 32;
        unsigned lo = __ballot_sync(0xffffffffu, i0 < dim && ldrow(r + i0) > threshold);
        unsigned hi = __ballot_sync(0xffffffffu, i1 < dim && ldrow(r + i1) > threshold);
        if (lane == 0) out[row * n_words + w] = ((uint64_t)hi << 32) | lo;
    }
}

// probability_row_stats for every row (flat_mmap.rs:949-983) — also used for the queries.
template <class RT>
__global__ void row_stats_kernel(const RT* __restrict__ rows, uint64_t n, int dim, float* __restrict__ stats) {
    uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    float inv, ent;
    probability_row_stats(rows + row * dim, dim, (dim & 3) == 0, &inv, &ent);
    stats[2 * row] = inv;
    stats[2 * row + 1] = ent;
}

// prepare_jensen_shannon_query (flat_mmap.rs:926-947): q * inv_mass; flag = 1 when the cached path cannot serve it.
static __global__ void js_prepare_queries_kernel(const float* __restrict__ queries, int nq, int dim,
                                          const float* __restrict__ qstats, float* __restrict__ nq_out,
                                          uint32_t* __restrict__ unhandled) {
    int q = blockIdx.x;
    float inv = qstats[2 * q], ent = qstats[2 * q + 1];
    bool bad = (inv != inv) || !isfinite(ent) || (inv != 0.0f && !isfinite(inv));
    if (threadIdx.x == 0) unhandled[q] = bad ? 1u : 0u;
    for (int d = threadIdx.x; d < dim; d += blockDim.x) {
        float v = queries[(size_t)q * dim + d];
        nq_out[(size_t)q * dim + d] = bad ? v : v * inv;
    }
}

// one pair (py_compute_distance)
static __global__ void pair_distance_kernel(const float* a, const float* b, int dim, int metric, float* out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = compute_distance<true>(metric, a, b, dim, (dim & 3) == 0);
}

static __global__ void fill_f32_kernel(float* out, uint64_t n, float value) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = value;
}

// synthetic corpora
static __global__ void synth_f32_kernel(float* out, uint64_t n_elems, uint64_t seed, uint64_t elem_offset) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n_elems; i += stride) out[i] = synth_f32(seed, elem_offset + i);
}
static __global__ void synth_u64_kernel(uint64_t* out, uint64_t n_elems, uint64_t seed, uint64_t elem_offset) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n_elems; i += stride) out[i] = synth_u64(seed, elem_offset + i);
}

}  // namespace lb
