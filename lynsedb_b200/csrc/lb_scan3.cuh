// lb_scan3.cuh — exact scan of the 8-lane f32 metrics for BATCHES of queries (a dozen and more): register tiles of
// 4 rows x 16 queries per thread, the block of rows resident in shared memory.
//
// Same reference loops and the same per-pair arithmetic as lb_scan2.cuh (fused_topk_parallel and friends,
// src/storage/flat_mmap.rs:4845-4982; the AVX2+FMA kernels of src/distance/simd.rs).  What changes is who owns what.
// lb_scan2.cuh gives a thread one row and all eight AVX lanes of up to eight queries: every 16 bytes of query that
// come out of shared memory feed eight arithmetic instructions, the load/store unit saturates at a quarter of the
// FP32 rate (Q=64 on 4M x 256: L1 15 ms), and the rows are re-read from L2 once per query tile.  Here EIGHT THREADS
// share a row: thread i owns AVX lane i (elements i, i+8, i+16, ... — its accumulator sees exactly the products the
// reference's lane i sees, in the same order) of R=4 rows and TQ=16 queries.  One step loads 4 row words and 16 query
// words and issues 64 (x2) arithmetic instructions: 0.16 shared-memory words per instruction, under the 0.25 the
// hardware can deliver.  The block's rows (64 KiB, TMA, 128-byte swizzle) stay in shared memory for all query tiles,
// so the corpus comes from HBM once per batch.  After the chunk loop the eight lane accumulators of a pair meet in a
// per-warp scratch area and one thread finishes the pair with Scan2Op::finish — the reference's horizontal reduction,
// scalar tail and final formula, unchanged.
#pragma once
#include "lb_scan2.cuh"

namespace lb {

constexpr int S4_R = 4;  // rows per group of eight threads

__host__ __device__ inline bool scan4_supported(int metric) {
    // (Canberra stays with lb_scan2.cuh: its IEEE division per element is the bound, and measured through this kernel it
    // was twice as slow — 90 against 47 ms at 64 queries over 4M x 256)
    return metric == LB_IP || metric == LB_L2 || metric == LB_COSINE || metric == LB_MANHATTAN || metric == LB_CHEBYSHEV ||
           metric == LB_BRAY_CURTIS;
}

// CTA shapes (lb_scan_dense.cuh picks per metric): up to 256 dims a 64-row block (at most 64 KiB, so two CTAs fit an SM and
// one computes while the other waits for its rows) worked by eight warps — two per row group, each with half of the tile's
// sixteen queries — or by four; up to 512 dims two warps over a 32-row block with eight queries per tile.
// scan4_warps: 8 = the 64-row shapes, 2 = the 32-row shape.  LYNSE_B200_SCAN_TILE_NW forces 8 or 2.
inline int scan4_warps(int dim) {
    const char* env = getenv("LYNSE_B200_SCAN_TILE_NW");
    if (env && (atoi(env) == 8 || atoi(env) == 2)) return atoi(env);
    return dim <= 256 ? 8 : 2;
}
inline int scan4_block_rows(int nw) { return nw == 8 ? 64 : 32; }

// TQW: queries per tile in shared memory; QSP: how many warps share a row group, each taking TQW / QSP of the tile's
// queries (QSP = 2: eight warps per CTA over the same 64 rows, 16 warps per SM instead of 8 — the chunk loop leaves a
// third of the issue slots empty with two warps per scheduler).  R x (TQW / QSP) x accumulators per pair <= 64 registers.
template <int METRIC, bool IP2, int TQW, int QSP = 1>
struct S4Cfg {
    using Op = Scan2Op<METRIC, IP2>;
    static constexpr int kTQ = TQW;
    static constexpr int kTQT = TQW / QSP;                   // queries per thread
    static_assert(TQW % QSP == 0 && kTQT % 4 == 0, "a thread reads its queries as 16-byte pieces");
    static_assert(S4_R * kTQT * (Op::kState / 8) <= 64, "64 accumulators per thread at most");
    static constexpr int kSRow = kTQT * Op::kState + 8;      // scratch words per row group of a warp: = 8 mod 32, conflict-free
    static constexpr int kQStride = kTQ + 4;                 // words per element of the transposed query tile (20 / 12)
    // one query tile as the kernel wants it in shared memory: [dim & ~7][kQStride] element-major, then [kTQ][8] tail elements
    static size_t tile_floats(int dim) { return (size_t)(dim & ~7) * kQStride + (size_t)kTQ * 8; }
    // top-k lists of every query of the batch in shared memory (keys, counts, gates) when they take at most 16 KiB
    static bool smem_lists(int nq, int k) { return (size_t)nq * k * 8 + (size_t)nq * 16 <= 16384; }
    // shared memory of a CTA of NW warps for rows of `dim` floats (+ 1 KiB alignment slack)
    static size_t smem_bytes(int nw, int dim, int nq, int k) {
        const size_t rb = (size_t)(nw / QSP) * 4 * S4_R;
        const size_t rows = (size_t)((dim + 31) / 32) * rb * 128;
        const size_t cand = (size_t)kTQ * rb * 8;
        const size_t q = (size_t)(dim & ~7) * kQStride * 4 + (size_t)2 * kTQ * 8 * 4;
        const size_t scratch = (size_t)nw * 4 * kSRow * 4;
        const size_t misc = (size_t)kTQ * 16 + 64 + (METRIC == LB_COSINE ? (size_t)((nq + 3) & ~3) * 4 : 0);
        const size_t lists = smem_lists(nq, k) ? (size_t)nq * k * 8 + (size_t)((nq + 1) & ~1) * 12 + 16 : 0;
        return rows + cand + q + scratch + misc + lists + 1024;
    }
};

// Query tiles in the kernel's shared-memory layout, written once per search: a tile then arrives with one bulk copy.
// out: [n_tiles][tile_floats]; element e < (dim & ~7) of query t of the tile at e * qs + t, tail element x at tail_off + t * 8 + x.
static __global__ void scan4_query_tiles_kernel(const float* __restrict__ queries, int nq, int dim, int tq, int qs, float* __restrict__ out) {
    const int tail0 = dim & ~7;
    const size_t tile_floats = (size_t)tail0 * qs + (size_t)tq * 8;
    const int tile = blockIdx.x;
    float* o = out + (size_t)tile * tile_floats;
    for (size_t i = threadIdx.x; i < tile_floats; i += blockDim.x) o[i] = 0.0f;
    __syncthreads();
    for (int idx = threadIdx.x; idx < tq * dim; idx += blockDim.x) {
        const int t = idx / dim, d = idx - t * dim;
        const int q = tile * tq + t;
        if (q < nq) {
            const float v = __ldg(queries + (size_t)q * dim + d);
            if (d < tail0) o[(size_t)d * qs + t] = v;
            else o[(size_t)tail0 * qs + t * 8 + (d - tail0)] = v;
        }
    }
}

__device__ __forceinline__ void bulk_copy_g2s(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst), "l"(gsrc), "r"(bytes),
                 "r"(bar)
                 : "memory");
}

// one element of AVX lane i: (q, c) into the pair's accumulator(s), exactly as Scan2Op::step does for that lane
template <int METRIC, bool IP2, bool ODD>
__device__ __forceinline__ void lane_step(float& s0, float& s1, float q, float c, bool two_acc) {
    if constexpr (METRIC == LB_IP) {
        if (IP2 && ODD) {
            if (two_acc) s1 = fmaf(q, c, s1);
            else s0 = fmaf(q, c, s0);
        } else {
            s0 = fmaf(q, c, s0);
        }
    } else if constexpr (METRIC == LB_L2) {
        const float d = q - c;
        if (ODD) s1 = fmaf(d, d, s1);
        else s0 = fmaf(d, d, s0);
    } else if constexpr (METRIC == LB_COSINE) {
        s0 = fmaf(q, c, s0);   // the row norm (s1 of the reference's state) is per row: taken by the caller
    } else if constexpr (METRIC == LB_MANHATTAN) {
        s0 = s0 + fabsf(q - c);
    } else if constexpr (METRIC == LB_CHEBYSHEV) {
        s0 = max_ps(s0, fabsf(q - c));
    } else if constexpr (METRIC == LB_CANBERRA) {
        const float num = fabsf(q - c);
        const float den = fabsf(q) + fabsf(c);
        const float quot = num / den;
        s0 = s0 + ((den != 0.0f) ? quot : 0.0f);
    } else {  // Bray-Curtis
        s0 = s0 + fabsf(q - c);
        s1 = s1 + fabsf(q + c);
    }
}

template <int METRIC, bool IP2, int NW, int TQW, int QSP = 1>
__global__ void __launch_bounds__(NW * 32, 2) scan_tile_kernel(const __grid_constant__ CUtensorMap tmap, ScanArgs a) {
    using Op = Scan2Op<METRIC, IP2>;
    using Cfg = S4Cfg<METRIC, IP2, TQW, QSP>;
    constexpr int R = S4_R, TQ = Cfg::kTQ, TQT = Cfg::kTQT, KS = Op::kState, SROW = Cfg::kSRow, QS = Cfg::kQStride;
    constexpr int RB = (NW / QSP) * 4 * R, NT = NW * 32;
    constexpr bool ASC = METRIC != LB_IP;
    constexpr bool TWO = KS == 16 && METRIC != LB_COSINE;  // two accumulators per pair
    extern __shared__ __align__(16) unsigned char smem_s4[];
    const uint32_t smem_base = (tc::smem_u32(smem_s4) + 1023u) & ~1023u;
    unsigned char* smem = smem_s4 + (smem_base - tc::smem_u32(smem_s4));
    const int dim = a.dim;
    const int chunks = dim >> 3, tail0 = chunks * 8;
    const int n_cc = (dim + 31) / 32;
    const int n_tiles = (a.nq + TQ - 1) / TQ;
    const uint32_t rows_bytes = (uint32_t)n_cc * RB * 128;
    const uint32_t qtile_bytes = (uint32_t)tail0 * QS * 4;                           // the element-major part of a query tile
    const size_t tile_floats = (size_t)tail0 * QS + (size_t)TQ * 8;
    uint64_t* cand = reinterpret_cast<uint64_t*>(smem + rows_bytes);                 // [TQ][RB]
    float* sqt = reinterpret_cast<float*>(cand + TQ * RB);                           // [tail0][QS]: element-major query tile
    float* stail = sqt + (size_t)tail0 * QS;                                         // [2][TQ][8]: elements tail0.. of each query
    float* scratch = stail + 2 * TQ * 8;                                             // [NW][4][SROW]
    uint64_t* tgate = reinterpret_cast<uint64_t*>(scratch + NW * 4 * SROW);          // [TQ] gates of the tile (global lists)
    uint32_t* scnt = reinterpret_cast<uint32_t*>(tgate + TQ);                        // [TQ]
    uint64_t* bar = reinterpret_cast<uint64_t*>(scnt + TQ + (TQ & 1));               // [0] rows, [1] query tile
    float* snorm = reinterpret_cast<float*>(bar + 2);                                // [nq] cosine: |q|^2
    // lists of the whole batch in shared memory (a.smem_lists): keys [nq][k], gates [nq], counts [nq]
    uint64_t* lkeys = reinterpret_cast<uint64_t*>(snorm + (METRIC == LB_COSINE ? ((a.nq + 3) & ~3) : 0));
    uint64_t* lgate = lkeys + (size_t)a.nq * a.k;
    uint32_t* lcnt = reinterpret_cast<uint32_t*>(lgate + a.nq);
    const bool slists = a.smem_lists != 0;
    const uint32_t bar_rows = tc::smem_u32(bar), bar_q = bar_rows + 8u;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, sub = lane >> 3, li = lane & 7;
    const int rwarp = warp / QSP, qh = warp % QSP;   // which 16 rows of the block, which TQT queries of the tile
    const int part = blockIdx.x;
    const uint64_t part_begin = (uint64_t)part * a.rows_per_part;
    uint64_t part_end = part_begin + a.rows_per_part;
    if (part_end > a.n_rows) part_end = a.n_rows;
    const uint32_t n_blocks = part_end > part_begin ? (uint32_t)((part_end - part_begin + RB - 1) / RB) : 0u;
    if (n_blocks == 0u) return;

    if (tid == 0) {
        tc::mbar_init(bar_rows, 1);
        tc::mbar_init(bar_q, 1);
        tc::fence_barrier_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    }
    if (slists) {
        for (int q = tid; q < a.nq; q += NT) {
            lgate[q] = KEY_NONE;
            lcnt[q] = 0u;
        }
    }
    if (tid < TQ) scnt[tid] = 0u;
    if (METRIC == LB_COSINE) {
        // |q|^2 of every query, once per CTA, in the reference's lane order (simd.rs:1583-1636)
        for (int q = tid; q < a.nq; q += NT) {
            const float* qp = a.queries + (size_t)q * dim;
            float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int j = 0; j < chunks; ++j) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float v = __ldg(qp + 8 * j + i);
                    acc[i] = fmaf(v, v, acc[i]);
                }
            }
            float na = hsum8(acc);
            for (int i = tail0; i < dim; ++i) na = na + __ldg(qp + i) * __ldg(qp + i);
            snorm[q] = na;
        }
    }
    // this thread's rows inside a block: the four row groups of a warp read rows two apart, so their 128-byte-swizzled
    // pieces fall into different banks
    uint32_t a0[R];
    int rl[R];
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
        rl[rr] = rwarp * (4 * R) + (rr >> 1) * 8 + sub * 2 + (rr & 1);
        a0[rr] = (uint32_t)rl[rr] * 128u + ((((uint32_t)li >> 2) ^ ((uint32_t)rl[rr] & 7u)) << 4) + ((uint32_t)li & 3u) * 4u;
    }
    __syncthreads();
    // the query tiles arrive one bulk copy each, the next one while the current tile's pairs are finished and folded
    auto fetch_tile = [&](int tile) {  // thread 0 only
        tc::mbar_arrive_expect_tx(bar_q, qtile_bytes);
        bulk_copy_g2s(tc::smem_u32(sqt), a.query_tiles + (size_t)tile * tile_floats, qtile_bytes, bar_q);
    };
    if (tid == 0 && qtile_bytes > 0) fetch_tile(0);

    // a block of rows: one TMA box per 32-float column chunk, all on one barrier; the next block is requested as soon as
    // the last query tile's chunk loop is over, under the finishing and folding of that tile
    auto fetch_rows = [&](uint32_t blk) {  // thread 0 only
        const uint64_t r0 = part_begin + (uint64_t)blk * RB;
        tc::mbar_arrive_expect_tx(bar_rows, rows_bytes);
        for (int cc = 0; cc < n_cc; ++cc) tc::tma_load_2d(smem_base + (uint32_t)cc * (RB * 128), &tmap, cc * 32, (int)r0, bar_rows);
    };
    if (tid == 0) fetch_rows(0);

    uint32_t phase_rows = 0, phase_q = 0;
    uint32_t tile_seq = 0;  // tiles consumed so far (selects the tail buffer)
    for (uint32_t blk = 0; blk < n_blocks; ++blk) {
        const uint64_t row0 = part_begin + (uint64_t)blk * RB;
        bool valid[R], two_acc[R];
        uint32_t row[R];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const uint64_t slot = row0 + (uint64_t)rl[rr];
            row[rr] = (uint32_t)slot;
            valid[rr] = slot < part_end && row_allowed(a.allow_bits, row[rr]);
            two_acc[rr] = METRIC == LB_IP && IP2 && (a.ip_single || (valid[rr] && a.n_small > 0 && in_small_segment(a.small_seg, a.n_small, row[rr])));
        }
        while (!tc::mbar_try_wait(bar_rows, phase_rows)) {
        }
        phase_rows ^= 1u;

        for (int tile = 0; tile < n_tiles; ++tile, ++tile_seq) {
            const int q0 = tile * TQ;
            const int tq = min(TQ, a.nq - q0);
            float* mytail = stail + (tile_seq & 1u) * (TQ * 8);
            // tail elements of the tile (dim % 8 of them per query) and, with global lists, the gates of its queries:
            // consumed after the barrier that follows the chunk loop
            if (tail0 < dim)
                for (int i = tid; i < TQ * 8; i += NT) mytail[i] = __ldg(a.query_tiles + (size_t)tile * tile_floats + (size_t)tail0 * QS + i);
            if (!slists && tid < TQ) {
                uint64_t g = 0ull;  // a gate no key passes (queries past the end of the batch)
                if (tid < tq) {
                    const size_t lq = (size_t)part * a.nq + (q0 + tid);
                    g = __ldcg(a.counts + lq) < (uint32_t)a.k ? KEY_NONE : __ldcg(a.thr + lq);
                }
                tgate[tid] = g;
            }
            if (qtile_bytes > 0) {
                while (!tc::mbar_try_wait(bar_q, phase_q)) {
                }
                phase_q ^= 1u;
            }

            float acc0[R][TQT], acc1[R][TQT], nb[R];
#pragma unroll
            for (int rr = 0; rr < R; ++rr) {
                nb[rr] = 0.0f;
#pragma unroll
                for (int t = 0; t < TQT; ++t) {
                    acc0[rr][t] = 0.0f;
                    acc1[rr][t] = 0.0f;
                }
            }
            // four chunks per round: the row words sit at a0 ^ (s << 5) of the round's column chunk, the query words
            // 8 * QS floats further per chunk — running pointers, no index arithmetic in the loop
            const unsigned char* rptr = smem;
            const float* qptr = sqt + (size_t)li * QS + qh * TQT;
            auto body = [&](auto s_tag) {
                constexpr int S = decltype(s_tag)::value;
                constexpr bool ODD = (S & 1) != 0;
                float cv[R];
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                    cv[rr] = *reinterpret_cast<const float*>(rptr + (a0[rr] ^ (uint32_t)(S << 5)));
                    if (METRIC == LB_COSINE) nb[rr] = fmaf(cv[rr], cv[rr], nb[rr]);
                }
                const float4* qp = reinterpret_cast<const float4*>(qptr + S * 8 * QS);
#pragma unroll
                for (int t4 = 0; t4 < TQT / 4; ++t4) {
                    const float4 q4 = qp[t4];
                    const float qv[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
#pragma unroll
                        for (int rr = 0; rr < R; ++rr)
                            lane_step<METRIC, IP2, ODD>(acc0[rr][4 * t4 + u], acc1[rr][4 * t4 + u], qv[u], cv[rr], two_acc[rr]);
                    }
                }
            };
            int j = 0;
#pragma unroll 1
            for (; j + 4 <= chunks; j += 4) {
                body(std::integral_constant<int, 0>{});
                body(std::integral_constant<int, 1>{});
                body(std::integral_constant<int, 2>{});
                body(std::integral_constant<int, 3>{});
                rptr += RB * 128;
                qptr += 32 * QS;
            }
            if (j < chunks) body(std::integral_constant<int, 0>{});
            if (j + 1 < chunks) body(std::integral_constant<int, 1>{});
            if (j + 2 < chunks) body(std::integral_constant<int, 2>{});

            __syncthreads();  // the query tile is no longer read; the previous tile's candidates are folded; tails and gates are visible
            if (tid == 0) {
                if (qtile_bytes > 0 && !(blk + 1 == n_blocks && tile + 1 == n_tiles)) fetch_tile(tile + 1 == n_tiles ? 0 : tile + 1);
                if (tile + 1 == n_tiles && blk + 1 < n_blocks) fetch_rows(blk + 1);  // this block's rows are no longer read
            }

            // the eight lanes of a pair meet in the warp's scratch; thread (sub, li) finishes queries li (and li + 8) of the
            // warp's share of the tile, for row rr
            float* ws = scratch + (size_t)(warp * 4 + sub) * SROW;
            const uint64_t* gate = slists ? lgate + q0 : tgate;
#pragma unroll
            for (int rr = 0; rr < R; ++rr) {
                __syncwarp();
#pragma unroll
                for (int t = 0; t < TQT; ++t) {
                    ws[t * KS + li] = acc0[rr][t];
                    if (KS == 16) ws[t * KS + 8 + li] = TWO ? acc1[rr][t] : nb[rr];
                }
                __syncwarp();
                if (valid[rr]) {
                    const float* c = a.corpus + (size_t)row[rr] * dim;
#pragma unroll
                    for (int tt = 0; tt < (TQT + 7) / 8; ++tt) {
                        const int tl = li + 8 * tt;            // query of this warp's share
                        const int t = qh * TQT + tl;           // ... of the tile
                        if (tl < TQT && t < tq) {
                            float s[KS];
#pragma unroll
                            for (int x = 0; x < KS; x += 4) {
                                const float4 v4 = *reinterpret_cast<const float4*>(ws + tl * KS + x);
                                s[x] = v4.x; s[x + 1] = v4.y; s[x + 2] = v4.z; s[x + 3] = v4.w;
                            }
                            PairConst pc;
                            const float v = Op::finish(s, mytail + t * 8 - tail0, c, tail0, dim, two_acc[rr], METRIC == LB_COSINE ? snorm[q0 + t] : 0.0f, pc);
                            const uint64_t key = make_key<ASC>(v, row[rr]);
                            if (key < gate[t]) {
                                const uint32_t pos = atomicAdd(&scnt[t], 1u);
                                cand[t * RB + pos] = key;
                            }
                        }
                    }
                }
            }
            __syncthreads();
            for (int t = warp; t < tq; t += NW) {
                const int n = (int)scnt[t];
                if (n > 0) {
                    if (slists) {
                        warp_fold_candidates_t<false>(cand + t * RB, n, lkeys + (size_t)(q0 + t) * a.k, lcnt + q0 + t, lgate + q0 + t, a.k, lane);
                    } else {
                        const size_t lq = (size_t)part * a.nq + (q0 + t);
                        warp_fold_candidates(cand + t * RB, n, a.lists + lq * a.k, a.counts + lq, a.thr + lq, a.k, lane);
                    }
                    if (lane == 0) scnt[t] = 0u;
                }
            }
        }
    }
    if (slists) {
        __syncthreads();
        for (int q = warp; q < a.nq; q += NW) {
            const size_t lq = (size_t)part * a.nq + q;
            const uint32_t cnt = lcnt[q];
            for (int i = lane; i < (int)cnt; i += 32) a.lists[lq * a.k + i] = lkeys[(size_t)q * a.k + i];
            if (lane == 0) {
                a.counts[lq] = cnt;
                a.thr[lq] = lgate[q];
            }
        }
    }
}

}  // namespace lb
