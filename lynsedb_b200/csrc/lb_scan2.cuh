// lb_scan2.cuh — streaming exact scan for the 8-lane f32 metrics (IP, L2, cosine, L1, Chebyshev, Canberra,
// Bray-Curtis) and the sequential f64 ones (correlation, Hellinger, Wasserstein): every corpus row is read ONCE per
// query tile and scored against up to 8 queries at a time.
//
// Replaces the same reference loops as lb_scan.cuh (fused_topk_ip_parallel / ip_scan_chunk_topk,
// fused_topk_parallel, fused_topk_parallel_filtered, direct_access_topk — src/storage/flat_mmap.rs:4845-4982,
// :5223-5274, :5439-5554) with the same per-pair arithmetic order as lb_metrics.cuh (i.e. as
// src/distance/simd.rs's AVX2+FMA kernels): the eight AVX lanes are eight accumulators per (row, query), the
// 8-float chunks are consumed in order, the horizontal reductions and the scalar tails follow the same tree.
// What changes against lb_scan.cuh is the loop nest: lb_scan.cuh evaluates one whole pair at a time and so
// re-reads the row (from L1) once per query; here the chunk loop is outermost and the accumulators of all the
// queries of the tile live in registers, which is what lets a small batch run at HBM speed.
#pragma once
#include "lb_packed.cuh"

namespace lb {

constexpr int S2_ROWS = 256;   // rows per block step == threads per CTA

__host__ __device__ inline bool scan2_f64(int metric) {
    return metric == LB_CORRELATION || metric == LB_HELLINGER || metric == LB_WASSERSTEIN;
}
__host__ __device__ inline bool scan2_supported(int metric) {
    return metric == LB_IP || metric == LB_L2 || metric == LB_COSINE || metric == LB_MANHATTAN || metric == LB_CHEBYSHEV ||
           metric == LB_CANBERRA || metric == LB_BRAY_CURTIS || scan2_f64(metric) || metric == LB_JENSEN_SHANNON;
}   // Wasserstein needs ScanArgs::row_mass, Jensen-Shannon ScanArgs::row_stats (the cached form; queries mass-normalised)

// Per-pair constants of the sequential f64 metrics.  Their reference loops (simd.rs:632-714) interleave sums that
// depend on one operand only with the cross terms; those are independent accumulator chains, so the query-side sums
// are taken once per query (qa, qb) and Wasserstein's row mass once per row (ra), bit for bit the same values.
struct PairConst {
    double qa = 0, qb = 0;  // correlation: sum a, sum a^2; Hellinger / Wasserstein: sum a (NaN: invalid value in the query)
    double ra = 0, rb = 0;  // Wasserstein: sum b (NaN: invalid value in the row) and its reciprocal; qb = 1 / qa there
    float q_inv = 0, q_ent = 0, r_inv = 0, r_ent = 0;  // Jensen-Shannon cached scan: (inverse mass, entropy) of query and row
};
template <int METRIC>
__device__ inline void scan2_query_consts(const float* __restrict__ q /*smem*/, int dim, double* out /*[2]*/) {
    double s = 0, ss = 0;
    bool bad = false;
    for (int i = 0; i < dim; ++i) {
        const double av = (double)q[i];
        s = s + av;
        if (METRIC == LB_CORRELATION) ss = ss + av * av;
        else bad = bad || invalid_mass_value(q[i]);
    }
    out[0] = bad ? __longlong_as_double(0x7ff8000000000000ll) : s;
    out[1] = ss;
}
// Wasserstein row masses, once per row (the first loop of wasserstein_1d_f32, simd.rs:691-698)
template <class RT>
__global__ void row_mass_kernel(const RT* __restrict__ rows, uint64_t n, int dim, double* __restrict__ mass) {
    const uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    const RT* c = rows + row * dim;
    const bool vec = (dim & 3) == 0;
    double s = 0;
    bool bad = false;
    const int chunks = dim >> 3;
    for (int j = 0; j < chunks; ++j) {
        const Vec8 b = load8<true>(c + 8 * j, vec);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            bad = bad || invalid_mass_value(b.v[i]);
            s = s + (double)b.v[i];
        }
    }
    for (int i = chunks * 8; i < dim; ++i) {
        const float b = ldrow(c + i);
        bad = bad || invalid_mass_value(b);
        s = s + (double)b;
    }
    mass[row] = bad ? __longlong_as_double(0x7ff8000000000000ll) : s;
}

// state layout per (row, query): s[0..7] first accumulator vector, s[8..15] second (metrics that have one)
// IP2: the launch may contain IP rows that take the two-accumulator (single-row) kernel — rows of segments under
// 4096 rows, or the stateless operator; without it IP needs one accumulator vector only.
template <int METRIC, bool IP2>
struct Scan2Op {
    static constexpr bool kF64 = METRIC == LB_CORRELATION || METRIC == LB_HELLINGER || METRIC == LB_WASSERSTEIN;
    using T = typename std::conditional<kF64, double, float>::type;
    static constexpr int kState = kF64 ? (METRIC == LB_WASSERSTEIN ? 2 : 3)
                                  : (((METRIC == LB_IP && IP2) || METRIC == LB_L2 || METRIC == LB_COSINE || METRIC == LB_BRAY_CURTIS) ? 16 : 8);
    static constexpr int kTQ = kF64 ? 4 : 64 / kState;  // queries per tile: <= 64 accumulator registers per thread

    // one 8-float chunk, index j; `two_acc` = IP rows that take the two-accumulator kernel
    static __device__ __forceinline__ void step(T* s, const Vec8& q, const Vec8& c, int j, bool two_acc, int dim, const PairConst& pc) {
        const bool odd = (j & 1) != 0;
        if constexpr (METRIC == LB_CORRELATION) {  // simd.rs:632-661: s = {sum b, sum b^2, sum ab}
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const double av = (double)q.v[i], bv = (double)c.v[i];
                s[0] = s[0] + bv;
                s[1] = s[1] + bv * bv;
                s[2] = s[2] + av * bv;
            }
        } else if constexpr (METRIC == LB_HELLINGER) {  // simd.rs:665-684: s = {sum b, sum sqrt(ab), invalid row value seen}
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (invalid_mass_value(c.v[i])) s[2] = 1.0;
                s[0] = s[0] + (double)c.v[i];
                s[1] = s[1] + sqrt((double)q.v[i] * (double)c.v[i]);
            }
        } else if constexpr (METRIC == LB_WASSERSTEIN) {  // simd.rs:688-714, second loop: s = {cdf delta, distance}
            const double inv_a = pc.qb, inv_b = pc.rb;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (8 * j + i < dim - 1) {  // the last bin never contributes
                    s[0] = s[0] + ((double)q.v[i] * inv_a - (double)c.v[i] * inv_b);
                    s[1] = s[1] + fabs(s[0]);
                }
            }
        } else if constexpr (METRIC == LB_JENSEN_SHANNON) {  // mixture term of the cached form, simd.rs:2316-2354
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float sum = q.v[i] + c.v[i] * pc.r_inv;
                s[i] = fmaf(sum, fast_ln<true>(max_ps(sum, kMinPositive)), s[i]);
            }
        } else if constexpr (METRIC == LB_IP) {
            // batch-8 order: one accumulator (simd.rs:1450-1525); single-row order: even chunks -> acc0, odd -> acc1
            // (simd.rs:1341-1396).  The trailing unpaired chunk has an even index, so it lands in acc0 as it must.
            if (IP2 && two_acc && odd) {
#pragma unroll
                for (int i = 0; i < 8; ++i) s[8 + i] = fmaf(q.v[i], c.v[i], s[8 + i]);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) s[i] = fmaf(q.v[i], c.v[i], s[i]);
            }
        } else if constexpr (METRIC == LB_L2) {  // simd.rs:1527-1581
            if (odd) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float d = q.v[i] - c.v[i];
                    s[8 + i] = fmaf(d, d, s[8 + i]);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float d = q.v[i] - c.v[i];
                    s[i] = fmaf(d, d, s[i]);
                }
            }
        } else if constexpr (METRIC == LB_COSINE) {  // simd.rs:1583-1636 (the query norm is the same chain for every row: done once)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                s[i] = fmaf(q.v[i], c.v[i], s[i]);
                s[8 + i] = fmaf(c.v[i], c.v[i], s[8 + i]);
            }
        } else if constexpr (METRIC == LB_MANHATTAN) {  // simd.rs:2134-2158
#pragma unroll
            for (int i = 0; i < 8; ++i) s[i] = s[i] + fabsf(q.v[i] - c.v[i]);
        } else if constexpr (METRIC == LB_CHEBYSHEV) {  // simd.rs:2715-2737
#pragma unroll
            for (int i = 0; i < 8; ++i) s[i] = max_ps(s[i], fabsf(q.v[i] - c.v[i]));
        } else if constexpr (METRIC == LB_CANBERRA) {  // simd.rs:2762-2793
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float num = fabsf(q.v[i] - c.v[i]);
                const float den = fabsf(q.v[i]) + fabsf(c.v[i]);
                const float quot = num / den;
                s[i] = s[i] + ((den != 0.0f) ? quot : 0.0f);
            }
        } else {  // Bray-Curtis, simd.rs:2824-2865
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                s[i] = s[i] + fabsf(q.v[i] - c.v[i]);
                s[8 + i] = s[8 + i] + fabsf(q.v[i] + c.v[i]);
            }
        }
    }

    // horizontal reduction + scalar tail (elements [tail0, dim)) + final formula
    template <class CP>
    static __device__ __forceinline__ float finish(T* s, const float* __restrict__ q /*smem*/, CP c /*global*/,
                                                   int tail0, int dim, bool two_acc, float q_norm2, const PairConst& pc) {
        if constexpr (METRIC == LB_CORRELATION) {
            if (dim == 0) return 0.0f;
            for (int i = tail0; i < dim; ++i) {
                const double av = (double)q[i], bv = (double)ldrow(c + i);
                s[0] = s[0] + bv;
                s[1] = s[1] + bv * bv;
                s[2] = s[2] + av * bv;
            }
            const double n = (double)dim, sa = pc.qa, saa = pc.qb, sb = s[0], sbb = s[1], sab = s[2];
            const double var_a = fmax(saa - sa * sa / n, 0.0);
            const double var_b = fmax(sbb - sb * sb / n, 0.0);
            const double denom = sqrt(var_a * var_b);
            if (denom <= 2.2204460492503131e-16) {
                bool same = true;
                for (int i = 0; i < dim; ++i) same = same && (q[i] == ldrow(c + i));
                return same ? 0.0f : 1.0f;
            }
            const double cov = sab - sa * sb / n;
            return (float)(1.0 - clamp_f64(cov / denom, -1.0, 1.0));
        } else if constexpr (METRIC == LB_HELLINGER) {
            for (int i = tail0; i < dim; ++i) {
                const float b = ldrow(c + i);
                if (invalid_mass_value(b)) s[2] = 1.0;
                s[0] = s[0] + (double)b;
                s[1] = s[1] + sqrt((double)q[i] * (double)b);
            }
            const double sa = pc.qa, sb = s[0];
            if (s[2] != 0.0 || sa != sa) return INFINITY;
            if (sa == 0.0 || sb == 0.0) return sa == sb ? 0.0f : 1.0f;
            const double cc = s[1] / sqrt(sa * sb);
            return (float)sqrt(1.0 - clamp_f64(cc, 0.0, 1.0));
        } else if constexpr (METRIC == LB_WASSERSTEIN) {
            const double sa = pc.qa, sb = pc.ra;
            if (sa != sa || sb != sb) return INFINITY;
            if (sa == 0.0 || sb == 0.0) return sa == sb ? 0.0f : INFINITY;
            const double inv_a = pc.qb, inv_b = pc.rb;
            for (int i = tail0; i < dim - 1; ++i) {
                s[0] = s[0] + ((double)q[i] * inv_a - (double)ldrow(c + i) * inv_b);
                s[1] = s[1] + fabs(s[0]);
            }
            return (float)s[1];
        } else if constexpr (METRIC == LB_JENSEN_SHANNON) {
            // the ranking value of jensen_shannon_cached_parallel (flat_mmap.rs:985-1111): squared distance
            const bool vec = (dim & 3) == 0;
            if (pc.q_inv == 0.0f) {  // zero-mass query (flat_mmap.rs:938-972)
                if (pc.r_inv != pc.r_inv || !isfinite(pc.r_ent)) return INFINITY;
                return pc.r_inv == 0.0f ? 0.0f : kLn2;
            }
            if (pc.r_inv <= 0.0f || !isfinite(pc.r_inv) || !isfinite(pc.r_ent)) {
                const float d = jensen_shannon_precomputed<false>(q, c, dim, vec, pc.q_ent, pc.r_inv, pc.r_ent);
                return d * d;
            }
            float mix = lane_sum8(s);
            for (int i = tail0; i < dim; ++i) {
                const float sm = q[i] + ldrow(c + i) * pc.r_inv;
                if (sm > 0.0f) mix = mix + sm * logf(sm);
            }
            const float divergence = fmaxf(kLn2 + 0.5f * (pc.q_ent + pc.r_ent - mix), 0.0f);
            if (divergence <= kJsStableDivergence) {
                const float d = jensen_shannon_normalized_query<false>(q, c, dim, vec, pc.r_inv);
                return d * d;
            }
            return divergence;
        } else if constexpr (METRIC == LB_IP) {
            if (IP2 && two_acc) {
#pragma unroll
                for (int i = 0; i < 8; ++i) s[i] = s[i] + s[8 + i];
            }
            float out = hsum8(s);
            for (int i = tail0; i < dim; ++i) out = out + q[i] * ldrow(c + i);
            return out;
        } else if constexpr (METRIC == LB_L2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) s[i] = s[i] + s[8 + i];
            float sum = hsum8(s);
            for (int i = tail0; i < dim; ++i) {
                const float diff = q[i] - ldrow(c + i);
                sum = sum + diff * diff;
            }
            return sum;
        } else if constexpr (METRIC == LB_COSINE) {
            float dot = hsum8(s), nb = hsum8(s + 8), na = q_norm2;
            for (int i = tail0; i < dim; ++i) {
                const float a = q[i], b = ldrow(c + i);
                dot = dot + a * b;
                nb = nb + b * b;
            }
            const float denom = sqrtf(na * nb);
            if (denom < 1e-30f) return 1.0f;
            return 1.0f - dot / denom;
        } else if constexpr (METRIC == LB_MANHATTAN) {
            float sum = lane_sum8(s);
            for (int i = tail0; i < dim; ++i) sum = sum + fabsf(q[i] - ldrow(c + i));
            return sum;
        } else if constexpr (METRIC == LB_CHEBYSHEV) {
            float m = 0.0f;
#pragma unroll
            for (int i = 0; i < 8; ++i) m = rust_max(m, s[i]);
            for (int i = tail0; i < dim; ++i) m = rust_max(m, fabsf(q[i] - ldrow(c + i)));
            return m;
        } else if constexpr (METRIC == LB_CANBERRA) {
            float sum = lane_sum8(s);
            for (int i = tail0; i < dim; ++i) {
                const float a = q[i], b = ldrow(c + i);
                const float den = fabsf(a) + fabsf(b);
                if (den != 0.0f) sum = sum + fabsf(a - b) / den;
            }
            return sum;
        } else {
            float num = lane_sum8(s), den = lane_sum8(s + 8);
            for (int i = tail0; i < dim; ++i) {
                const float a = q[i], b = ldrow(c + i);
                num = num + fabsf(a - b);
                den = den + fabsf(a + b);
            }
            if (den == 0.0f) return num == 0.0f ? 0.0f : INFINITY;
            return num / den;
        }
    }
};

// the query-side half of the cosine kernel (simd.rs:1583-1636): |q|^2 with the same lane order, once per query
__device__ inline float cosine_query_norm2(const float* __restrict__ q /*smem*/, int dim) {
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int chunks = dim >> 3;
    for (int j = 0; j < chunks; ++j) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaf(q[8 * j + i], q[8 * j + i], a[i]);
    }
    float na = hsum8(a);
    for (int i = chunks * 8; i < dim; ++i) na = na + q[i] * q[i];
    return na;
}

template <int METRIC, bool IP2, class RT = float>
__global__ void __launch_bounds__(S2_ROWS, 2) scan_stream_kernel(ScanArgs a) {
    using Op = Scan2Op<METRIC, IP2>;
    constexpr int S2_TQ = Op::kTQ;
    constexpr bool ASC = METRIC != LB_IP;
    extern __shared__ __align__(16) unsigned char smem_s2[];
    const int dim = a.dim, dim_pad = (dim + 3) & ~3;
    uint64_t* cand = reinterpret_cast<uint64_t*>(smem_s2);                              // [TQ][256]
    float* sq = reinterpret_cast<float*>(smem_s2 + S2_TQ * S2_ROWS * 8);                // [TQ][dim_pad]
    uint64_t* sthr = reinterpret_cast<uint64_t*>(sq + S2_TQ * dim_pad);                 // [TQ]
    uint32_t* scnt = reinterpret_cast<uint32_t*>(sthr + S2_TQ);                         // [TQ]
    float* sna = reinterpret_cast<float*>(scnt + S2_TQ);                                // [TQ] cosine |q|^2
    double* sqc = reinterpret_cast<double*>(sna + S2_TQ);                               // [TQ][2] f64 metrics: query sums
    SmemLists sl;
    sl.keys = reinterpret_cast<uint64_t*>(sqc + 2 * S2_TQ);                             // [TQ][k] when a.smem_lists
    sl.counts = reinterpret_cast<uint32_t*>(sl.keys + (size_t)S2_TQ * a.k);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int part = blockIdx.x;
    const bool vec = (dim & 3) == 0;
    const int chunks = dim >> 3;
    const uint64_t part_begin = (uint64_t)part * a.rows_per_part;
    uint64_t part_end = part_begin + a.rows_per_part;
    if (part_end > a.n_rows) part_end = a.n_rows;
    const bool single_tile = a.nq <= S2_TQ;

    auto load_tile = [&](int q0, int tq) {
        for (int i = tid; i < tq * dim; i += S2_ROWS) {
            const int qq = i / dim, d = i - qq * dim;
            sq[qq * dim_pad + d] = __ldg(a.queries + (size_t)(q0 + qq) * dim + d);
        }
        __syncthreads();
        if (METRIC == LB_COSINE && tid < tq) sna[tid] = cosine_query_norm2(sq + tid * dim_pad, dim);
        if (Op::kF64 && tid < tq) scan2_query_consts<METRIC>(sq + tid * dim_pad, dim, sqc + 2 * tid);
    };
    if (single_tile) {
        load_tile(0, a.nq);
        if (tid < a.nq) {
            sthr[tid] = KEY_NONE;  // the lists of this launch start empty (counts are zeroed by the host)
            scnt[tid] = 0u;
        }
        if (a.smem_lists) sl.init(a.nq, tid, S2_ROWS);
        __syncthreads();
    }

    for (uint64_t blk = part_begin; blk < part_end; blk += S2_ROWS) {
        const uint64_t slot = blk + tid;
        bool valid = slot < part_end;
        uint32_t row = 0;
        if (valid) row = a.row_ids ? __ldg(a.row_ids + slot) : (uint32_t)slot;
        if (valid && !row_allowed(a.allow_bits, row)) valid = false;
        const RT* c = corpus_rows<RT>(a) + (size_t)row * dim;
        if (a.row_ids != nullptr) {
            // gathered rows (IVF lists, id filters): neighbouring threads read unrelated rows, so nothing arrives in L2
            // ahead of the one-chunk-ahead loads and every 32-byte piece costs a DRAM round trip (measured: 2.3 us per
            // chunk).  Ask L2 for the whole row now — the next block's row, and on the first block this one's too.
            auto prefetch_row = [&](uint32_t r) {
                const char* p = reinterpret_cast<const char*>(corpus_rows<RT>(a) + (size_t)r * dim);
                for (int off = 0; off < dim * (int)sizeof(RT); off += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + off));
            };
            if (blk == part_begin && valid) prefetch_row(row);
            const uint64_t nslot = slot + S2_ROWS;
            if (nslot < part_end) prefetch_row(__ldg(a.row_ids + nslot));
        }
        const bool two_acc = METRIC == LB_IP && (a.ip_single || (valid && a.n_small > 0 && in_small_segment(a.small_seg, a.n_small, row)));
        for (int q0 = 0; q0 < a.nq; q0 += S2_TQ) {
            const int tq = min(S2_TQ, a.nq - q0);
            if (!single_tile) {
                __syncthreads();  // the previous tile's queries and candidates are no longer read
                load_tile(q0, tq);
            }
            if (!single_tile && tid < tq) {
                const size_t lq = (size_t)part * a.nq + (q0 + tid);
                sthr[tid] = __ldcg(a.counts + lq) < (uint32_t)a.k ? KEY_NONE : __ldcg(a.thr + lq);
                scnt[tid] = 0u;
            }
            if (!single_tile) __syncthreads();
            if (valid) {
                typename Op::T st[S2_TQ][Op::kState];
                PairConst pc[S2_TQ];
#pragma unroll
                for (int t = 0; t < S2_TQ; ++t) {
#pragma unroll
                    for (int i = 0; i < Op::kState; ++i) st[t][i] = 0;
                    if (Op::kF64 && t < tq) {
                        pc[t].qa = sqc[2 * t];
                        pc[t].qb = sqc[2 * t + 1];
                        if (METRIC == LB_WASSERSTEIN) {
                            pc[t].qb = 1.0 / pc[t].qa;
                            pc[t].ra = __ldg(a.row_mass + row);
                            pc[t].rb = 1.0 / pc[t].ra;
                        }
                    }
                    if (METRIC == LB_JENSEN_SHANNON && t < tq) {
                        pc[t].q_inv = __ldg(a.query_stats + 2 * (q0 + t));
                        pc[t].q_ent = __ldg(a.query_stats + 2 * (q0 + t) + 1);
                        pc[t].r_inv = __ldg(a.row_stats + 2 * (size_t)row);
                        pc[t].r_ent = __ldg(a.row_stats + 2 * (size_t)row + 1);
                    }
                }
                // chunk loop outermost: the row streams through registers once, one chunk ahead of the arithmetic
                Vec8 cv = chunks > 0 ? load8<true>(c, vec) : Vec8{};
                for (int j = 0; j < chunks; ++j) {
                    Vec8 nx = cv;
                    if (j + 1 < chunks) nx = load8<true>(c + 8 * (j + 1), vec);
                    // one row per thread means one 128-byte line per thread in flight at a time: ask for the lines
                    // two ahead so that enough bytes are outstanding to cover the HBM latency
                    if ((j & 3) == 0 && j + 8 < chunks) asm volatile("prefetch.global.L1 [%0];" ::"l"(c + 8 * (j + 8)));
#pragma unroll
                    for (int t = 0; t < S2_TQ; ++t) {
                        if (t < tq) {
                            const Vec8 qv = load8<false>(sq + t * dim_pad + 8 * j, vec);  // broadcast
                            Op::step(st[t], qv, cv, j, two_acc, dim, pc[t]);
                        }
                    }
                    cv = nx;
                }
#pragma unroll
                for (int t = 0; t < S2_TQ; ++t) {
                    if (t < tq) {
                        const float v = Op::finish(st[t], sq + t * dim_pad, c, chunks * 8, dim, two_acc, METRIC == LB_COSINE ? sna[t] : 0.0f, pc[t]);
                        const uint64_t key = make_key<ASC>(v, row);
                        if (key < sthr[t]) {
                            const uint32_t pos = atomicAdd(&scnt[t], 1u);
                            cand[t * S2_ROWS + pos] = key;
                        }
                    }
                }
            }
            __syncthreads();
            for (int j = warp; j < tq; j += S2_ROWS / 32) {
                const int n = (int)scnt[j];
                if (n > 0) {
                    const size_t lq = (size_t)part * a.nq + (q0 + j);
                    const uint64_t g = (single_tile && a.smem_lists)
                                           ? warp_fold_candidates_t<false>(cand + j * S2_ROWS, n, sl.keys + (size_t)j * a.k, sl.counts + j, sthr + j, a.k, lane)
                                           : warp_fold_candidates(cand + j * S2_ROWS, n, a.lists + lq * a.k, a.counts + lq, a.thr + lq, a.k, lane);
                    if (lane == 0) {
                        sthr[j] = g;   // a single-tile batch keeps its gates in shared memory across row blocks
                        scnt[j] = 0u;
                    }
                }
            }
            if (single_tile) __syncthreads();  // candidates folded before the next block reuses the buffers
        }
    }
    if (single_tile && a.smem_lists) sl.write_back(a, part, a.nq, sthr, warp, lane, S2_ROWS / 32);
}

// ---- the same scan with the rows staged through shared memory by TMA -------------------------------------------
// One row per thread reading straight from global memory touches 32 rows per load instruction, 32 bytes each:
// HBM tops out near 3.7 TB/s on that pattern.  Here a block of 256 rows is fetched in column chunks of 32 floats:
// one TMA box [256 rows x 128 B] per chunk (SWIZZLE_128B, so thread r finds 16-byte piece c of its row at piece
// c ^ (r & 7) and the 128-bit reads are bank-conflict free), four stages in flight.  The chunk order, and with it
// every accumulator chain, is unchanged.
constexpr int S3_NSTAGES = 6;                  // 192 KiB in flight per SM: the pass is latency-bound below that
constexpr int S3_STAGE_BYTES = S2_ROWS * 128;  // 32 KiB: 256 rows x 32 floats
constexpr int S3_TQ = 4;                       // queries per tile (this kernel serves batches of <= 4)

// binary16 rows carry twice the arithmetic per staged byte: two CTAs per SM (16 warps), each with a ring of two stages (the
// query tile, candidate buffer and lists of a CTA take ~21 KiB, so three stages would not leave room for two CTAs), overlap
// one CTA's arithmetic with the other's loads.
template <class RT>
struct S3Cfg {
    static constexpr int kStages = std::is_same<RT, float>::value ? S3_NSTAGES : 2;
    static constexpr int kCtasPerSm = std::is_same<RT, float>::value ? 1 : 2;
};

template <int METRIC, bool IP2, class RT = float>
__global__ void __launch_bounds__(S2_ROWS, S3Cfg<RT>::kCtasPerSm) scan_stream_tma_kernel(const __grid_constant__ CUtensorMap tmap, ScanArgs a) {
    constexpr int EPB = 128 / (int)sizeof(RT);   // elements per 128-byte stage row: 32 floats or 64 halves
    constexpr int CPB = EPB / 8;                 // 8-element chunks per stage row
    constexpr int NST = S3Cfg<RT>::kStages;
    using Op = Scan2Op<METRIC, IP2>;
    constexpr int S2_TQ = Op::kTQ < S3_TQ ? Op::kTQ : S3_TQ;
    constexpr bool ASC = METRIC != LB_IP;
    extern __shared__ __align__(16) unsigned char smem_s3[];
    const uint32_t smem_base = (tc::smem_u32(smem_s3) + 1023u) & ~1023u;
    unsigned char* smem = smem_s3 + (smem_base - tc::smem_u32(smem_s3));
    const int dim = a.dim, dim_pad = (dim + 3) & ~3;
    uint64_t* cand = reinterpret_cast<uint64_t*>(smem + NST * S3_STAGE_BYTES);   // [TQ][256]
    float* sq = reinterpret_cast<float*>(cand + S2_TQ * S2_ROWS);                       // [TQ][dim_pad]
    uint64_t* sthr = reinterpret_cast<uint64_t*>(sq + S2_TQ * dim_pad);                 // [TQ]
    uint32_t* scnt = reinterpret_cast<uint32_t*>(sthr + S2_TQ);                         // [TQ]
    float* sna = reinterpret_cast<float*>(scnt + S2_TQ);                                // [TQ]
    double* sqc = reinterpret_cast<double*>(sna + S2_TQ + 2);                           // [TQ][2] f64 metrics: query sums
    uint64_t* bars = reinterpret_cast<uint64_t*>(sqc + 2 * S2_TQ);                      // full[NSTAGES]
    const uint32_t full0 = tc::smem_u32(bars);
    SmemLists sl;
    sl.keys = bars + NST;                                                        // [TQ][k] when a.smem_lists
    sl.counts = reinterpret_cast<uint32_t*>(sl.keys + (size_t)S2_TQ * a.k);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int part = blockIdx.x;
    const int chunks = dim >> 3;                 // whole 8-float chunks (the rest is the scalar tail)
    const int n_cc = (dim + EPB - 1) / EPB;      // column chunks of 128 bytes per row block
    const uint64_t part_begin = (uint64_t)part * a.rows_per_part;
    uint64_t part_end = part_begin + a.rows_per_part;
    if (part_end > a.n_rows) part_end = a.n_rows;
    const uint32_t n_blocks = part_end > part_begin ? (uint32_t)((part_end - part_begin + S2_ROWS - 1) / S2_ROWS) : 0u;
    const int n_tiles = (a.nq + S2_TQ - 1) / S2_TQ;
    const bool single_tile = n_tiles == 1;
    // the stream of boxes this CTA consumes: for every row block, for every query tile, its n_cc column chunks
    const uint64_t n_boxes = (uint64_t)n_blocks * n_tiles * n_cc;

    if (tid == 0) {
        for (int s = 0; s < NST; ++s) tc::mbar_init(full0 + 8u * s, 1);
        tc::fence_barrier_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    }
    auto issue = [&](uint64_t box) {  // thread 0 only
        const uint32_t blk = (uint32_t)(box / ((uint64_t)n_tiles * n_cc));
        const int cc = (int)(box % n_cc);
        const uint32_t stage = (uint32_t)(box % NST);
        tc::mbar_arrive_expect_tx(full0 + 8u * stage, S3_STAGE_BYTES);
        tc::tma_load_2d(smem_base + stage * S3_STAGE_BYTES, &tmap, cc * EPB, (int)(part_begin + (uint64_t)blk * S2_ROWS), full0 + 8u * stage);
    };
    auto load_tile = [&](int q0, int tq) {
        for (int i = tid; i < tq * dim; i += S2_ROWS) {
            const int qq = i / dim, d = i - qq * dim;
            sq[qq * dim_pad + d] = __ldg(a.queries + (size_t)(q0 + qq) * dim + d);
        }
        __syncthreads();
        if (METRIC == LB_COSINE && tid < tq) sna[tid] = cosine_query_norm2(sq + tid * dim_pad, dim);
        if (Op::kF64 && tid < tq) scan2_query_consts<METRIC>(sq + tid * dim_pad, dim, sqc + 2 * tid);
    };
    __syncthreads();
    if (tid == 0)
        for (uint64_t b = 0; b < (uint64_t)NST && b < n_boxes; ++b) issue(b);
    if (single_tile) {
        load_tile(0, a.nq);
        if (tid < a.nq) {
            sthr[tid] = KEY_NONE;
            scnt[tid] = 0u;
        }
        if (a.smem_lists) sl.init(a.nq, tid, S2_ROWS);
        __syncthreads();
    }

    uint64_t box = 0;
    for (uint32_t blk = 0; blk < n_blocks; ++blk) {
        const uint64_t slot = part_begin + (uint64_t)blk * S2_ROWS + tid;
        bool valid = slot < part_end;
        const uint32_t row = (uint32_t)slot;
        if (valid && !row_allowed(a.allow_bits, row)) valid = false;
        const RT* c = corpus_rows<RT>(a) + (size_t)row * dim;
        const bool two_acc = METRIC == LB_IP && (a.ip_single || (valid && a.n_small > 0 && in_small_segment(a.small_seg, a.n_small, row)));
        for (int q0 = 0; q0 < a.nq; q0 += S2_TQ) {
            const int tq = min(S2_TQ, a.nq - q0);
            if (!single_tile) {
                __syncthreads();
                load_tile(q0, tq);
                if (tid < tq) {
                    const size_t lq = (size_t)part * a.nq + (q0 + tid);
                    sthr[tid] = __ldcg(a.counts + lq) < (uint32_t)a.k ? KEY_NONE : __ldcg(a.thr + lq);
                    scnt[tid] = 0u;
                }
                __syncthreads();
            }
            typename Op::T st[S2_TQ][Op::kState];
            PairConst pc[S2_TQ];
#pragma unroll
            for (int t = 0; t < S2_TQ; ++t) {
#pragma unroll
                for (int i = 0; i < Op::kState; ++i) st[t][i] = 0;
                if (Op::kF64 && t < tq) {
                    pc[t].qa = sqc[2 * t];
                    pc[t].qb = sqc[2 * t + 1];
                    if (METRIC == LB_WASSERSTEIN && valid) {
                        pc[t].qb = 1.0 / pc[t].qa;
                        pc[t].ra = __ldg(a.row_mass + row);
                        pc[t].rb = 1.0 / pc[t].ra;
                    }
                }
            }
            for (int cc = 0; cc < n_cc; ++cc, ++box) {
                const uint32_t stage = (uint32_t)(box % NST), phase = (uint32_t)((box / NST) & 1u);
                while (!tc::mbar_try_wait(full0 + 8u * stage, phase)) {
                }
                // my 128 bytes of this column chunk: 8 pieces of 16 bytes, piece p at p ^ (row & 7), kept raw (binary16 rows
                // are decoded chunk by chunk below: half the registers of holding 64 floats)
                uint4 raw[8];
                {
                    const uint4* rowp = reinterpret_cast<const uint4*>(smem + stage * S3_STAGE_BYTES + tid * 128);
#pragma unroll
                    for (int p = 0; p < 8; ++p) raw[p] = rowp[p ^ (tid & 7)];
                }
                __syncthreads();  // every thread holds its piece: the stage can be refilled
                if (tid == 0 && box + NST < n_boxes) issue(box + NST);
                if (valid) {
#pragma unroll
                    for (int sub = 0; sub < CPB; ++sub) {
                        const int j = cc * CPB + sub;
                        Vec8 cvs;
                        if constexpr (std::is_same<RT, float>::value) {
                            const uint4 lo = raw[2 * sub], hi = raw[2 * sub + 1];
                            cvs.v[0] = __uint_as_float(lo.x); cvs.v[1] = __uint_as_float(lo.y); cvs.v[2] = __uint_as_float(lo.z); cvs.v[3] = __uint_as_float(lo.w);
                            cvs.v[4] = __uint_as_float(hi.x); cvs.v[5] = __uint_as_float(hi.y); cvs.v[6] = __uint_as_float(hi.z); cvs.v[7] = __uint_as_float(hi.w);
                        } else {
                            const uint32_t w[4] = {raw[sub].x, raw[sub].y, raw[sub].z, raw[sub].w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
                                cvs.v[2 * i] = f.x;
                                cvs.v[2 * i + 1] = f.y;
                            }
                        }
                        if (j < chunks) {
#pragma unroll
                            for (int t = 0; t < S2_TQ; ++t) {
                                if (t < tq) {
                                    const Vec8 qv = load8<false>(sq + t * dim_pad + 8 * j, true);  // broadcast
                                    Op::step(st[t], qv, cvs, j, two_acc, dim, pc[t]);
                                }
                            }
                        }
                    }
                }
            }
            if (valid) {
#pragma unroll
                for (int t = 0; t < S2_TQ; ++t) {
                    if (t < tq) {
                        const float v = Op::finish(st[t], sq + t * dim_pad, c, chunks * 8, dim, two_acc, METRIC == LB_COSINE ? sna[t] : 0.0f, pc[t]);
                        const uint64_t key = make_key<ASC>(v, row);
                        if (key < sthr[t]) {
                            const uint32_t pos = atomicAdd(&scnt[t], 1u);
                            cand[t * S2_ROWS + pos] = key;
                        }
                    }
                }
            }
            __syncthreads();
            for (int j = warp; j < tq; j += S2_ROWS / 32) {
                const int n = (int)scnt[j];
                if (n > 0) {
                    const size_t lq = (size_t)part * a.nq + (q0 + j);
                    const uint64_t g = (single_tile && a.smem_lists)
                                           ? warp_fold_candidates_t<false>(cand + j * S2_ROWS, n, sl.keys + (size_t)j * a.k, sl.counts + j, sthr + j, a.k, lane)
                                           : warp_fold_candidates(cand + j * S2_ROWS, n, a.lists + lq * a.k, a.counts + lq, a.thr + lq, a.k, lane);
                    if (lane == 0) {
                        sthr[j] = g;
                        scnt[j] = 0u;
                    }
                }
            }
            if (single_tile) __syncthreads();
        }
    }
    if (single_tile && a.smem_lists) sl.write_back(a, part, a.nq, sthr, warp, lane, S2_ROWS / 32);
}

}  // namespace lb
