// lb_metrics.cuh — exact per-pair distance functions, one CUDA thread per pair.
//
// Every function reproduces the arithmetic ORDER of the reference's x86-64
// AVX2+FMA kernel (reference src/distance/simd.rs; the branch taken under
// is_x86_feature_detected!("avx2") && "fma"): the eight AVX lanes become eight
// scalar accumulators in one thread, fused multiply-adds stay fused (fmaf),
// everything else is rounded separately (the library is compiled with
// -fmad=false so the compiler never contracts a*b+c on its own), and the
// horizontal reductions follow the same tree.  With IEEE division and square
// root (-prec-div=true -prec-sqrt=true, nvcc defaults) the f32 results are
// bit-identical to the CPU path; the transcendental tails (libm log / sin /
// cos / asin) agree to the last ulp or two.
#pragma once
#include <cuda_fp16.h>

#include "lb_common.cuh"

namespace lb {

constexpr float kLn2 = 0.693147180559945309417232121458176568f;
constexpr float kFrac1Sqrt2 = 0.707106781186547524400844362104849039f;
constexpr float kMinPositive = 1.17549435e-38f;
constexpr float kJsStableDivergence = 1e-5f;  // simd.rs:8

struct Vec8 {
    float v[8];
};

// 8 consecutive floats; `vec` = pointer is 16-byte aligned for every i (dim % 4 == 0)
template <bool GLOBAL>
__device__ __forceinline__ Vec8 load8(const float* __restrict__ p, bool vec) {
    Vec8 r;
    if (vec) {
        float4 a, b;
        if (GLOBAL) {
            a = __ldg(reinterpret_cast<const float4*>(p));
            b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        } else {
            a = reinterpret_cast<const float4*>(p)[0];
            b = reinterpret_cast<const float4*>(p)[1];
        }
        r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
        r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) r.v[i] = GLOBAL ? __ldg(p + i) : p[i];
    }
    return r;
}

// binary16 rows (float16 collections keep their rows as IEEE half in HBM; decoding is exact): the same 8 elements
template <bool GLOBAL>
__device__ __forceinline__ Vec8 load8(const __half* __restrict__ p, bool vec) {
    Vec8 r;
    if (vec && (reinterpret_cast<uintptr_t>(p) & 15u) == 0u) {
        const uint4 x = GLOBAL ? __ldg(reinterpret_cast<const uint4*>(p)) : *reinterpret_cast<const uint4*>(p);
        const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
            r.v[2 * i] = f.x;
            r.v[2 * i + 1] = f.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) r.v[i] = __half2float(GLOBAL ? __ldg(p + i) : p[i]);
    }
    return r;
}
// one row element from global memory
__device__ __forceinline__ float ldrow(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ldrow(const __half* p) { return __half2float(__ldg(p)); }

// simd.rs:1427-1436 — extractf128+add, movehdup+add, movehl+add_ss
__device__ __forceinline__ float hsum8(const float a[8]) {
    float s0 = a[0] + a[4], s1 = a[1] + a[5], s2 = a[2] + a[6], s3 = a[3] + a[7];
    return (s0 + s1) + (s2 + s3);
}
// `lanes.into_iter().sum()` (simd.rs:2151-2153): sequential from 0.0
__device__ __forceinline__ float lane_sum8(const float a[8]) {
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s = s + a[i];
    return s;
}
// Rust f32::max — NaN-ignoring
__device__ __forceinline__ float rust_max(float x, float y) {
    if (x != x) return y;
    if (y != y) return x;
    return x > y ? x : y;
}
// MAXPS(a, b) = a > b ? a : b
__device__ __forceinline__ float max_ps(float a, float b) { return a > b ? a : b; }

// ---- inner product -----------------------------------------------------------------
// batch-8 order (simd.rs:1450-1525): ONE accumulator vector per row.
template <bool QG, class CP>
__device__ float ip_batch8_order(const float* __restrict__ q, CP c, int dim, bool vec) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int chunks = dim >> 3;
#pragma unroll 2
    for (int j = 0; j < chunks; ++j) {
        Vec8 qv = load8<QG>(q + 8 * j, vec), cv = load8<true>(c + 8 * j, vec);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(qv.v[i], cv.v[i], acc[i]);
    }
    float out = hsum8(acc);
    for (int i = chunks * 8; i < dim; ++i) out = out + (QG ? __ldg(q + i) : q[i]) * ldrow(c + i);
    return out;
}
// single-row order (simd.rs:1341-1396): TWO accumulator vectors over a 16-stride.
template <bool QG, class CP>
__device__ float ip_single_order(const float* __restrict__ q, CP c, int dim, bool vec) {
    float acc0[8] = {0, 0, 0, 0, 0, 0, 0, 0}, acc1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int chunks = dim >> 3;
    int j = 0;
    for (; j + 1 < chunks; j += 2) {
        Vec8 qa = load8<QG>(q + 8 * j, vec), ca = load8<true>(c + 8 * j, vec);
        Vec8 qb = load8<QG>(q + 8 * j + 8, vec), cb = load8<true>(c + 8 * j + 8, vec);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            acc0[i] = fmaf(qa.v[i], ca.v[i], acc0[i]);
            acc1[i] = fmaf(qb.v[i], cb.v[i], acc1[i]);
        }
    }
    if (j < chunks) {
        Vec8 qa = load8<QG>(q + 8 * j, vec), ca = load8<true>(c + 8 * j, vec);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc0[i] = fmaf(qa.v[i], ca.v[i], acc0[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc0[i] = acc0[i] + acc1[i];
    float out = hsum8(acc0);
    for (int i = chunks * 8; i < dim; ++i) out = out + (QG ? __ldg(q + i) : q[i]) * ldrow(c + i);
    return out;
}

// ---- squared L2 (simd.rs:1527-1581) -------------------------------------------------
template <bool QG, class CP>
__device__ float l2_squared(const float* __restrict__ q, CP c, int dim, bool vec) {
    float acc0[8] = {0, 0, 0, 0, 0, 0, 0, 0}, acc1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int chunks = dim >> 3;
    int j = 0;
    for (; j + 1 < chunks; j += 2) {
        Vec8 qa = load8<QG>(q + 8 * j, vec), ca = load8<true>(c + 8 * j, vec);
        Vec8 qb = load8<QG>(q + 8 * j + 8, vec), cb = load8<true>(c + 8 * j + 8, vec);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float d0 = qa.v[i] - ca.v[i];
            acc0[i] = fmaf(d0, d0, acc0[i]);
            float d1 = qb.v[i] - cb.v[i];
            acc1[i] = fmaf(d1, d1, acc1[i]);
        }
    }
    if (j < chunks) {
        Vec8 qa = load8<QG>(q + 8 * j, vec), ca = load8<true>(c + 8 * j, vec);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float d0 = qa.v[i] - ca.v[i];
            acc0[i] = fmaf(d0, d0, acc0[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc0[i] = acc0[i] + acc1[i];
    float sum = hsum8(acc0);
    for (int i = chunks * 8; i < dim; ++i) {
        float diff = (QG ? __ldg(q + i) : q[i]) - ldrow(c + i);
        sum = sum + diff * diff;
    }
    return sum;
}

// ---- cosine distance (simd.rs:1583-1636) ----------------------------------------------
template <bool QG, class CP>
__device__ float cosine_distance(const float* __restrict__ q, CP c, int dim, bool vec) {
    float dacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, aacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, bacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int chunks = dim >> 3;
    for (int j = 0; j < chunks; ++j) {
        Vec8 a = load8<QG>(q + 8 * j, vec), b = load8<true>(c + 8 * j, vec);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            dacc[i] = fmaf(a.v[i], b.v[i], dacc[i]);
            aacc[i] = fmaf(a.v[i], a.v[i], aacc[i]);
            bacc[i] = fmaf(b.v[i], b.v[i], bacc[i]);
        }
    }
    float dot = hsum8(dacc), na = hsum8(aacc), nb = hsum8(bacc);
    for (int i = chunks * 8; i < dim; ++i) {
        float a = QG ? __ldg(q + i) : q[i], b = ldrow(c + i);
        dot = dot + a * b;
        na = na + a * a;
        nb = nb + b * b;
    }
    float denom = sqrtf(na * nb);
    if (denom < 1e-30f) return 1.0f;
    return 1.0f - dot / denom;
}

// ---- L1 (simd.rs:2134-2158) ---------------------------------------------------------------
template <bool QG, class CP>
__device__ float manhattan(const float* __restrict__ q, CP c, int dim, bool vec) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int chunks = dim >> 3;
    for (int j = 0; j < chunks; ++j) {
        Vec8 a = load8<QG>(q + 8 * j, vec), b = load8<true>(c + 8 * j, vec);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = acc[i] + fabsf(a.v[i] - b.v[i]);
    }
    float sum = lane_sum8(acc);
    for (int i = chunks * 8; i < dim; ++i) sum = sum + fabsf((QG ? __ldg(q + i) : q[i]) - ldrow(c + i));
    return sum;
}

// ---- Chebyshev (simd.rs:2715-2737) -----------------------------------------------------------
template <bool QG, class CP>
__device__ float chebyshev(const float* __restrict__ q, CP c, int dim, bool vec) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int chunks = dim >> 3;
    for (int j = 0; j < chunks; ++j) {
        Vec8 a = load8<QG>(q + 8 * j, vec), b = load8<true>(c + 8 * j, vec);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = max_ps(acc[i], fabsf(a.v[i] - b.v[i]));
    }
    float m = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) m = rust_max(m, acc[i]);
    for (int i = chunks * 8; i < dim; ++i) m = rust_max(m, fabsf((QG ? __ldg(q + i) : q[i]) - ldrow(c + i)));
    return m;
}

// ---- Canberra (simd.rs:2762-2793) ---------------------------------------------------------------
template <bool QG, class CP>
__device__ float canberra(const float* __restrict__ q, CP c, int dim, bool vec) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int chunks = dim >> 3;
    for (int j = 0; j < chunks; ++j) {
        Vec8 a = load8<QG>(q + 8 * j, vec), b = load8<true>(c + 8 * j, vec);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float num = fabsf(a.v[i] - b.v[i]);
            float den = fabsf(a.v[i]) + fabsf(b.v[i]);
            float quot = num / den;
            acc[i] = acc[i] + ((den != 0.0f) ? quot : 0.0f);  // and_ps(nonzero, quotient); NEQ_OQ is false on NaN
        }
    }
    float sum = lane_sum8(acc);
    for (int i = chunks * 8; i < dim; ++i) {
        float a = QG ? __ldg(q + i) : q[i], b = ldrow(c + i);
        float den = fabsf(a) + fabsf(b);
        if (den != 0.0f) sum = sum + fabsf(a - b) / den;
    }
    return sum;
}

// ---- Bray-Curtis (simd.rs:2824-2865) ----------------------------------------------------------------
template <bool QG, class CP>
__device__ float bray_curtis(const float* __restrict__ q, CP c, int dim, bool vec) {
    float nacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, dacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int chunks = dim >> 3;
    for (int j = 0; j < chunks; ++j) {
        Vec8 a = load8<QG>(q + 8 * j, vec), b = load8<true>(c + 8 * j, vec);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            nacc[i] = nacc[i] + fabsf(a.v[i] - b.v[i]);
            dacc[i] = dacc[i] + fabsf(a.v[i] + b.v[i]);
        }
    }
    float num = lane_sum8(nacc), den = lane_sum8(dacc);
    for (int i = chunks * 8; i < dim; ++i) {
        float a = QG ? __ldg(q + i) : q[i], b = ldrow(c + i);
        num = num + fabsf(a - b);
        den = den + fabsf(a + b);
    }
    if (den == 0.0f) return num == 0.0f ? 0.0f : INFINITY;
    return num / den;
}

// ---- thresholded binary metrics on f32 (simd.rs:175-209, :718-736) -----------------------------------
template <bool QG, class CP>
__device__ float hamming_f32(const float* __restrict__ q, CP c, int dim) {
    uint32_t count = 0;
    for (int i = 0; i < dim; ++i) count += (((QG ? __ldg(q + i) : q[i]) > 0.5f) != (ldrow(c + i) > 0.5f));
    return (float)count;
}
template <bool QG, class CP>
__device__ float jaccard_f32(const float* __restrict__ q, CP c, int dim) {
    uint32_t inter = 0, uni = 0;
    for (int i = 0; i < dim; ++i) {
        bool ab = (QG ? __ldg(q + i) : q[i]) > 0.5f, bb = ldrow(c + i) > 0.5f;
        uni += (ab || bb);
        inter += (ab && bb);
    }
    return uni == 0 ? 0.0f : 1.0f - ((float)inter / (float)uni);
}
template <bool QG, class CP>
__device__ float dice_f32(const float* __restrict__ q, CP c, int dim) {
    uint32_t inter = 0, ca = 0, cb = 0;
    for (int i = 0; i < dim; ++i) {
        bool ab = (QG ? __ldg(q + i) : q[i]) > 0.5f, bb = ldrow(c + i) > 0.5f;
        ca += ab;
        cb += bb;
        inter += (ab && bb);
    }
    uint32_t total = ca + cb;
    return total == 0 ? 0.0f : 1.0f - (float)(2 * inter) / (float)total;
}

// ---- scalar f64 metrics ---------------------------------------------------------------------------------
// Elements of a pair in index order: loads are 8 values at a time (two 128-bit loads per operand when `vec`), the
// arithmetic one element at a time — for the kernels whose reference loop is a plain sequential one.
template <bool QG, class F, class CP>
__device__ __forceinline__ void scalar_order_foreach(const float* __restrict__ q, CP c, int dim,
                                                     bool vec, F&& f) {
    int chunks = dim >> 3;
    for (int j = 0; j < chunks; ++j) {
        Vec8 a = load8<QG>(q + 8 * j, vec), b = load8<true>(c + 8 * j, vec);
#pragma unroll
        for (int i = 0; i < 8; ++i) f(a.v[i], b.v[i]);
    }
    for (int i = chunks * 8; i < dim; ++i) f(QG ? __ldg(q + i) : q[i], ldrow(c + i));
}

__device__ __forceinline__ double clamp_f64(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }
__device__ __forceinline__ bool invalid_mass_value(float v) { return !isfinite(v) || v < 0.0f; }

// simd.rs:603-628
template <bool QG, class CP>
__device__ float haversine_meters(const float* __restrict__ q, CP c, int dim) {
    if (dim != 2) return INFINITY;
    const double R = 6371008.8;
    const double k = 3.14159265358979323846264338327950288 / 180.0;
    float a0 = QG ? __ldg(q) : q[0], a1 = QG ? __ldg(q + 1) : q[1], b0 = ldrow(c), b1 = ldrow(c + 1);
    double lon1 = (double)a0 * k, lat1 = (double)a1 * k, lon2 = (double)b0 * k, lat2 = (double)b1 * k;
    if (!isfinite(lon1) || !isfinite(lat1) || !isfinite(lon2) || !isfinite(lat2) || fabsf(a1) > 90.0f ||
        fabsf(b1) > 90.0f)
        return INFINITY;
    double dlat = lat2 - lat1, dlon = lon2 - lon1;
    double sin_lat = sin(dlat * 0.5), sin_lon = sin(dlon * 0.5);
    double h = clamp_f64(sin_lat * sin_lat + cos(lat1) * cos(lat2) * sin_lon * sin_lon, 0.0, 1.0);
    return (float)(2.0 * R * asin(sqrt(h)));
}

// simd.rs:632-661
template <bool QG, class CP>
__device__ float correlation_distance(const float* __restrict__ q, CP c, int dim, bool vec) {
    if (dim == 0) return 0.0f;
    double n = (double)dim, sa = 0, sb = 0, saa = 0, sbb = 0, sab = 0;
    scalar_order_foreach<QG>(q, c, dim, vec, [&](float a, float b) {
        double av = (double)a, bv = (double)b;
        sa = sa + av;
        sb = sb + bv;
        saa = saa + av * av;
        sbb = sbb + bv * bv;
        sab = sab + av * bv;
    });
    double var_a = fmax(saa - sa * sa / n, 0.0);
    double var_b = fmax(sbb - sb * sb / n, 0.0);
    double denom = sqrt(var_a * var_b);
    if (denom <= 2.2204460492503131e-16) {
        bool same = true;
        for (int i = 0; i < dim; ++i) same = same && ((QG ? __ldg(q + i) : q[i]) == ldrow(c + i));
        return same ? 0.0f : 1.0f;
    }
    double cov = sab - sa * sb / n;
    return (float)(1.0 - clamp_f64(cov / denom, -1.0, 1.0));
}

// simd.rs:665-684 (the early return on an invalid value is the same +inf whichever element trips it)
template <bool QG, class CP>
__device__ float hellinger_distance(const float* __restrict__ q, CP c, int dim, bool vec) {
    double sa = 0, sb = 0, coef = 0;
    bool bad = false;
    scalar_order_foreach<QG>(q, c, dim, vec, [&](float a, float b) {
        bad = bad || invalid_mass_value(a) || invalid_mass_value(b);
        sa = sa + (double)a;
        sb = sb + (double)b;
        coef = coef + sqrt((double)a * (double)b);
    });
    if (bad) return INFINITY;
    if (sa == 0.0 || sb == 0.0) return sa == sb ? 0.0f : 1.0f;
    double cc = coef / sqrt(sa * sb);
    return (float)sqrt(1.0 - clamp_f64(cc, 0.0, 1.0));
}

// simd.rs:688-714; DIVIDE = wasserstein_1d_f16 (simd.rs:1046-1071), which divides by the masses instead of
// multiplying by their reciprocals
template <bool QG, bool DIVIDE, class CP>
__device__ float wasserstein_1d_impl(const float* __restrict__ q, CP c, int dim, bool vec) {
    double sa = 0, sb = 0;
    bool bad = false;
    scalar_order_foreach<QG>(q, c, dim, vec, [&](float a, float b) {
        bad = bad || invalid_mass_value(a) || invalid_mass_value(b);
        sa = sa + (double)a;
        sb = sb + (double)b;
    });
    if (bad) return INFINITY;
    if (sa == 0.0 || sb == 0.0) return sa == sb ? 0.0f : INFINITY;
    const double inv_a = 1.0 / sa, inv_b = 1.0 / sb;
    double cdf = 0, dist = 0;
    scalar_order_foreach<QG>(q, c, dim > 0 ? dim - 1 : 0, vec, [&](float a, float b) {  // the last bin never contributes
        cdf = cdf + (DIVIDE ? ((double)a / sa - (double)b / sb) : ((double)a * inv_a - (double)b * inv_b));
        dist = dist + fabs(cdf);
    });
    return (float)dist;
}
template <bool QG, class CP>
__device__ float wasserstein_1d(const float* __restrict__ q, CP c, int dim, bool vec) {
    return wasserstein_1d_impl<QG, false>(q, c, dim, vec);
}

// ---- Jensen-Shannon ----------------------------------------------------------------------------------------
// Cephes-style ln, one AVX lane (simd.rs:2164-2205; COMPACT = :2211-2247)
template <bool COMPACT>
__device__ __forceinline__ float fast_ln(float x) {
    uint32_t bits = __float_as_uint(x);
    uint32_t exponent_bits = bits >> 23;
    x = __uint_as_float((bits & 0x007fffffu) | 0x3f000000u);
    float exponent = (float)((int)exponent_bits - 0x7f);
    exponent = exponent + 1.0f;
    bool mask = x < kFrac1Sqrt2;
    float tmp = mask ? x : 0.0f;
    x = x - 1.0f;
    exponent = exponent - (mask ? 1.0f : 0.0f);
    x = x + tmp;
    float z = x * x;
    float y;
    if (!COMPACT) {
        y = 7.0376836E-2f;
        y = fmaf(y, x, -1.151461E-1f);
        y = fmaf(y, x, 1.1676998E-1f);
        y = fmaf(y, x, -1.2420141E-1f);
    } else {
        y = -1.2420141E-1f;
    }
    y = fmaf(y, x, 1.4249323E-1f);
    y = fmaf(y, x, -1.6668057E-1f);
    y = fmaf(y, x, 2.0000714E-1f);
    y = fmaf(y, x, -2.4999994E-1f);
    y = fmaf(y, x, 3.333333E-1f);
    y = (y * x) * z;
    y = fmaf(exponent, -2.1219444E-4f, y);
    y = fmaf(-z, 0.5f, y);  // fnmadd(z, 0.5, y)
    x = x + y;
    return fmaf(exponent, 0.6933594f, x);
}

// simd.rs:2249-2286
template <bool QG, class CP>
__device__ float jensen_shannon_avx(const float* __restrict__ a, CP b, int dim, bool vec,
                                    float inv_a, float inv_b) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int chunks = dim >> 3;
    for (int j = 0; j < chunks; ++j) {
        Vec8 va = load8<QG>(a + 8 * j, vec), vb = load8<true>(b + 8 * j, vec);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float p = va.v[i] * inv_a, q = vb.v[i] * inv_b;
            float m = (p + q) * 0.5f;
            float safe_m = max_ps(m, kMinPositive);
            float log_p = fast_ln<false>(max_ps(p, kMinPositive) / safe_m);
            float log_q = fast_ln<false>(max_ps(q, kMinPositive) / safe_m);
            float terms = p * log_p + q * log_q;
            acc[i] = fmaf(terms, 0.5f, acc[i]);
        }
    }
    float divergence = lane_sum8(acc);
    for (int i = chunks * 8; i < dim; ++i) {
        float p = (QG ? __ldg(a + i) : a[i]) * inv_a, q = ldrow(b + i) * inv_b;
        float m = 0.5f * (p + q);
        if (p > 0.0f) divergence = divergence + 0.5f * p * logf(p / m);
        if (q > 0.0f) divergence = divergence + 0.5f * q * logf(q / m);
    }
    return sqrtf(fmaxf(divergence, 0.0f));
}

// simd.rs:1161-1178
template <bool QG, class CP>
__device__ float jensen_shannon_scalar_f64(const float* __restrict__ a, CP b, int dim,
                                           double sum_a, double sum_b) {
    double inv_a = 1.0 / sum_a, inv_b = 1.0 / sum_b, divergence = 0;
    for (int i = 0; i < dim; ++i) {
        double p = (double)(QG ? __ldg(a + i) : a[i]) * inv_a, q = (double)ldrow(b + i) * inv_b, m = 0.5 * (p + q);
        if (p > 0.0) divergence = divergence + 0.5 * p * log(p / m);
        if (q > 0.0) divergence = divergence + 0.5 * q * log(q / m);
    }
    return (float)sqrt(fmax(divergence, 0.0));
}

// simd.rs:235-284 (+ refine_small_jensen_shannon, :1118-1125)
template <bool QG, class CP>
__device__ float jensen_shannon_distance(const float* __restrict__ a, CP b, int dim, bool vec) {
    double sum_a = 0, sum_b = 0;
    for (int i = 0; i < dim; ++i) {
        float x = QG ? __ldg(a + i) : a[i], y = ldrow(b + i);
        if (invalid_mass_value(x) || invalid_mass_value(y)) return INFINITY;
        sum_a = sum_a + (double)x;
        sum_b = sum_b + (double)y;
    }
    if (sum_a == 0.0 || sum_b == 0.0) return sum_a == sum_b ? 0.0f : sqrtf(kLn2);
    float inv_a = (float)(1.0 / sum_a), inv_b = (float)(1.0 / sum_b);
    if (!isfinite(inv_a) || !isfinite(inv_b) || inv_a == 0.0f || inv_b == 0.0f)
        return jensen_shannon_scalar_f64<QG>(a, b, dim, sum_a, sum_b);
    float distance = jensen_shannon_avx<QG>(a, b, dim, vec, inv_a, inv_b);
    if (distance * distance <= kJsStableDivergence) {
        bool same = true;
        for (int i = 0; i < dim; ++i) same = same && ((QG ? __ldg(a + i) : a[i]) == ldrow(b + i));
        if (!same) return jensen_shannon_scalar_f64<QG>(a, b, dim, sum_a, sum_b);
    }
    return distance;
}

// simd.rs:2288-2312
template <class RP>
__device__ inline float probability_entropy_avx(RP row, int dim, bool vec, float inv_mass) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int chunks = dim >> 3;
    for (int j = 0; j < chunks; ++j) {
        Vec8 r = load8<true>(row + 8 * j, vec);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float p = r.v[i] * inv_mass;
            float log_p = fast_ln<true>(max_ps(p, kMinPositive));
            acc[i] = fmaf(p, log_p, acc[i]);
        }
    }
    float entropy = lane_sum8(acc);
    for (int i = chunks * 8; i < dim; ++i) {
        float p = ldrow(row + i) * inv_mass;
        if (p > 0.0f) entropy = entropy + p * logf(p);
    }
    return entropy;
}

// simd.rs:291-331 — (inverse mass, sum p ln p)
template <class RP>
__device__ inline void probability_row_stats(RP row, int dim, bool vec, float* inv_mass_out,
                                             float* entropy_out) {
    double sum = 0;
    for (int i = 0; i < dim; ++i) {
        float v = ldrow(row + i);
        if (invalid_mass_value(v)) {
            *inv_mass_out = __int_as_float(0x7fc00000);
            *entropy_out = INFINITY;
            return;
        }
        sum = sum + (double)v;
    }
    if (sum == 0.0) {
        *inv_mass_out = 0.0f;
        *entropy_out = 0.0f;
        return;
    }
    float inv_mass = (float)(1.0 / sum);
    if (!isfinite(inv_mass) || inv_mass == 0.0f) {
        double e = 0;
        for (int i = 0; i < dim; ++i) {
            float v = ldrow(row + i);
            if (v > 0.0f) {
                double p = (double)v / sum;
                e = e + p * log(p);
            }
        }
        *inv_mass_out = inv_mass;
        *entropy_out = (float)e;
        return;
    }
    *inv_mass_out = inv_mass;
    *entropy_out = probability_entropy_avx(row, dim, vec, inv_mass);
}

// simd.rs:498-536
template <bool QG, class CP>
__device__ float jensen_shannon_normalized_query(const float* __restrict__ nq, CP cand, int dim,
                                                 bool vec, float cand_inv_mass) {
    float distance = jensen_shannon_avx<QG>(nq, cand, dim, vec, 1.0f, cand_inv_mass);
    if (distance * distance <= kJsStableDivergence) return jensen_shannon_distance<QG>(nq, cand, dim, vec);
    return distance;
}

// Σ s·ln(s), s = p + c·inv_c (simd.rs:2330-2345, :2376-2402)
template <bool QG, class CP>
__device__ float js_mixture_term(const float* __restrict__ nq, CP cand, int dim, bool vec,
                                 float cand_inv_mass) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int chunks = dim >> 3;
    for (int j = 0; j < chunks; ++j) {
        Vec8 p = load8<QG>(nq + 8 * j, vec), cv = load8<true>(cand + 8 * j, vec);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float q = cv.v[i] * cand_inv_mass;
            float sum = p.v[i] + q;
            float log_sum = fast_ln<true>(max_ps(sum, kMinPositive));
            acc[i] = fmaf(sum, log_sum, acc[i]);
        }
    }
    float mix = lane_sum8(acc);
    for (int i = chunks * 8; i < dim; ++i) {
        float s = (QG ? __ldg(nq + i) : nq[i]) + ldrow(cand + i) * cand_inv_mass;
        if (s > 0.0f) mix = mix + s * logf(s);
    }
    return mix;
}

// simd.rs:337-389 + :2316-2354 — entropy-form distance
template <bool QG, class CP>
__device__ float jensen_shannon_precomputed(const float* __restrict__ nq, CP cand, int dim,
                                            bool vec, float query_entropy, float cand_inv_mass, float cand_entropy) {
    if (cand_inv_mass == 0.0f) return sqrtf(kLn2);
    if (!isfinite(cand_entropy)) return INFINITY;
    if (!isfinite(cand_inv_mass)) return jensen_shannon_distance<QG>(nq, cand, dim, vec);
    float mix = js_mixture_term<QG>(nq, cand, dim, vec, cand_inv_mass);
    float divergence = fmaxf(kLn2 + 0.5f * (query_entropy + cand_entropy - mix), 0.0f);
    if (divergence <= kJsStableDivergence) return jensen_shannon_normalized_query<QG>(nq, cand, dim, vec, cand_inv_mass);
    return sqrtf(divergence);
}

// simd.rs:418-496 + :2356-2423 — squared distance used for ranking
template <bool QG, class CP>
__device__ float jensen_shannon_precomputed_divergence(const float* __restrict__ nq, CP cand,
                                                       int dim, bool vec, float query_entropy, float inv_mass,
                                                       float entropy) {
    if (inv_mass <= 0.0f || !isfinite(inv_mass) || !isfinite(entropy)) {
        float d = jensen_shannon_precomputed<QG>(nq, cand, dim, vec, query_entropy, inv_mass, entropy);
        return d * d;
    }
    float mix = js_mixture_term<QG>(nq, cand, dim, vec, inv_mass);
    float divergence = fmaxf(kLn2 + 0.5f * (query_entropy + entropy - mix), 0.0f);
    if (divergence <= kJsStableDivergence) {
        float d = jensen_shannon_normalized_query<QG>(nq, cand, dim, vec, inv_mass);
        return d * d;
    }
    return divergence;
}

// ---- compute_distance_f32 dispatch (src/distance/mod.rs:193-213) ------------------------------------------------------
// IP here is the single-row (two-accumulator) kernel, as in the reference.
template <bool QG, class CP>
__device__ float compute_distance(int metric, const float* __restrict__ q, CP c, int dim,
                                  bool vec) {
    switch (metric) {
        case LB_IP: return ip_single_order<QG>(q, c, dim, vec);
        case LB_L2: return l2_squared<QG>(q, c, dim, vec);
        case LB_COSINE: return cosine_distance<QG>(q, c, dim, vec);
        case LB_HAMMING: return hamming_f32<QG>(q, c, dim);
        case LB_JACCARD:
        case LB_TANIMOTO: return jaccard_f32<QG>(q, c, dim);
        case LB_MANHATTAN: return manhattan<QG>(q, c, dim, vec);
        case LB_HAVERSINE: return haversine_meters<QG>(q, c, dim);
        case LB_CORRELATION: return correlation_distance<QG>(q, c, dim, vec);
        case LB_HELLINGER: return hellinger_distance<QG>(q, c, dim, vec);
        case LB_WASSERSTEIN: return wasserstein_1d<QG>(q, c, dim, vec);
        case LB_DICE: return dice_f32<QG>(q, c, dim);
        case LB_JENSEN_SHANNON: return jensen_shannon_distance<QG>(q, c, dim, vec);
        case LB_CHEBYSHEV: return chebyshev<QG>(q, c, dim, vec);
        case LB_CANBERRA: return canberra<QG>(q, c, dim, vec);
        case LB_BRAY_CURTIS: return bray_curtis<QG>(q, c, dim, vec);
    }
    return __int_as_float(0x7fc00000);
}

// ---- f32 query x binary16 row: the scalar kernels of the F16 storage dtype (simd.rs:805-1092) -------------------------
// Rows of a float16 collection hold exactly binary16-representable values, so `c` may be the decoded row
// (half::f16::to_f32 is exact).  Every sum is the reference's sequential scalar loop: element order, products and
// sums rounded separately.  Loads are 8 values at a time, the arithmetic one element at a time.
template <bool QG, class CP>
__device__ float inner_product_f16order(const float* __restrict__ q, CP c, int dim, bool vec) {
    float sum = 0.0f;
    scalar_order_foreach<QG>(q, c, dim, vec, [&](float a, float b) { sum = sum + a * b; });
    return sum;
}
template <bool QG, class CP>
__device__ float l2_squared_f16order(const float* __restrict__ q, CP c, int dim, bool vec) {
    float sum = 0.0f;
    scalar_order_foreach<QG>(q, c, dim, vec, [&](float a, float b) {
        float diff = a - b;
        sum = sum + diff * diff;
    });
    return sum;
}
template <bool QG, class CP>
__device__ float cosine_distance_f16order(const float* __restrict__ q, CP c, int dim, bool vec) {
    float dot = 0.0f, nq = 0.0f, nc = 0.0f;
    scalar_order_foreach<QG>(q, c, dim, vec, [&](float a, float b) {
        dot = dot + a * b;
        nq = nq + a * a;
        nc = nc + b * b;
    });
    if (nq == 0.0f || nc == 0.0f) return 1.0f;
    return 1.0f - dot / (sqrtf(nq) * sqrtf(nc));
}
template <bool QG, class CP>
__device__ float manhattan_f16order(const float* __restrict__ q, CP c, int dim, bool vec) {
    float sum = 0.0f;
    scalar_order_foreach<QG>(q, c, dim, vec, [&](float a, float b) { sum = sum + fabsf(a - b); });
    return sum;
}
template <bool QG, class CP>
__device__ float chebyshev_f16order(const float* __restrict__ q, CP c, int dim, bool vec) {
    float m = 0.0f;
    scalar_order_foreach<QG>(q, c, dim, vec, [&](float a, float b) { m = rust_max(m, fabsf(a - b)); });
    return m;
}
template <bool QG, class CP>
__device__ float canberra_f16order(const float* __restrict__ q, CP c, int dim, bool vec) {
    float sum = 0.0f;
    scalar_order_foreach<QG>(q, c, dim, vec, [&](float a, float b) {
        float den = fabsf(a) + fabsf(b);
        sum = sum + (den == 0.0f ? 0.0f : fabsf(a - b) / den);
    });
    return sum;
}
template <bool QG, class CP>
__device__ float bray_curtis_f16order(const float* __restrict__ q, CP c, int dim, bool vec) {
    float num = 0.0f, den = 0.0f;
    scalar_order_foreach<QG>(q, c, dim, vec, [&](float a, float b) {
        num = num + fabsf(a - b);
        den = den + fabsf(a + b);
    });
    if (den == 0.0f) return num == 0.0f ? 0.0f : INFINITY;
    return num / den;
}
template <bool QG, class CP>
__device__ float jensen_shannon_f16order(const float* __restrict__ q, CP c, int dim, bool vec) {
    double sa = 0, sb = 0;
    bool bad = false;
    scalar_order_foreach<QG>(q, c, dim, vec, [&](float a, float b) {
        bad = bad || invalid_mass_value(a) || invalid_mass_value(b);
        sa = sa + (double)a;
        sb = sb + (double)b;
    });
    if (bad) return INFINITY;
    if (sa == 0.0 || sb == 0.0) return sa == sb ? 0.0f : sqrtf(kLn2);
    double divergence = 0;
    scalar_order_foreach<QG>(q, c, dim, vec, [&](float a, float b) {
        double p = (double)a / sa, qq = (double)b / sb, m = 0.5 * (p + qq);
        if (p > 0.0) divergence = divergence + 0.5 * p * log(p / m);
        if (qq > 0.0) divergence = divergence + 0.5 * qq * log(qq / m);
    });
    return (float)sqrt(fmax(divergence, 0.0));
}
template <bool QG, class CP>
__device__ float wasserstein_1d_f16order(const float* __restrict__ q, CP c, int dim, bool vec) {
    return wasserstein_1d_impl<QG, true>(q, c, dim, vec);
}

// compute_distance_f16 dispatch (src/distance/mod.rs:217-237).  Haversine, correlation and Hellinger repeat their f32
// formulas on the decoded row; the binary metrics count thresholded bits element by element, as the f32 ones do.
template <bool QG, class CP>
__device__ float compute_distance_f16order(int metric, const float* __restrict__ q, CP c, int dim,
                                           bool vec) {
    switch (metric) {
        case LB_IP: return inner_product_f16order<QG>(q, c, dim, vec);
        case LB_L2: return l2_squared_f16order<QG>(q, c, dim, vec);
        case LB_COSINE: return cosine_distance_f16order<QG>(q, c, dim, vec);
        case LB_HAMMING: return hamming_f32<QG>(q, c, dim);
        case LB_JACCARD:
        case LB_TANIMOTO: return jaccard_f32<QG>(q, c, dim);
        case LB_MANHATTAN: return manhattan_f16order<QG>(q, c, dim, vec);
        case LB_HAVERSINE: return haversine_meters<QG>(q, c, dim);
        case LB_CORRELATION: return correlation_distance<QG>(q, c, dim, vec);
        case LB_HELLINGER: return hellinger_distance<QG>(q, c, dim, vec);
        case LB_WASSERSTEIN: return wasserstein_1d_f16order<QG>(q, c, dim, vec);
        case LB_DICE: return dice_f32<QG>(q, c, dim);
        case LB_JENSEN_SHANNON: return jensen_shannon_f16order<QG>(q, c, dim, vec);
        case LB_CHEBYSHEV: return chebyshev_f16order<QG>(q, c, dim, vec);
        case LB_CANBERRA: return canberra_f16order<QG>(q, c, dim, vec);
        case LB_BRAY_CURTIS: return bray_curtis_f16order<QG>(q, c, dim, vec);
    }
    return __int_as_float(0x7fc00000);
}

// ---- packed one-bit rows (simd.rs:765-801) ------------------------------------------------------------------------------
__device__ __forceinline__ float packed_finish(int metric, uint32_t x /*xor or inter*/, uint32_t y /*union or count*/) {
    if (metric == LB_HAMMING) return (float)x;
    if (metric == LB_DICE) return y == 0 ? 0.0f : 1.0f - (float)(2 * x) / (float)y;
    return y == 0 ? 0.0f : 1.0f - (float)x / (float)y;  // Jaccard / Tanimoto
}

}  // namespace lb
