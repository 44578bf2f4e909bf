// lynse_oracle.cpp — CPU restatement of LynseDB's batched distance + top-k path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product path (lynsedb_b200/) may
// import, link or execute this file.  Allowed callers: tests/,
// __graft_entry__.smoke(), and bench.py's cpu_baseline / --impl reference legs.
//
// What it is: a function-by-function restatement, in C++ with the same
// AVX2+FMA intrinsics in the same order, of the reference's Rust CPU kernels
// (all citations relative to the reference tree):
//   src/distance/simd.rs        per-pair kernels (x86_64 AVX2+FMA branches)
//   src/distance/mod.rs         compute_distance_f32, top_k_search, quickselect_k
//   src/storage/flat_mmap.rs    fused scan + sorted-array top-k + chunk merge,
//                               packed-binary search, Jensen-Shannon cached scan
//   src/storage/vector_store.rs segment fan-out + (score,row) merge
//   src/index/ivf.rs            IVF probe + list scan (given centroids/lists)
//
// Why a restatement: the reference is Rust (pyo3 / maturin); this image has no
// cargo / rustc, so the reference cannot be compiled or imported here
// (oracle/_ref is therefore absent).  Parity is PINNED instead against every
// known-answer test the reference holds for this path — see
// tests/test_oracle_golden.py, which replays them as data.
//
// Build: g++ -O2 -mavx2 -mfma -mpopcnt -ffp-contract=off -fopenmp (oracle/Makefile).
// -ffp-contract=off keeps scalar tails unfused, as rustc does.
//
// Known, documented deviations from a real Rust run:
//   * `sort_unstable_by` on the fill-phase top-k buffer: Rust's std uses a
//     stable insertion sort for len <= 20; for larger k the order among equal
//     distances is unspecified.  We use a stable sort for every k.
//   * rayon's chunking depends on the live thread count; here it is the
//     explicit `n_threads` argument (chunk = max(n / n_threads, 512)).

#include <immintrin.h>
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace {

enum Metric : int {
    IP = 0,
    L2 = 1,
    COSINE = 2,
    HAMMING = 3,
    JACCARD = 4,
    MANHATTAN = 5,
    HAVERSINE = 6,
    CORRELATION = 7,
    HELLINGER = 8,
    WASSERSTEIN = 9,
    DICE = 10,
    TANIMOTO = 11,
    JENSEN_SHANNON = 12,
    CHEBYSHEV = 13,
    CANBERRA = 14,
    BRAY_CURTIS = 15,
    METRIC_COUNT = 16,
};

constexpr float kInf = std::numeric_limits<float>::infinity();
constexpr float kLn2 = 0.693147180559945309417232121458176568f;  // f32::consts::LN_2
constexpr float kFrac1Sqrt2 = 0.707106781186547524400844362104849039f;
constexpr float kMinPositive = 1.17549435e-38f;  // f32::MIN_POSITIVE
constexpr float kJsStableDivergence = 1e-5f;     // simd.rs:8

inline bool is_ascending(int metric) { return metric != IP; }  // distance/mod.rs:111-116
inline bool is_binary(int metric) {                             // distance/mod.rs:161-166
    return metric == HAMMING || metric == JACCARD || metric == DICE || metric == TANIMOTO;
}

// ---- simd.rs:1427-1436 — horizontal sum used by IP / L2 / cosine ----------
inline float hsum8(__m256 acc) {
    __m128 hi = _mm256_extractf128_ps(acc, 1);
    __m128 lo = _mm256_castps256_ps128(acc);
    __m128 sum128 = _mm_add_ps(lo, hi);
    __m128 shuf = _mm_movehdup_ps(sum128);
    __m128 sums = _mm_add_ps(sum128, shuf);
    __m128 shuf2 = _mm_movehl_ps(sums, sums);
    __m128 result = _mm_add_ss(sums, shuf2);
    return _mm_cvtss_f32(result);
}

// `lanes.into_iter().sum()` — sequential left-to-right (simd.rs:2151-2153)
inline float lane_sum(__m256 acc) {
    float lanes[8];
    _mm256_storeu_ps(lanes, acc);
    float s = 0.0f;
    for (int i = 0; i < 8; ++i) s += lanes[i];
    return s;
}

// ---- simd.rs:1341-1396 — single dot product, two accumulators ---------------
float inner_product_single(const float* a, const float* b, size_t n) {
    size_t chunks = n / 8, remainder = n % 8;
    __m256 acc0 = _mm256_setzero_ps(), acc1 = _mm256_setzero_ps();
    size_t double_chunks = chunks / 2, single_remaining = chunks % 2;
    for (size_t i = 0; i < double_chunks; ++i) {
        size_t base = i * 16;
        acc0 = _mm256_fmadd_ps(_mm256_loadu_ps(a + base), _mm256_loadu_ps(b + base), acc0);
        acc1 = _mm256_fmadd_ps(_mm256_loadu_ps(a + base + 8), _mm256_loadu_ps(b + base + 8), acc1);
    }
    if (single_remaining > 0) {
        size_t base = double_chunks * 16;
        acc0 = _mm256_fmadd_ps(_mm256_loadu_ps(a + base), _mm256_loadu_ps(b + base), acc0);
    }
    acc0 = _mm256_add_ps(acc0, acc1);
    float sum = hsum8(acc0);
    size_t base = chunks * 8;
    for (size_t i = 0; i < remainder; ++i) sum += a[base + i] * b[base + i];
    return sum;
}

// ---- simd.rs:1450-1525 — eight dot products sharing the query loads --------
void inner_product_batch8(const float* q, const float* const v[8], size_t n, float out[8]) {
    size_t chunks = n / 8, remainder = n % 8;
    __m256 acc[8];
    for (int r = 0; r < 8; ++r) acc[r] = _mm256_setzero_ps();
    for (size_t i = 0; i < chunks; ++i) {
        size_t base = i * 8;
        __m256 qv = _mm256_loadu_ps(q + base);
        for (int r = 0; r < 8; ++r)
            acc[r] = _mm256_fmadd_ps(qv, _mm256_loadu_ps(v[r] + base), acc[r]);
    }
    for (int r = 0; r < 8; ++r) out[r] = hsum8(acc[r]);
    size_t base = chunks * 8;
    for (size_t i = 0; i < remainder; ++i) {
        float qq = q[base + i];
        for (int r = 0; r < 8; ++r) out[r] += qq * v[r][base + i];
    }
}

// ---- simd.rs:1527-1581 ------------------------------------------------------
float l2_squared(const float* a, const float* b, size_t n) {
    size_t chunks = n / 8, remainder = n % 8;
    __m256 acc0 = _mm256_setzero_ps(), acc1 = _mm256_setzero_ps();
    size_t double_chunks = chunks / 2, single_remaining = chunks % 2;
    for (size_t i = 0; i < double_chunks; ++i) {
        size_t base = i * 16;
        __m256 d0 = _mm256_sub_ps(_mm256_loadu_ps(a + base), _mm256_loadu_ps(b + base));
        acc0 = _mm256_fmadd_ps(d0, d0, acc0);
        __m256 d1 = _mm256_sub_ps(_mm256_loadu_ps(a + base + 8), _mm256_loadu_ps(b + base + 8));
        acc1 = _mm256_fmadd_ps(d1, d1, acc1);
    }
    if (single_remaining > 0) {
        size_t base = double_chunks * 16;
        __m256 d = _mm256_sub_ps(_mm256_loadu_ps(a + base), _mm256_loadu_ps(b + base));
        acc0 = _mm256_fmadd_ps(d, d, acc0);
    }
    acc0 = _mm256_add_ps(acc0, acc1);
    float sum = hsum8(acc0);
    size_t base = chunks * 8;
    for (size_t i = 0; i < remainder; ++i) {
        float diff = a[base + i] - b[base + i];
        sum += diff * diff;
    }
    return sum;
}

// ---- simd.rs:1583-1636 ------------------------------------------------------
float cosine_distance(const float* a, const float* b, size_t n) {
    size_t chunks = n / 8, remainder = n % 8;
    __m256 dot_acc = _mm256_setzero_ps(), na_acc = _mm256_setzero_ps(), nb_acc = _mm256_setzero_ps();
    for (size_t i = 0; i < chunks; ++i) {
        size_t base = i * 8;
        __m256 va = _mm256_loadu_ps(a + base), vb = _mm256_loadu_ps(b + base);
        dot_acc = _mm256_fmadd_ps(va, vb, dot_acc);
        na_acc = _mm256_fmadd_ps(va, va, na_acc);
        nb_acc = _mm256_fmadd_ps(vb, vb, nb_acc);
    }
    float dot = hsum8(dot_acc), norm_a = hsum8(na_acc), norm_b = hsum8(nb_acc);
    size_t base = chunks * 8;
    for (size_t i = 0; i < remainder; ++i) {
        dot += a[base + i] * b[base + i];
        norm_a += a[base + i] * a[base + i];
        norm_b += b[base + i] * b[base + i];
    }
    float denom = std::sqrt(norm_a * norm_b);
    if (denom < 1e-30f) return 1.0f;
    return 1.0f - dot / denom;
}

// ---- simd.rs:175-209, 718-736 — thresholded binary metrics on f32 ----------
float hamming_f32(const float* a, const float* b, size_t n) {
    uint32_t count = 0;
    for (size_t i = 0; i < n; ++i) count += ((a[i] > 0.5f) != (b[i] > 0.5f));
    return (float)count;
}
float jaccard_f32(const float* a, const float* b, size_t n) {
    uint32_t inter = 0, uni = 0;
    for (size_t i = 0; i < n; ++i) {
        bool ab = a[i] > 0.5f, bb = b[i] > 0.5f;
        if (ab || bb) {
            ++uni;
            if (ab && bb) ++inter;
        }
    }
    return uni == 0 ? 0.0f : 1.0f - ((float)inter / (float)uni);
}
float dice_f32(const float* a, const float* b, size_t n) {
    uint32_t inter = 0, ca = 0, cb = 0;
    for (size_t i = 0; i < n; ++i) {
        bool ab = a[i] > 0.5f, bb = b[i] > 0.5f;
        ca += ab;
        cb += bb;
        inter += (ab && bb);
    }
    uint32_t total = ca + cb;
    return total == 0 ? 0.0f : 1.0f - (float)(2 * inter) / (float)total;
}

// ---- simd.rs:2134-2158 ------------------------------------------------------
float manhattan(const float* a, const float* b, size_t n) {
    size_t chunks = n / 8;
    __m256 acc = _mm256_setzero_ps();
    const __m256 sign_mask = _mm256_set1_ps(-0.0f);
    for (size_t i = 0; i < chunks; ++i) {
        __m256 diff = _mm256_sub_ps(_mm256_loadu_ps(a + i * 8), _mm256_loadu_ps(b + i * 8));
        acc = _mm256_add_ps(acc, _mm256_andnot_ps(sign_mask, diff));
    }
    float sum = lane_sum(acc);
    for (size_t i = chunks * 8; i < n; ++i) sum += std::fabs(a[i] - b[i]);
    return sum;
}

// Rust's f32::max: NaN-ignoring maximum.
inline float rust_max(float x, float y) {
    if (x != x) return y;
    if (y != y) return x;
    return x > y ? x : y;
}

// ---- simd.rs:2715-2737 ------------------------------------------------------
float chebyshev(const float* a, const float* b, size_t n) {
    size_t chunks = n / 8;
    __m256 acc = _mm256_setzero_ps();
    const __m256 sign_mask = _mm256_set1_ps(-0.0f);
    for (size_t i = 0; i < chunks; ++i) {
        __m256 diff = _mm256_andnot_ps(
            sign_mask, _mm256_sub_ps(_mm256_loadu_ps(a + i * 8), _mm256_loadu_ps(b + i * 8)));
        acc = _mm256_max_ps(acc, diff);
    }
    float lanes[8];
    _mm256_storeu_ps(lanes, acc);
    float maximum = 0.0f;
    for (int i = 0; i < 8; ++i) maximum = rust_max(maximum, lanes[i]);
    for (size_t i = chunks * 8; i < n; ++i) maximum = rust_max(maximum, std::fabs(a[i] - b[i]));
    return maximum;
}

// ---- simd.rs:2762-2793 ------------------------------------------------------
float canberra(const float* a, const float* b, size_t n) {
    size_t chunks = n / 8;
    __m256 acc = _mm256_setzero_ps();
    const __m256 zero = _mm256_setzero_ps();
    const __m256 sign_mask = _mm256_set1_ps(-0.0f);
    for (size_t i = 0; i < chunks; ++i) {
        __m256 va = _mm256_loadu_ps(a + i * 8), vb = _mm256_loadu_ps(b + i * 8);
        __m256 numerator = _mm256_andnot_ps(sign_mask, _mm256_sub_ps(va, vb));
        __m256 denominator =
            _mm256_add_ps(_mm256_andnot_ps(sign_mask, va), _mm256_andnot_ps(sign_mask, vb));
        __m256 nonzero = _mm256_cmp_ps(denominator, zero, _CMP_NEQ_OQ);
        __m256 quotient = _mm256_div_ps(numerator, denominator);
        acc = _mm256_add_ps(acc, _mm256_and_ps(nonzero, quotient));
    }
    float sum = lane_sum(acc);
    for (size_t i = chunks * 8; i < n; ++i) {
        float denominator = std::fabs(a[i]) + std::fabs(b[i]);
        if (denominator != 0.0f) sum += std::fabs(a[i] - b[i]) / denominator;
    }
    return sum;
}

// ---- simd.rs:2824-2865 ------------------------------------------------------
float bray_curtis(const float* a, const float* b, size_t n) {
    size_t chunks = n / 8;
    const __m256 sign_mask = _mm256_set1_ps(-0.0f);
    __m256 num_acc = _mm256_setzero_ps(), den_acc = _mm256_setzero_ps();
    for (size_t i = 0; i < chunks; ++i) {
        __m256 va = _mm256_loadu_ps(a + i * 8), vb = _mm256_loadu_ps(b + i * 8);
        num_acc = _mm256_add_ps(num_acc, _mm256_andnot_ps(sign_mask, _mm256_sub_ps(va, vb)));
        den_acc = _mm256_add_ps(den_acc, _mm256_andnot_ps(sign_mask, _mm256_add_ps(va, vb)));
    }
    float numerator = lane_sum(num_acc), denominator = lane_sum(den_acc);
    for (size_t i = chunks * 8; i < n; ++i) {
        numerator += std::fabs(a[i] - b[i]);
        denominator += std::fabs(a[i] + b[i]);
    }
    if (denominator == 0.0f) return numerator == 0.0f ? 0.0f : kInf;
    return numerator / denominator;
}

// ---- simd.rs:603-628 --------------------------------------------------------
inline double clamp_f64(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

float haversine_meters(const float* a, const float* b, size_t n) {
    if (n != 2) return kInf;
    const double R = 6371008.8;
    const double k = 3.14159265358979323846264338327950288 / 180.0;
    double lon1 = (double)a[0] * k, lat1 = (double)a[1] * k;
    double lon2 = (double)b[0] * k, lat2 = (double)b[1] * k;
    if (!std::isfinite(lon1) || !std::isfinite(lat1) || !std::isfinite(lon2) || !std::isfinite(lat2) ||
        std::fabs(a[1]) > 90.0f || std::fabs(b[1]) > 90.0f)
        return kInf;
    double dlat = lat2 - lat1, dlon = lon2 - lon1;
    double sin_lat = std::sin(dlat * 0.5), sin_lon = std::sin(dlon * 0.5);
    double h = clamp_f64(sin_lat * sin_lat + std::cos(lat1) * std::cos(lat2) * sin_lon * sin_lon, 0.0, 1.0);
    return (float)(2.0 * R * std::asin(std::sqrt(h)));
}

// ---- simd.rs:632-661 --------------------------------------------------------
float correlation_distance(const float* a, const float* b, size_t len) {
    if (len == 0) return 0.0f;
    double n = (double)len, sa = 0, sb = 0, saa = 0, sbb = 0, sab = 0;
    for (size_t i = 0; i < len; ++i) {
        double av = a[i], bv = b[i];
        sa += av;
        sb += bv;
        saa += av * av;
        sbb += bv * bv;
        sab += av * bv;
    }
    double var_a = std::max(saa - sa * sa / n, 0.0);
    double var_b = std::max(sbb - sb * sb / n, 0.0);
    double denom = std::sqrt(var_a * var_b);
    if (denom <= 2.2204460492503131e-16) {
        bool same = true;
        for (size_t i = 0; i < len; ++i) same = same && (a[i] == b[i]);
        return same ? 0.0f : 1.0f;
    }
    double cov = sab - sa * sb / n;
    return (float)(1.0 - clamp_f64(cov / denom, -1.0, 1.0));
}

inline bool invalid_mass_value(float v) { return !std::isfinite(v) || v < 0.0f; }

// ---- simd.rs:665-684 --------------------------------------------------------
float hellinger_distance(const float* a, const float* b, size_t n) {
    double sa = 0, sb = 0, coef = 0;
    for (size_t i = 0; i < n; ++i) {
        if (invalid_mass_value(a[i]) || invalid_mass_value(b[i])) return kInf;
        sa += (double)a[i];
        sb += (double)b[i];
        coef += std::sqrt((double)a[i] * (double)b[i]);
    }
    if (sa == 0.0 || sb == 0.0) return sa == sb ? 0.0f : 1.0f;
    double c = coef / std::sqrt(sa * sb);
    return (float)std::sqrt(1.0 - clamp_f64(c, 0.0, 1.0));
}

// ---- simd.rs:688-714 --------------------------------------------------------
float wasserstein_1d(const float* a, const float* b, size_t n) {
    double sa = 0, sb = 0;
    for (size_t i = 0; i < n; ++i) {
        if (invalid_mass_value(a[i]) || invalid_mass_value(b[i])) return kInf;
        sa += (double)a[i];
        sb += (double)b[i];
    }
    if (sa == 0.0 || sb == 0.0) return sa == sb ? 0.0f : kInf;
    double inv_a = 1.0 / sa, inv_b = 1.0 / sb, cdf = 0, dist = 0;
    for (size_t i = 0; i + 1 < n; ++i) {
        cdf += (double)a[i] * inv_a - (double)b[i] * inv_b;
        dist += std::fabs(cdf);
    }
    return (float)dist;
}

// ---- simd.rs:2164-2205 / 2211-2247 — Cephes-style 8-lane ln -----------------
template <bool COMPACT>
inline __m256 fast_ln(__m256 x) {
    __m256i bits = _mm256_castps_si256(x);
    __m256i exponent_bits = _mm256_srli_epi32(bits, 23);
    __m256i mantissa_bits = _mm256_or_si256(_mm256_and_si256(bits, _mm256_set1_epi32(0x007fffff)),
                                            _mm256_set1_epi32(0x3f000000));
    x = _mm256_castsi256_ps(mantissa_bits);
    __m256 exponent = _mm256_cvtepi32_ps(_mm256_sub_epi32(exponent_bits, _mm256_set1_epi32(0x7f)));
    exponent = _mm256_add_ps(exponent, _mm256_set1_ps(1.0f));
    __m256 mask = _mm256_cmp_ps(x, _mm256_set1_ps(kFrac1Sqrt2), _CMP_LT_OQ);
    __m256 tmp = _mm256_and_ps(x, mask);
    x = _mm256_sub_ps(x, _mm256_set1_ps(1.0f));
    exponent = _mm256_sub_ps(exponent, _mm256_and_ps(_mm256_set1_ps(1.0f), mask));
    x = _mm256_add_ps(x, tmp);
    __m256 z = _mm256_mul_ps(x, x);
    __m256 y;
    if (!COMPACT) {
        y = _mm256_set1_ps(7.0376836E-2f);
        y = _mm256_fmadd_ps(y, x, _mm256_set1_ps(-1.151461E-1f));
        y = _mm256_fmadd_ps(y, x, _mm256_set1_ps(1.1676998E-1f));
        y = _mm256_fmadd_ps(y, x, _mm256_set1_ps(-1.2420141E-1f));
    } else {
        y = _mm256_set1_ps(-1.2420141E-1f);
    }
    y = _mm256_fmadd_ps(y, x, _mm256_set1_ps(1.4249323E-1f));
    y = _mm256_fmadd_ps(y, x, _mm256_set1_ps(-1.6668057E-1f));
    y = _mm256_fmadd_ps(y, x, _mm256_set1_ps(2.0000714E-1f));
    y = _mm256_fmadd_ps(y, x, _mm256_set1_ps(-2.4999994E-1f));
    y = _mm256_fmadd_ps(y, x, _mm256_set1_ps(3.333333E-1f));
    y = _mm256_mul_ps(_mm256_mul_ps(y, x), z);
    y = _mm256_fmadd_ps(exponent, _mm256_set1_ps(-2.1219444E-4f), y);
    y = _mm256_fnmadd_ps(z, _mm256_set1_ps(0.5f), y);
    x = _mm256_add_ps(x, y);
    return _mm256_fmadd_ps(exponent, _mm256_set1_ps(0.6933594f), x);
}

// ---- simd.rs:2249-2286 ------------------------------------------------------
float jensen_shannon_avx(const float* a, const float* b, size_t n, float inv_a_s, float inv_b_s) {
    size_t chunks = n / 8;
    __m256 acc = _mm256_setzero_ps();
    __m256 inv_a = _mm256_set1_ps(inv_a_s), inv_b = _mm256_set1_ps(inv_b_s);
    __m256 half = _mm256_set1_ps(0.5f), min_normal = _mm256_set1_ps(kMinPositive);
    for (size_t i = 0; i < chunks; ++i) {
        __m256 p = _mm256_mul_ps(_mm256_loadu_ps(a + i * 8), inv_a);
        __m256 q = _mm256_mul_ps(_mm256_loadu_ps(b + i * 8), inv_b);
        __m256 m = _mm256_mul_ps(_mm256_add_ps(p, q), half);
        __m256 safe_m = _mm256_max_ps(m, min_normal);
        __m256 log_p = fast_ln<false>(_mm256_div_ps(_mm256_max_ps(p, min_normal), safe_m));
        __m256 log_q = fast_ln<false>(_mm256_div_ps(_mm256_max_ps(q, min_normal), safe_m));
        __m256 terms = _mm256_add_ps(_mm256_mul_ps(p, log_p), _mm256_mul_ps(q, log_q));
        acc = _mm256_fmadd_ps(terms, half, acc);
    }
    float divergence = lane_sum(acc);
    for (size_t i = chunks * 8; i < n; ++i) {
        float p = a[i] * inv_a_s, q = b[i] * inv_b_s;
        float m = 0.5f * (p + q);
        if (p > 0.0f) divergence += 0.5f * p * std::log(p / m);
        if (q > 0.0f) divergence += 0.5f * q * std::log(q / m);
    }
    return std::sqrt(std::max(divergence, 0.0f));
}

// ---- simd.rs:1161-1178 ------------------------------------------------------
float jensen_shannon_scalar_f64(const float* a, const float* b, size_t n, double sum_a, double sum_b) {
    double inv_a = 1.0 / sum_a, inv_b = 1.0 / sum_b, divergence = 0;
    for (size_t i = 0; i < n; ++i) {
        double p = (double)a[i] * inv_a, q = (double)b[i] * inv_b, m = 0.5 * (p + q);
        if (p > 0.0) divergence += 0.5 * p * std::log(p / m);
        if (q > 0.0) divergence += 0.5 * q * std::log(q / m);
    }
    return (float)std::sqrt(std::max(divergence, 0.0));
}

inline bool slices_equal(const float* a, const float* b, size_t n) {
    for (size_t i = 0; i < n; ++i)
        if (!(a[i] == b[i])) return false;
    return true;
}

// ---- simd.rs:235-284 (+ refine_small, 1118-1125) ----------------------------
float jensen_shannon_distance(const float* a, const float* b, size_t n) {
    double sum_a = 0, sum_b = 0;
    for (size_t i = 0; i < n; ++i) {
        if (invalid_mass_value(a[i]) || invalid_mass_value(b[i])) return kInf;
        sum_a += (double)a[i];
        sum_b += (double)b[i];
    }
    if (sum_a == 0.0 || sum_b == 0.0) return sum_a == sum_b ? 0.0f : std::sqrt(kLn2);
    float inv_a = (float)(1.0 / sum_a), inv_b = (float)(1.0 / sum_b);
    if (!std::isfinite(inv_a) || !std::isfinite(inv_b) || inv_a == 0.0f || inv_b == 0.0f)
        return jensen_shannon_scalar_f64(a, b, n, sum_a, sum_b);
    float distance = jensen_shannon_avx(a, b, n, inv_a, inv_b);
    if (distance * distance <= kJsStableDivergence && !slices_equal(a, b, n))
        return jensen_shannon_scalar_f64(a, b, n, sum_a, sum_b);
    return distance;
}

// ---- simd.rs:2288-2312 ------------------------------------------------------
float probability_entropy_avx(const float* row, size_t n, float inv_mass) {
    size_t chunks = n / 8;
    __m256 acc = _mm256_setzero_ps();
    __m256 inv = _mm256_set1_ps(inv_mass), min_normal = _mm256_set1_ps(kMinPositive);
    for (size_t i = 0; i < chunks; ++i) {
        __m256 p = _mm256_mul_ps(_mm256_loadu_ps(row + i * 8), inv);
        __m256 log_p = fast_ln<true>(_mm256_max_ps(p, min_normal));
        acc = _mm256_fmadd_ps(p, log_p, acc);
    }
    float entropy = lane_sum(acc);
    for (size_t i = chunks * 8; i < n; ++i) {
        float p = row[i] * inv_mass;
        if (p > 0.0f) entropy += p * std::log(p);
    }
    return entropy;
}

// ---- simd.rs:291-331 — (inverse mass, sum p ln p) ---------------------------
void probability_row_stats(const float* row, size_t n, float* inv_mass_out, float* entropy_out) {
    double sum = 0;
    for (size_t i = 0; i < n; ++i) {
        if (invalid_mass_value(row[i])) {
            *inv_mass_out = std::numeric_limits<float>::quiet_NaN();
            *entropy_out = kInf;
            return;
        }
        sum += (double)row[i];
    }
    if (sum == 0.0) {
        *inv_mass_out = 0.0f;
        *entropy_out = 0.0f;
        return;
    }
    float inv_mass = (float)(1.0 / sum);
    if (!std::isfinite(inv_mass) || inv_mass == 0.0f) {
        double e = 0;
        for (size_t i = 0; i < n; ++i)
            if (row[i] > 0.0f) {
                double p = (double)row[i] / sum;
                e += p * std::log(p);
            }
        *inv_mass_out = inv_mass;
        *entropy_out = (float)e;
        return;
    }
    *inv_mass_out = inv_mass;
    *entropy_out = probability_entropy_avx(row, n, inv_mass);
}

// ---- simd.rs:498-536 --------------------------------------------------------
float jensen_shannon_normalized_query(const float* nq, const float* cand, size_t n, float cand_inv_mass) {
    float distance = jensen_shannon_avx(nq, cand, n, 1.0f, cand_inv_mass);
    if (distance * distance <= kJsStableDivergence) return jensen_shannon_distance(nq, cand, n);
    return distance;
}

// Mixture term Σ s·ln(s), s = p + c·inv_c  (simd.rs:2330-2345 / 2376-2402; the
// batch-2 kernel performs the same per-row arithmetic as the single kernel).
float js_mixture_term(const float* nq, const float* cand, size_t n, float cand_inv_mass) {
    size_t chunks = n / 8;
    __m256 acc = _mm256_setzero_ps();
    __m256 inv = _mm256_set1_ps(cand_inv_mass), min_normal = _mm256_set1_ps(kMinPositive);
    for (size_t i = 0; i < chunks; ++i) {
        __m256 p = _mm256_loadu_ps(nq + i * 8);
        __m256 q = _mm256_mul_ps(_mm256_loadu_ps(cand + i * 8), inv);
        __m256 sum = _mm256_add_ps(p, q);
        __m256 log_sum = fast_ln<true>(_mm256_max_ps(sum, min_normal));
        acc = _mm256_fmadd_ps(sum, log_sum, acc);
    }
    float mix = lane_sum(acc);
    for (size_t i = chunks * 8; i < n; ++i) {
        float s = nq[i] + cand[i] * cand_inv_mass;
        if (s > 0.0f) mix += s * std::log(s);
    }
    return mix;
}

// ---- simd.rs:337-389 + 2316-2354 — entropy-form distance --------------------
float jensen_shannon_precomputed(const float* nq, const float* cand, size_t n, float query_entropy,
                                 float cand_inv_mass, float cand_entropy) {
    if (cand_inv_mass == 0.0f) return std::sqrt(kLn2);
    if (!std::isfinite(cand_entropy)) return kInf;
    if (!std::isfinite(cand_inv_mass)) return jensen_shannon_distance(nq, cand, n);
    float mix = js_mixture_term(nq, cand, n, cand_inv_mass);
    float divergence = std::max(kLn2 + 0.5f * (query_entropy + cand_entropy - mix), 0.0f);
    if (divergence <= kJsStableDivergence) return jensen_shannon_normalized_query(nq, cand, n, cand_inv_mass);
    return std::sqrt(divergence);
}

// ---- simd.rs:418-496 + 2356-2423 — squared distance for ranking -------------
float jensen_shannon_precomputed_divergence(const float* nq, const float* cand, size_t n, float query_entropy,
                                            float inv_mass, float entropy) {
    if (inv_mass <= 0.0f || !std::isfinite(inv_mass) || !std::isfinite(entropy)) {
        float d = jensen_shannon_precomputed(nq, cand, n, query_entropy, inv_mass, entropy);
        return d * d;
    }
    float mix = js_mixture_term(nq, cand, n, inv_mass);
    float divergence = std::max(kLn2 + 0.5f * (query_entropy + entropy - mix), 0.0f);
    if (divergence <= kJsStableDivergence) {
        float d = jensen_shannon_normalized_query(nq, cand, n, inv_mass);
        return d * d;
    }
    return divergence;
}

// ---- distance/mod.rs:193-213 ------------------------------------------------
float compute_distance(const float* a, const float* b, size_t n, int metric) {
    switch (metric) {
        case IP: return inner_product_single(a, b, n);
        case L2: return l2_squared(a, b, n);
        case COSINE: return cosine_distance(a, b, n);
        case HAMMING: return hamming_f32(a, b, n);
        case JACCARD:
        case TANIMOTO: return jaccard_f32(a, b, n);
        case MANHATTAN: return manhattan(a, b, n);
        case HAVERSINE: return haversine_meters(a, b, n);
        case CORRELATION: return correlation_distance(a, b, n);
        case HELLINGER: return hellinger_distance(a, b, n);
        case WASSERSTEIN: return wasserstein_1d(a, b, n);
        case DICE: return dice_f32(a, b, n);
        case JENSEN_SHANNON: return jensen_shannon_distance(a, b, n);
        case CHEBYSHEV: return chebyshev(a, b, n);
        case CANBERRA: return canberra(a, b, n);
        case BRAY_CURTIS: return bray_curtis(a, b, n);
    }
    return std::numeric_limits<float>::quiet_NaN();
}

// ---- simd.rs:805-1092 — f32 query x binary16 candidate row, the scalar kernels of the F16 storage dtype ------
// `c` holds the candidate's DECODED values (half::f16::to_f32 is exact, so decoding first changes nothing); every
// accumulation below is the reference's sequential scalar loop (no SIMD lanes, products and sums unfused).
float inner_product_f16(const float* q, const float* c, size_t n) {
    float sum = 0.0f;
    for (size_t i = 0; i < n; ++i) sum += q[i] * c[i];
    return sum;
}
float l2_squared_f16(const float* q, const float* c, size_t n) {
    float sum = 0.0f;
    for (size_t i = 0; i < n; ++i) {
        float diff = q[i] - c[i];
        sum += diff * diff;
    }
    return sum;
}
float cosine_distance_f16(const float* q, const float* c, size_t n) {
    float dot = 0.0f, nq = 0.0f, nc = 0.0f;
    for (size_t i = 0; i < n; ++i) {
        dot += q[i] * c[i];
        nq += q[i] * q[i];
        nc += c[i] * c[i];
    }
    if (nq == 0.0f || nc == 0.0f) return 1.0f;
    return 1.0f - dot / (std::sqrt(nq) * std::sqrt(nc));
}
float manhattan_f16(const float* q, const float* c, size_t n) {
    float sum = 0.0f;
    for (size_t i = 0; i < n; ++i) sum += std::fabs(q[i] - c[i]);
    return sum;
}
float chebyshev_f16(const float* q, const float* c, size_t n) {
    float m = 0.0f;
    for (size_t i = 0; i < n; ++i) m = rust_max(m, std::fabs(q[i] - c[i]));
    return m;
}
float canberra_f16(const float* q, const float* c, size_t n) {
    float sum = 0.0f;
    for (size_t i = 0; i < n; ++i) {
        float den = std::fabs(q[i]) + std::fabs(c[i]);
        sum += den == 0.0f ? 0.0f : std::fabs(q[i] - c[i]) / den;
    }
    return sum;
}
float bray_curtis_f16(const float* q, const float* c, size_t n) {
    float num = 0.0f, den = 0.0f;
    for (size_t i = 0; i < n; ++i) {
        num += std::fabs(q[i] - c[i]);
        den += std::fabs(q[i] + c[i]);
    }
    if (den == 0.0f) return num == 0.0f ? 0.0f : kInf;
    return num / den;
}
float jensen_shannon_distance_f16(const float* a, const float* b, size_t n) {
    double sum_a = 0, sum_b = 0;
    for (size_t i = 0; i < n; ++i) {
        if (invalid_mass_value(a[i]) || invalid_mass_value(b[i])) return kInf;
        sum_a += (double)a[i];
        sum_b += (double)b[i];
    }
    if (sum_a == 0.0 || sum_b == 0.0) return sum_a == sum_b ? 0.0f : std::sqrt(kLn2);
    double divergence = 0;
    for (size_t i = 0; i < n; ++i) {
        double p = (double)a[i] / sum_a, q = (double)b[i] / sum_b, m = 0.5 * (p + q);
        if (p > 0.0) divergence += 0.5 * p * std::log(p / m);
        if (q > 0.0) divergence += 0.5 * q * std::log(q / m);
    }
    return (float)std::sqrt(std::max(divergence, 0.0));
}
float wasserstein_1d_f16(const float* a, const float* b, size_t n) {
    double sa = 0, sb = 0;
    for (size_t i = 0; i < n; ++i) {
        if (invalid_mass_value(a[i]) || invalid_mass_value(b[i])) return kInf;
        sa += (double)a[i];
        sb += (double)b[i];
    }
    if (sa == 0.0 || sb == 0.0) return sa == sb ? 0.0f : kInf;
    double cdf = 0, dist = 0;
    for (size_t i = 0; i + 1 < n; ++i) {
        cdf += (double)a[i] / sa - (double)b[i] / sb;
        dist += std::fabs(cdf);
    }
    return (float)dist;
}

// ---- distance/mod.rs:217-237 — compute_distance_f16 ---------------------------------------------------------
// (Haversine, correlation and Hellinger repeat their f32 formulas on the decoded row; the binary metrics count
// thresholded bits, which the f32 versions above already do element by element.)
float compute_distance_f16(const float* q, const float* c, size_t n, int metric) {
    switch (metric) {
        case IP: return inner_product_f16(q, c, n);
        case L2: return l2_squared_f16(q, c, n);
        case COSINE: return cosine_distance_f16(q, c, n);
        case HAMMING: return hamming_f32(q, c, n);
        case JACCARD:
        case TANIMOTO: return jaccard_f32(q, c, n);
        case MANHATTAN: return manhattan_f16(q, c, n);
        case HAVERSINE: return haversine_meters(q, c, n);
        case CORRELATION: return correlation_distance(q, c, n);
        case HELLINGER: return hellinger_distance(q, c, n);
        case WASSERSTEIN: return wasserstein_1d_f16(q, c, n);
        case DICE: return dice_f32(q, c, n);
        case JENSEN_SHANNON: return jensen_shannon_distance_f16(q, c, n);
        case CHEBYSHEV: return chebyshev_f16(q, c, n);
        case CANBERRA: return canberra_f16(q, c, n);
        case BRAY_CURTIS: return bray_curtis_f16(q, c, n);
    }
    return std::numeric_limits<float>::quiet_NaN();
}

// ---- flat_mmap.rs:1453-1476, 2132-2176 — sorted-array top-k -----------------
struct Entry {
    float dist;
    uint32_t idx;
};

struct TopK {
    std::vector<Entry> top;
    size_t k;
    bool asc;
    float threshold;
    bool filled = false;
    TopK(size_t k_, bool asc_) : k(k_), asc(asc_), threshold(asc_ ? kInf : -kInf) { top.reserve(k_); }

    void sort_now() {
        if (asc)
            std::stable_sort(top.begin(), top.end(), [](const Entry& a, const Entry& b) { return a.dist < b.dist; });
        else
            std::stable_sort(top.begin(), top.end(), [](const Entry& a, const Entry& b) { return a.dist > b.dist; });
    }
    inline bool passes(float d) const { return asc ? d < threshold : d > threshold; }
    // `if !filled || passes_threshold(dist, threshold) { topk_insert(..) }`
    inline void offer(float dist, uint32_t idx) {
        if (filled && !passes(dist)) return;
        if (!filled) {
            top.push_back({dist, idx});
            if (top.size() == k) {
                sort_now();
                threshold = top[k - 1].dist;
                filled = true;
            }
        } else {
            top[k - 1] = {dist, idx};
            size_t j = k - 1;
            if (asc) {
                while (j > 0 && top[j].dist < top[j - 1].dist) {
                    std::swap(top[j], top[j - 1]);
                    --j;
                }
            } else {
                while (j > 0 && top[j].dist > top[j - 1].dist) {
                    std::swap(top[j], top[j - 1]);
                    --j;
                }
            }
            threshold = top[k - 1].dist;
        }
    }
    void finish() {
        if (!filled && !top.empty()) sort_now();
    }
};

// ---- flat_mmap.rs:5183-5214 -------------------------------------------------
std::vector<Entry> merge_topk(const std::vector<std::vector<Entry>>& chunks, size_t k, bool asc) {
    TopK merged(k, asc);
    for (const auto& c : chunks)
        for (const auto& e : c) merged.offer(e.dist, e.idx);
    merged.finish();
    return merged.top;
}

// ---- flat_mmap.rs:4985-5044 -------------------------------------------------
template <class DistFn>
std::vector<Entry> fused_topk_seq(const float* cands, size_t n, size_t dim, size_t k, bool asc, DistFn&& fn) {
    TopK t(k, asc);
    for (size_t i = 0; i < n; ++i) t.offer(fn(cands + i * dim, i), (uint32_t)i);
    t.finish();
    return t.top;
}

// ---- flat_mmap.rs:4876-4982 — generic parallel scan ------------------------
// (the reference's 2-rows-per-iteration loop evaluates rows in index order, so
// a plain in-order loop is equivalent)
template <class DistFn>
std::vector<Entry> fused_topk_parallel(const float* cands, size_t n, size_t dim, size_t k, bool asc,
                                       int n_threads, DistFn&& fn) {
    if (n < 4096) return fused_topk_seq(cands, n, dim, k, asc, fn);
    size_t chunk_vecs = std::max<size_t>(n / (size_t)n_threads, 512);
    size_t n_chunks = (n + chunk_vecs - 1) / chunk_vecs;
    std::vector<std::vector<Entry>> results(n_chunks);
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
    for (long c = 0; c < (long)n_chunks; ++c) {
        size_t base = (size_t)c * chunk_vecs;
        size_t cnt = std::min(chunk_vecs, n - base);
        TopK t(k, asc);
        for (size_t i = 0; i < cnt; ++i) t.offer(fn(cands + (base + i) * dim, base + i), (uint32_t)(base + i));
        t.finish();
        results[c] = std::move(t.top);
    }
    return merge_topk(results, k, asc);
}

// ---- flat_mmap.rs:2179-2256 + 4845-4869 — IP scan ---------------------------
std::vector<Entry> ip_scan_chunk(const float* q, const float* chunk, size_t n_in_chunk, size_t dim, size_t k,
                                 size_t base_idx) {
    TopK t(k, false);
    size_t blocks8 = n_in_chunk / 8;
    for (size_t b = 0; b < blocks8; ++b) {
        const float* v[8];
        for (int r = 0; r < 8; ++r) v[r] = chunk + (b * 8 + r) * dim;
        float d[8];
        inner_product_batch8(q, v, dim, d);
        for (int r = 0; r < 8; ++r) t.offer(d[r], (uint32_t)(base_idx + b * 8 + r));
    }
    for (size_t i = blocks8 * 8; i < n_in_chunk; ++i)
        t.offer(inner_product_single(q, chunk + i * dim, dim), (uint32_t)(base_idx + i));
    t.finish();
    return t.top;
}

std::vector<Entry> fused_topk_ip_parallel(const float* q, const float* cands, size_t n, size_t dim, size_t k,
                                          int n_threads) {
    if (n < 4096)
        return fused_topk_seq(cands, n, dim, k, false,
                              [&](const float* c, size_t) { return inner_product_single(q, c, dim); });
    size_t chunk_vecs = std::max<size_t>(n / (size_t)n_threads, 512);
    size_t n_chunks = (n + chunk_vecs - 1) / chunk_vecs;
    std::vector<std::vector<Entry>> results(n_chunks);
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
    for (long c = 0; c < (long)n_chunks; ++c) {
        size_t base = (size_t)c * chunk_vecs;
        size_t cnt = std::min(chunk_vecs, n - base);
        results[c] = ip_scan_chunk(q, cands + base * dim, cnt, dim, k, base);
    }
    return merge_topk(results, k, false);
}

// ---- simd.rs:750-801 + flat_mmap.rs:1283-1343 — packed binary --------------
inline void pack_row(const float* src, size_t dim, uint64_t* words, size_t n_words, float threshold) {
    for (size_t w = 0; w < n_words; ++w) words[w] = 0;
    for (size_t i = 0; i < dim; ++i)
        if (src[i] > threshold) words[i / 64] |= (uint64_t)1 << (i % 64);
}

inline float packed_distance(const uint64_t* a, const uint64_t* b, size_t words, int metric) {
    if (metric == HAMMING) {
        uint32_t s = 0;
        for (size_t i = 0; i < words; ++i) s += (uint32_t)__builtin_popcountll(a[i] ^ b[i]);
        return (float)s;
    }
    if (metric == JACCARD || metric == TANIMOTO) {
        uint32_t inter = 0, uni = 0;
        for (size_t i = 0; i < words; ++i) {
            inter += (uint32_t)__builtin_popcountll(a[i] & b[i]);
            uni += (uint32_t)__builtin_popcountll(a[i] | b[i]);
        }
        return uni == 0 ? 0.0f : 1.0f - (float)inter / (float)uni;
    }
    uint32_t inter = 0, count = 0;  // DICE
    for (size_t i = 0; i < words; ++i) {
        inter += (uint32_t)__builtin_popcountll(a[i] & b[i]);
        count += (uint32_t)__builtin_popcountll(a[i]) + (uint32_t)__builtin_popcountll(b[i]);
    }
    return count == 0 ? 0.0f : 1.0f - (float)(2 * inter) / (float)count;
}

// ---- flat_mmap.rs:1345-1409 -------------------------------------------------
std::vector<Entry> packed_binary_search(const uint64_t* query, const uint64_t* data, size_t words, size_t n,
                                        size_t k, int metric, int n_threads) {
    if (n < 4096) {
        TopK t(k, true);
        for (size_t i = 0; i < n; ++i) t.offer(packed_distance(query, data + i * words, words, metric), (uint32_t)i);
        return t.top;  // (reference returns without the !filled sort; k <= n so it is always filled)
    }
    size_t chunk_rows = std::max<size_t>(n / (size_t)std::max(n_threads, 1), 1024);
    size_t n_chunks = (n + chunk_rows - 1) / chunk_rows;
    std::vector<std::vector<Entry>> results(n_chunks);
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
    for (long c = 0; c < (long)n_chunks; ++c) {
        size_t base = (size_t)c * chunk_rows;
        size_t cnt = std::min(chunk_rows, n - base);
        TopK t(k, true);
        for (size_t i = 0; i < cnt; ++i)
            t.offer(packed_distance(query, data + (base + i) * words, words, metric), (uint32_t)(base + i));
        results[c] = std::move(t.top);
    }
    return merge_topk(results, k, true);
}

// ---- flat_mmap.rs:926-1111 — Jensen-Shannon cached scan ---------------------
std::vector<Entry> jensen_shannon_cached_search(const float* query, const float* cands, const float* stats,
                                                size_t n, size_t dim, size_t k, int n_threads, bool* handled) {
    float inv_mass, entropy;
    probability_row_stats(query, dim, &inv_mass, &entropy);
    *handled = true;
    if (std::isnan(inv_mass) || !std::isfinite(entropy) || (inv_mass != 0.0f && !std::isfinite(inv_mass))) {
        *handled = false;  // prepare_jensen_shannon_query -> None: caller falls back to exact_flat_search
        return {};
    }
    if (inv_mass == 0.0f) {  // zero-mass query (flat_mmap.rs:938-972)
        return fused_topk_parallel(cands, n, dim, k, true, n_threads, [&](const float*, size_t row) {
            float im = stats[2 * row], en = stats[2 * row + 1];
            if (std::isnan(im) || !std::isfinite(en)) return kInf;
            if (im == 0.0f) return 0.0f;
            return std::sqrt(kLn2);
        });
    }
    std::vector<float> nq(dim);
    for (size_t i = 0; i < dim; ++i) nq[i] = query[i] * inv_mass;
    if (n < 4096) {
        return fused_topk_seq(cands, n, dim, k, true, [&](const float* c, size_t row) {
            return jensen_shannon_precomputed(nq.data(), c, dim, entropy, stats[2 * row], stats[2 * row + 1]);
        });
    }
    // parallel path ranks on divergence (squared distance), sqrt on the final k
    size_t chunk_vecs = std::max<size_t>(n / (size_t)n_threads, 512);
    size_t n_chunks = (n + chunk_vecs - 1) / chunk_vecs;
    std::vector<std::vector<Entry>> results(n_chunks);
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
    for (long c = 0; c < (long)n_chunks; ++c) {
        size_t base = (size_t)c * chunk_vecs;
        size_t cnt = std::min(chunk_vecs, n - base);
        TopK t(k, true);
        size_t pairs = cnt / 2;
        for (size_t i = 0; i < pairs * 2; ++i) {
            size_t row = base + i;
            t.offer(jensen_shannon_precomputed_divergence(nq.data(), cands + row * dim, dim, entropy, stats[2 * row],
                                                          stats[2 * row + 1]),
                    (uint32_t)row);
        }
        if (cnt % 2 == 1) {
            size_t row = base + cnt - 1;
            float d = jensen_shannon_precomputed(nq.data(), cands + row * dim, dim, entropy, stats[2 * row],
                                                 stats[2 * row + 1]);
            t.offer(d * d, (uint32_t)row);
        }
        t.finish();
        results[c] = std::move(t.top);
    }
    auto merged = merge_topk(results, k, true);
    for (auto& e : merged) e.dist = std::sqrt(e.dist);
    return merged;
}

// ---- flat_mmap.rs:824-923, 1173-1230 — FlatMmap::search for one segment -----
// (exact unfiltered f32 path; approx / SQ8 / f16 branches are out of scope)
std::vector<Entry> flat_search(const float* cands, size_t n, size_t dim, const float* query, size_t k, int metric,
                               int n_threads) {
    if (n == 0 || k == 0) return {};
    k = std::min(k, n);
    if (is_binary(metric)) {
        size_t words = (dim + 63) / 64;
        std::vector<uint64_t> data(n * words), q(words);
#pragma omp parallel for schedule(static) num_threads(n_threads)
        for (long r = 0; r < (long)n; ++r) pack_row(cands + (size_t)r * dim, dim, data.data() + (size_t)r * words, words, 0.5f);
        pack_row(query, dim, q.data(), words, 0.5f);
        return packed_binary_search(q.data(), data.data(), words, n, k, metric, n_threads);
    }
    if (metric == JENSEN_SHANNON) {
        std::vector<float> stats(2 * n);
#pragma omp parallel for schedule(static) num_threads(n_threads)
        for (long r = 0; r < (long)n; ++r)
            probability_row_stats(cands + (size_t)r * dim, dim, &stats[2 * r], &stats[2 * r + 1]);
        bool handled = false;
        auto res = jensen_shannon_cached_search(query, cands, stats.data(), n, dim, k, n_threads, &handled);
        if (handled) return res;
    }
    if (metric == IP) return fused_topk_ip_parallel(query, cands, n, dim, k, n_threads);
    return fused_topk_parallel(cands, n, dim, k, true, n_threads,
                               [&](const float* c, size_t) { return compute_distance(query, c, dim, metric); });
}

// ---- flat_mmap.rs:905-907, 1259-1281, 5047-5180 — FlatMmap::search on an F16 segment -----------------------
// binary metrics take the packed cache first (as for f32 rows); everything else is fused_topk_parallel_f16 over
// the scalar kernels (no batch-8 inner product, no Jensen-Shannon cache).
std::vector<Entry> flat_search_f16(const float* cands, size_t n, size_t dim, const float* query, size_t k, int metric,
                                   int n_threads) {
    if (n == 0 || k == 0) return {};
    if (is_binary(metric)) return flat_search(cands, n, dim, query, k, metric, n_threads);
    k = std::min(k, n);
    return fused_topk_parallel(cands, n, dim, k, is_ascending(metric), n_threads,
                               [&](const float* c, size_t) { return compute_distance_f16(query, c, dim, metric); });
}

// ---- vector_store.rs:953-1004 — segment fan-out + global merge --------------
struct Hit {
    uint64_t row;
    float dist;
};

std::vector<Hit> store_search(const float* cands, const uint64_t* seg_rows, size_t n_segs, size_t dim,
                              const float* query, size_t k, int metric, int n_threads, bool f16_rows = false) {
    std::vector<Hit> merged;
    uint64_t base = 0;
    for (size_t s = 0; s < n_segs; ++s) {
        auto local = f16_rows ? flat_search_f16(cands + base * dim, seg_rows[s], dim, query, k, metric, n_threads)
                              : flat_search(cands + base * dim, seg_rows[s], dim, query, k, metric, n_threads);
        for (const auto& e : local) merged.push_back({base + e.idx, e.dist});
        base += seg_rows[s];
    }
    bool asc = is_ascending(metric);
    // partial_cmp().unwrap_or(Equal), reversed for IP, then row ascending
    std::stable_sort(merged.begin(), merged.end(), [asc](const Hit& a, const Hit& b) {
        if (a.dist < b.dist) return asc;
        if (a.dist > b.dist) return !asc;
        return a.row < b.row;
    });
    if (merged.size() > k) merged.resize(k);
    return merged;
}

// ---- distance/mod.rs:304-362 — median-of-3 Lomuto quickselect ---------------
struct Pair {
    float d;
    uint32_t i;
};
inline int cmp_pair(const Pair& a, const Pair& b, bool asc) {
    // Ordering of a vs b: -1 Less, 0 Equal (incl. NaN), +1 Greater
    float x = asc ? a.d : b.d, y = asc ? b.d : a.d;
    if (x < y) return -1;
    if (x > y) return 1;
    return 0;
}
void quickselect_k(std::vector<Pair>& arr, size_t k, bool asc) {
    size_t n = arr.size();
    if (n <= k || k == 0) return;
    size_t target = k - 1, lo = 0, hi = n - 1;
    while (lo < hi) {
        if (hi - lo >= 2) {
            size_t mid = lo + (hi - lo) / 2;
            if (cmp_pair(arr[lo], arr[mid], asc) > 0) std::swap(arr[lo], arr[mid]);
            if (cmp_pair(arr[lo], arr[hi], asc) > 0) std::swap(arr[lo], arr[hi]);
            if (cmp_pair(arr[mid], arr[hi], asc) > 0) std::swap(arr[mid], arr[hi]);
            std::swap(arr[mid], arr[hi]);
        }
        Pair pivot = arr[hi];
        size_t store = lo;
        for (size_t j = lo; j < hi; ++j) {
            if (cmp_pair(arr[j], pivot, asc) <= 0) {
                std::swap(arr[store], arr[j]);
                ++store;
            }
        }
        std::swap(arr[store], arr[hi]);
        if (store == target) return;
        if (store < target)
            lo = store + 1;
        else
            hi = store - 1;
    }
}

// ---- distance/mod.rs:373-422 ------------------------------------------------
std::vector<Pair> top_k_search(const float* q, const float* cands, size_t n, size_t dim, size_t k, int metric,
                               int n_threads) {
    k = std::min(k, n);
    if (n == 0 || k == 0) return {};
    bool asc = is_ascending(metric);
    std::vector<Pair> pairs(n);
#pragma omp parallel for schedule(static) num_threads(n_threads) if (n >= 8192)
    for (long i = 0; i < (long)n; ++i) pairs[i] = {compute_distance(q, cands + (size_t)i * dim, dim, metric), (uint32_t)i};
    quickselect_k(pairs, k, asc);
    pairs.resize(k);
    std::stable_sort(pairs.begin(), pairs.end(), [asc](const Pair& a, const Pair& b) { return asc ? a.d < b.d : a.d > b.d; });
    return pairs;
}

// ---- src/index/kmeans.rs — IVF k-means (farthest-point init on a seeded sample + Lloyd) ----------------
// kmeans.rs:21-48
struct FastRng {
    uint64_t s;
    explicit FastRng(uint64_t seed) : s(seed) {}
    double next_f64() {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        return (double)(s >> 33) / (double)(1ull << 31);
    }
    std::vector<size_t> sample_indices(size_t n, size_t count) {
        count = std::min(count, n);
        std::vector<size_t> idx(n);
        for (size_t i = 0; i < n; ++i) idx[i] = i;
        for (size_t i = 0; i < count; ++i) {
            size_t j = i + std::min((size_t)(next_f64() * (double)(n - i)), n - i - 1);
            std::swap(idx[i], idx[j]);
        }
        idx.resize(count);
        return idx;
    }
};
inline size_t adaptive_init_sample_size(size_t n, size_t k) {  // kmeans.rs:50-53
    size_t v = std::min<size_t>(std::max<size_t>(k * 32, 2048), 10000);
    return std::min(n, v);
}
// kmeans.rs:237-264 — first best under strict <
void assign_metric(const float* data, size_t n, const float* centroids, size_t dim, size_t nc, int metric, uint32_t* out) {
    bool asc = is_ascending(metric);
    for (size_t i = 0; i < n; ++i) {
        size_t best = 0;
        float best_rank = std::numeric_limits<float>::max();
        for (size_t c = 0; c < nc; ++c) {
            float raw = compute_distance(data + i * dim, centroids + c * dim, dim, metric);
            float rank = asc ? raw : -raw;
            if (rank < best_rank) {
                best_rank = rank;
                best = c;
            }
        }
        out[i] = (uint32_t)best;
    }
}
// kmeans.rs:141-196
std::vector<float> kmeans_pp_init(const float* data, size_t n, size_t dim, size_t k, int metric) {
    FastRng rng(42);
    size_t sample_n = adaptive_init_sample_size(n, k);
    std::vector<size_t> sidx;
    if (sample_n >= n) {
        sidx.resize(n);
        for (size_t i = 0; i < n; ++i) sidx[i] = i;
    } else {
        sidx = rng.sample_indices(n, sample_n);
    }
    sample_n = sidx.size();
    std::vector<float> sample(sample_n * dim);
    for (size_t i = 0; i < sample_n; ++i) memcpy(&sample[i * dim], data + sidx[i] * dim, dim * 4);
    bool asc = is_ascending(metric);
    std::vector<float> centroids(k * dim, 0.0f);
    size_t first = (size_t)(rng.next_f64() * (double)sample_n) % sample_n;
    memcpy(&centroids[0], &sample[first * dim], dim * 4);
    std::vector<float> min_ranks(sample_n, std::numeric_limits<float>::max());
    for (size_t c = 1; c < k; ++c) {
        const float* prev = &centroids[(c - 1) * dim];
        for (size_t i = 0; i < sample_n; ++i) {
            float raw = compute_distance(&sample[i * dim], prev, dim, metric);
            float rank = asc ? raw : -raw;
            if (rank < min_ranks[i]) min_ranks[i] = rank;
        }
        // Iterator::max_by keeps the LAST of several equal maxima
        size_t best = 0;
        for (size_t i = 1; i < sample_n; ++i)
            if (!(min_ranks[i] < min_ranks[best])) best = i;
        memcpy(&centroids[c * dim], &sample[best * dim], dim * 4);
    }
    return centroids;
}
// kmeans.rs:74-139 (centroid sums in row order, the n < 8192 branch of accumulate_centroid_sums :266-315; the
// parallel fold/reduce branch for larger n is order-dependent in the reference itself)
size_t kmeans_train(const float* data, size_t n, size_t dim, size_t requested, size_t max_iter, int metric,
                    std::vector<float>& centroids, std::vector<uint32_t>& assignments) {
    size_t nc = std::min(requested, n);
    centroids.clear();
    assignments.clear();
    if (n == 0 || nc == 0 || dim == 0) return 0;
    centroids = kmeans_pp_init(data, n, dim, nc, metric);
    assignments.assign(n, 0xFFFFFFFFu);
    std::vector<uint32_t> fresh(n);
    for (size_t it = 0; it < max_iter; ++it) {
        assign_metric(data, n, centroids.data(), dim, nc, metric, fresh.data());
        bool changed = fresh != assignments;
        assignments = fresh;
        std::vector<float> sums(nc * dim, 0.0f);
        std::vector<uint32_t> counts(nc, 0);
        for (size_t i = 0; i < n; ++i) {
            uint32_t c = assignments[i];
            counts[c] += 1;
            for (size_t d = 0; d < dim; ++d) sums[c * dim + d] += data[i * dim + d];
        }
        size_t max_c = 0;
        uint32_t max_count = 0;
        for (size_t c = 0; c < nc; ++c)  // max_by_key keeps the last maximum
            if (counts[c] >= max_count) {
                max_count = counts[c];
                max_c = c;
            }
        for (size_t c = 0; c < nc; ++c) {
            if (counts[c] > 0) {
                float inv = 1.0f / (float)counts[c];
                for (size_t d = 0; d < dim; ++d) centroids[c * dim + d] = sums[c * dim + d] * inv;
            } else if (max_count > 1) {
                for (size_t d = 0; d < dim; ++d) centroids[c * dim + d] = centroids[max_c * dim + d] * (1.0f + 1e-4f * (float)d);
            }
        }
        if (!changed) break;
    }
    assign_metric(data, n, centroids.data(), dim, nc, metric, assignments.data());
    return nc;
}

// ---- src/index/ivf.rs:181-348 — IVFIndex::search without a quantizer ---------------------------------------
// ids are row positions; `allow` (optional) is the subset filter as a row bitset.
std::vector<Pair> ivf_search(const float* data, size_t n, size_t dim, const float* centroids, size_t nc,
                             const uint32_t* assignments, const float* query, size_t k, size_t nprobe, int metric,
                             const uint64_t* allow) {
    if (n == 0) return {};
    nprobe = std::max<size_t>(nprobe, 1);
    bool asc = is_ascending(metric), binary = is_binary(metric);
    int routing = binary ? L2 : metric;  // ivf.rs:80-87
    bool rasc = is_ascending(routing);
    std::vector<std::pair<float, size_t>> cd(nc);
    for (size_t c = 0; c < nc; ++c) cd[c] = {compute_distance(query, centroids + c * dim, dim, routing), c};
    std::stable_sort(cd.begin(), cd.end(), [rasc](const auto& a, const auto& b) { return rasc ? a.first < b.first : a.first > b.first; });
    // inverted lists in row order (kmeans.rs:317-345)
    std::vector<std::vector<uint32_t>> lists(nc);
    for (size_t i = 0; i < n; ++i) lists[assignments[i]].push_back((uint32_t)i);
    auto allowed = [&](uint32_t r) { return allow == nullptr || ((allow[r >> 6] >> (r & 63)) & 1ull); };
    std::vector<uint32_t> cand;
    for (size_t p = 0; p < std::min(nprobe, nc); ++p)
        for (uint32_t r : lists[cd[p].second])
            if (allowed(r)) cand.push_back(r);
    if (cand.empty())
        for (size_t r = 0; r < n; ++r)
            if (allowed((uint32_t)r)) cand.push_back((uint32_t)r);
    if (cand.empty()) return {};
    size_t pool = std::min(k, cand.size());
    std::vector<Pair> scored(cand.size());
    if (binary) {
        size_t words = (dim + 63) / 64;
        std::vector<uint64_t> pq(words), pr(words);
        pack_row(query, dim, pq.data(), words, 0.5f);
        for (size_t i = 0; i < cand.size(); ++i) {
            pack_row(data + (size_t)cand[i] * dim, dim, pr.data(), words, 0.5f);
            scored[i] = {packed_distance(pq.data(), pr.data(), words, metric), (uint32_t)i};
        }
    } else {
        for (size_t i = 0; i < cand.size(); ++i) scored[i] = {compute_distance(query, data + (size_t)cand[i] * dim, dim, metric), (uint32_t)i};
    }
    quickselect_k(scored, pool, asc);
    scored.resize(pool);
    std::stable_sort(scored.begin(), scored.end(), [asc](const Pair& a, const Pair& b) { return asc ? a.d < b.d : a.d > b.d; });
    for (auto& p : scored) p.i = cand[p.i];
    return scored;
}

// ---- src/storage/ivf_flat_mmap.rs — the standalone IVF_FLAT index (`_core.IvfFlatIndex`) ----------------------
// Training is kmeans::train_l2 = train_for_metric(.., L2Squared) (kmeans.rs:53-72), i.e. kmeans_train above.
// select_routing_dims (ivf_flat_mmap.rs:312-345): the 16 centroid dimensions of highest variance, ascending.
// The reference picks them with select_nth_unstable_by (ties arbitrary); here ties break on the lower dimension.
std::vector<size_t> ivf_flat_routing_dims(const float* centroids, size_t dim, size_t nc) {
    if (nc == 0 || dim == 0 || dim < 64 || nc < 64) return {};
    std::vector<float> sums(dim, 0.0f), sq(dim, 0.0f);
    for (size_t c = 0; c < nc; ++c)
        for (size_t d = 0; d < dim; ++d) {
            float v = centroids[c * dim + d];
            sums[d] += v;
            sq[d] += v * v;
        }
    float inv_k = 1.0f / (float)nc;
    size_t keep = std::min<size_t>(16, dim);
    std::vector<std::pair<float, size_t>> dims(dim);
    for (size_t d = 0; d < dim; ++d) {
        float mean = sums[d] * inv_k;
        float variance = sq[d] * inv_k - mean * mean;
        dims[d] = {variance, d};
    }
    std::stable_sort(dims.begin(), dims.end(), [](const auto& a, const auto& b) { return a.first > b.first; });
    std::vector<size_t> sel(keep);
    for (size_t i = 0; i < keep; ++i) sel[i] = dims[i].second;
    std::sort(sel.begin(), sel.end());
    return sel;
}

// find_nearest_centroids (ivf_flat_mmap.rs:383-446): every partition when nprobe >= n_partitions; for inner
// product on dim >= 64 with >= 64 partitions a routing-dimension shortlist (clamp(3*nprobe, 24, 96) entries,
// shortlist_insert replaces the first worst entry when the new score is strictly larger) re-ranked with the full
// metric; otherwise the nprobe nearest centroids under the SEARCH metric.  Selection ties are arbitrary in the
// reference (select_nth_unstable_by); here they break on the lower centroid index.
std::vector<size_t> ivf_flat_nearest_centroids(const float* query, const float* centroids, size_t dim, size_t nc,
                                               size_t nprobe, int metric, const std::vector<size_t>& routing_dims) {
    std::vector<size_t> out;
    if (nprobe >= nc) {
        for (size_t c = 0; c < nc; ++c) out.push_back(c);
        return out;
    }
    bool asc = is_ascending(metric);
    std::vector<std::pair<float, size_t>> dists;
    if (metric == IP && dim >= 64 && nc >= 64 && !routing_dims.empty()) {
        size_t shortlist = std::min<size_t>(std::max<size_t>(std::min<size_t>(nprobe * 3, 96), 24), nc);
        std::vector<std::pair<float, uint32_t>> best(shortlist, {-INFINITY, 0u});
        size_t len = 0;
        for (size_t c = 0; c < nc; ++c) {
            float score = 0.0f;
            for (size_t d : routing_dims) score += query[d] * centroids[c * dim + d];
            if (len < best.size()) {
                best[len++] = {score, (uint32_t)c};
                continue;
            }
            size_t worst = 0;
            float worst_score = best[0].first;
            for (size_t i = 1; i < best.size(); ++i)
                if (best[i].first < worst_score) {
                    worst_score = best[i].first;
                    worst = i;
                }
            if (score > worst_score) best[worst] = {score, (uint32_t)c};
        }
        for (size_t i = 0; i < len; ++i)
            dists.push_back({compute_distance(query, centroids + (size_t)best[i].second * dim, dim, metric), best[i].second});
        asc = false;
    } else {
        for (size_t c = 0; c < nc; ++c) dists.push_back({compute_distance(query, centroids + c * dim, dim, metric), c});
    }
    std::sort(dists.begin(), dists.end(), [asc](const auto& a, const auto& b) {
        if (a.first != b.first) return asc ? a.first < b.first : a.first > b.first;
        return a.second < b.second;
    });
    for (size_t i = 0; i < std::min(nprobe, dists.size()); ++i) out.push_back(dists[i].second);
    return out;
}

// IvfFlatMmap::search (ivf_flat_mmap.rs:225-300): compute_distance_f32 over every row of the probed partitions,
// select_nth + sort on the distance alone (ties arbitrary in the reference; here the lower original id first).
std::vector<Pair> ivf_flat_search(const float* data, size_t n, size_t dim, const float* centroids, size_t nc,
                                  const uint32_t* assignments, const float* query, size_t k, size_t nprobe, int metric) {
    if (n == 0 || k == 0) return {};
    k = std::min(k, n);
    nprobe = std::min(std::max<size_t>(nprobe, 1), nc);
    bool asc = is_ascending(metric);
    auto rdims = ivf_flat_routing_dims(centroids, dim, nc);
    auto parts = ivf_flat_nearest_centroids(query, centroids, dim, nc, nprobe, metric, rdims);
    std::vector<char> probed(nc, 0);
    for (size_t p : parts) probed[p] = 1;
    std::vector<Pair> best;
    for (size_t i = 0; i < n; ++i)
        if (probed[assignments[i]]) best.push_back({compute_distance(query, data + i * dim, dim, metric), (uint32_t)i});
    std::sort(best.begin(), best.end(), [asc](const Pair& a, const Pair& b) {
        if (a.d != b.d) return asc ? a.d < b.d : a.d > b.d;
        return a.i < b.i;
    });
    if (best.size() > k) best.resize(k);
    return best;
}

}  // namespace

// =============================== C ABI =======================================
extern "C" {

int lo_metric_count(void) { return METRIC_COUNT; }

float lo_compute_distance(const float* a, const float* b, uint64_t dim, int metric) {
    return compute_distance(a, b, dim, metric);
}

// IP in the batch-8 lane order for a single row (the order the scan uses for
// rows inside full blocks of 8; simd.rs:1450-1525).
float lo_inner_product_batch8_order(const float* q, const float* row, uint64_t dim) {
    const float* v[8] = {row, row, row, row, row, row, row, row};
    float out[8];
    inner_product_batch8(q, v, dim, out);
    return out[0];
}

void lo_probability_row_stats(const float* rows, uint64_t n, uint64_t dim, float* stats /*2n*/) {
    for (uint64_t r = 0; r < n; ++r) probability_row_stats(rows + r * dim, dim, &stats[2 * r], &stats[2 * r + 1]);
}

float lo_jensen_shannon_precomputed(const float* nq, const float* cand, uint64_t dim, float query_entropy,
                                    float inv_mass, float entropy) {
    return jensen_shannon_precomputed(nq, cand, dim, query_entropy, inv_mass, entropy);
}

float lo_jensen_shannon_precomputed_divergence(const float* nq, const float* cand, uint64_t dim, float query_entropy,
                                               float inv_mass, float entropy) {
    return jensen_shannon_precomputed_divergence(nq, cand, dim, query_entropy, inv_mass, entropy);
}

void lo_pack_binary(const float* rows, uint64_t n, uint64_t dim, float threshold, uint64_t* words_out) {
    size_t words = (dim + 63) / 64;
    for (uint64_t r = 0; r < n; ++r) pack_row(rows + r * dim, dim, words_out + r * words, words, threshold);
}

float lo_packed_distance(const uint64_t* a, const uint64_t* b, uint64_t words, int metric) {
    return packed_distance(a, b, words, metric);
}

// Search packed rows directly (flat_mmap.rs:1345-1409).
uint32_t lo_packed_search(const uint64_t* query, const uint64_t* data, uint64_t words, uint64_t n, uint32_t k,
                          int metric, int n_threads, uint32_t* ids, float* dists) {
    if (n == 0 || k == 0) return 0;
    size_t kk = std::min<size_t>(k, n);
    auto res = packed_binary_search(query, data, words, n, kk, metric, std::max(n_threads, 1));
    for (size_t i = 0; i < res.size(); ++i) {
        ids[i] = res[i].idx;
        dists[i] = res[i].dist;
    }
    return (uint32_t)res.size();
}

// distance::top_k_search (stateless operator)
uint32_t lo_top_k_search(const float* q, const float* cands, uint64_t n, uint64_t dim, uint32_t k, int metric,
                         int n_threads, uint32_t* ids, float* dists) {
    auto res = top_k_search(q, cands, n, dim, k, metric, std::max(n_threads, 1));
    for (size_t i = 0; i < res.size(); ++i) {
        ids[i] = res[i].i;
        dists[i] = res[i].d;
    }
    return (uint32_t)res.size();
}

// FlatMmap::search on one segment.
uint32_t lo_flat_search(const float* cands, uint64_t n, uint64_t dim, const float* query, uint32_t k, int metric,
                        int n_threads, uint32_t* ids, float* dists) {
    auto res = flat_search(cands, n, dim, query, k, metric, std::max(n_threads, 1));
    for (size_t i = 0; i < res.size(); ++i) {
        ids[i] = res[i].idx;
        dists[i] = res[i].dist;
    }
    return (uint32_t)res.size();
}

// VectorStore::search for a batch of queries = sequential loop over queries
// (engine.rs:5484-5497).  ids/dists are [nq][k]; counts[nq].
void lo_store_batch_search(const float* cands, const uint64_t* seg_rows, uint64_t n_segs, uint64_t dim,
                           const float* queries, uint64_t nq, uint32_t k, int metric, int n_threads, uint64_t* ids,
                           float* dists, uint32_t* counts) {
    for (uint64_t q = 0; q < nq; ++q) {
        auto res = store_search(cands, seg_rows, n_segs, dim, queries + q * dim, k, metric, std::max(n_threads, 1));
        counts[q] = (uint32_t)res.size();
        for (size_t i = 0; i < res.size(); ++i) {
            ids[q * k + i] = res[i].row;
            dists[q * k + i] = res[i].dist;
        }
    }
}

// Packed-binary store search on pre-packed rows (C4-shaped workloads, where the
// f32 source would be 64x larger than the fingerprints).  One segment.
void lo_packed_batch_search(const uint64_t* data, uint64_t words, uint64_t n, const uint64_t* queries, uint64_t nq,
                            uint32_t k, int metric, int n_threads, uint64_t* ids, float* dists, uint32_t* counts) {
    for (uint64_t q = 0; q < nq; ++q) {
        size_t kk = std::min<size_t>(k, n);
        std::vector<Entry> res;
        if (n > 0 && kk > 0) res = packed_binary_search(queries + q * words, data, words, n, kk, metric, std::max(n_threads, 1));
        counts[q] = (uint32_t)res.size();
        for (size_t i = 0; i < res.size(); ++i) {
            ids[q * k + i] = res[i].idx;
            dists[q * k + i] = res[i].dist;
        }
    }
}

// compute_distance_f16 on a decoded binary16 row
float lo_compute_distance_f16(const float* q, const float* row, uint64_t dim, int metric) {
    return compute_distance_f16(q, row, dim, metric);
}

// VectorStore::search of a float16 collection, one query (each segment: FlatMmap::search on F16 rows).
// `cands` holds the decoded rows.
uint32_t lo_store_search_f16(const float* cands, const uint64_t* seg_rows, uint64_t n_segs, uint64_t dim, const float* query,
                             uint32_t k, int metric, int n_threads, uint64_t* ids, float* dists) {
    auto res = store_search(cands, seg_rows, n_segs, dim, query, k, metric, std::max(n_threads, 1), true);
    for (size_t i = 0; i < res.size(); ++i) {
        ids[i] = res[i].row;
        dists[i] = res[i].dist;
    }
    return (uint32_t)res.size();
}

// kmeans::train_for_metric; returns the number of centroids (min(requested, n))
uint32_t lo_kmeans_train(const float* data, uint64_t n, uint64_t dim, uint32_t requested, uint32_t max_iter, int metric,
                         float* centroids_out /*[requested][dim]*/, uint32_t* assignments_out /*[n]*/) {
    std::vector<float> c;
    std::vector<uint32_t> a;
    size_t nc = kmeans_train(data, n, dim, requested, max_iter, metric, c, a);
    if (nc) {
        memcpy(centroids_out, c.data(), nc * dim * 4);
        memcpy(assignments_out, a.data(), n * 4);
    }
    return (uint32_t)nc;
}

// IVFIndex::search for one query given centroids and assignments
uint32_t lo_ivf_search(const float* data, uint64_t n, uint64_t dim, const float* centroids, uint32_t nc,
                       const uint32_t* assignments, const float* query, uint32_t k, uint32_t nprobe, int metric,
                       const uint64_t* allow_bits, uint32_t* ids, float* dists) {
    auto res = ivf_search(data, n, dim, centroids, nc, assignments, query, k, nprobe, metric, allow_bits);
    for (size_t i = 0; i < res.size(); ++i) {
        ids[i] = res[i].i;
        dists[i] = res[i].d;
    }
    return (uint32_t)res.size();
}

// IvfFlatMmap::search for one query given centroids and assignments; probes_out (optional) receives the probed
// partitions, n_probes_out their count
uint32_t lo_ivf_flat_search(const float* data, uint64_t n, uint64_t dim, const float* centroids, uint32_t nc,
                            const uint32_t* assignments, const float* query, uint32_t k, uint32_t nprobe, int metric,
                            uint32_t* ids, float* dists, uint32_t* probes_out, uint32_t* n_probes_out) {
    auto res = ivf_flat_search(data, n, dim, centroids, nc, assignments, query, k, nprobe, metric);
    for (size_t i = 0; i < res.size(); ++i) {
        ids[i] = res[i].i;
        dists[i] = res[i].d;
    }
    if (probes_out && n_probes_out && n && k) {
        size_t np = std::min<size_t>(std::max<uint32_t>(nprobe, 1), nc);
        auto parts = ivf_flat_nearest_centroids(query, centroids, dim, nc, np, metric, ivf_flat_routing_dims(centroids, dim, nc));
        *n_probes_out = (uint32_t)parts.size();
        for (size_t i = 0; i < parts.size(); ++i) probes_out[i] = (uint32_t)parts[i];
    }
    return (uint32_t)res.size();
}

// select_routing_dims; returns the count (0 or 16)
uint32_t lo_ivf_flat_routing_dims(const float* centroids, uint64_t dim, uint32_t nc, uint32_t* dims_out) {
    auto d = ivf_flat_routing_dims(centroids, dim, nc);
    for (size_t i = 0; i < d.size(); ++i) dims_out[i] = (uint32_t)d[i];
    return (uint32_t)d.size();
}

int lo_max_threads(void) { return omp_get_max_threads(); }

}  // extern "C"
