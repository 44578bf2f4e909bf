// lb_common.cuh — shared definitions for liblynse_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <string>

#include "../../include/lynse_b200.h"

namespace lb {

// ---- error plumbing ---------------------------------------------------------
void set_error(const std::string& msg);
int fail(int status, const std::string& msg);

#define LB_CUDA_TRY(expr)                                                                          \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            return ::lb::fail(LB_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));        \
        }                                                                                          \
    } while (0)

#define LB_TRY(expr)                       \
    do {                                   \
        int _s = (expr);                   \
        if (_s != LB_OK) return _s;        \
    } while (0)

// ---- metric helpers (src/distance/mod.rs:111-116, :161-166) ------------------
__host__ __device__ inline bool metric_ascending(int m) { return m != LB_IP; }
__host__ __device__ inline bool metric_binary(int m) {
    return m == LB_HAMMING || m == LB_JACCARD || m == LB_DICE || m == LB_TANIMOTO;
}

// ---- 64-bit ranking keys ------------------------------------------------------
// The reference's result order is (score best-first, row ascending)
// (src/storage/vector_store.rs:959-967, flat_mmap.rs:2132-2176, :5183-5214).
// key = (rank(score) << 32) | row, with rank() monotone so that a SMALLER key is a
// BETTER hit; the k smallest keys, ascending, are the reference's answer.
__host__ __device__ inline uint32_t f32_orderable(float f) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline uint32_t f32_orderable_bits(uint32_t u) { return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__host__ __device__ inline float f32_from_orderable(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
constexpr uint64_t KEY_NONE = 0xFFFFFFFFFFFFFFFFull;
constexpr uint32_t ROW_NONE = 0xFFFFFFFFu;

template <bool ASC>
__host__ __device__ inline uint64_t make_key(float score, uint32_t row) {
    score = score + 0.0f;  // -0.0 -> +0.0 so equal scores tie on row, as partial_cmp does
    uint32_t o = f32_orderable(score);
    if (!ASC) o = ~o;
    return ((uint64_t)o << 32) | row;
}
template <bool ASC>
__host__ __device__ inline float key_score(uint64_t key) {
    uint32_t o = (uint32_t)(key >> 32);
    if (!ASC) o = ~o;
    return f32_from_orderable(o);
}
__host__ __device__ inline uint32_t key_row(uint64_t key) { return (uint32_t)key; }

// ---- synthetic data (bench + large parity tests) ---------------------------------
// murmur3 fmix64 of (seed, index); reproduced in numpy by lynsedb_b200/synthetic.py
__host__ __device__ inline uint64_t mix64(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}
__host__ __device__ inline uint64_t synth_u64(uint64_t seed, uint64_t index) {
    return mix64(index * 0x9E3779B97F4A7C15ull + mix64(seed + 0x632BE59BD9B4E019ull));
}
__host__ __device__ inline float synth_f32(uint64_t seed, uint64_t index) {
    return (float)(uint32_t)(synth_u64(seed, index) >> 40) * (1.0f / 16777216.0f);
}

inline uint64_t ceil_div(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

}  // namespace lb
