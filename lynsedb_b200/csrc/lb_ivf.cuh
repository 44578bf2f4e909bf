// lb_ivf.cuh — IVF build kernels: k-means assignment, centroid update, farthest-point initialisation.
//
// Replaces src/index/kmeans.rs (train_for_metric :74-139, kmeans_pp_init_metric :141-196, assign_metric :237-264,
// accumulate_centroid_sums :266-315).  Distances are compute_distance_f32(vector, centroid, metric) in the
// reference's order (lb_metrics.cuh); sums run over a cluster's members in row order, i.e. exactly the
// reference's sequential branch (its rayon fold/reduce branch for n >= 8192 is order-dependent in the reference
// itself, so above that size ours is one valid outcome).
#pragma once
#include "lb_scan.cuh"

namespace lb {

// assign_metric: first centroid of minimal rank (rank = distance, or -score for IP), strict <
static __global__ void ivf_assign_kernel(const float* __restrict__ rows, uint64_t n, int dim, const float* __restrict__ centroids, int nc,
                                  int metric, uint32_t* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool vec = (dim & 3) == 0;
    const float* v = rows + i * dim;
    const bool asc = metric_ascending(metric);
    uint32_t best = 0;
    float best_rank = 3.402823466e+38f;
    for (int c = 0; c < nc; ++c) {
        const float raw = compute_distance<true>(metric, v, centroids + (size_t)c * dim, dim, vec);
        const float rank = asc ? raw : -raw;
        if (rank < best_rank) {
            best_rank = rank;
            best = (uint32_t)c;
        }
    }
    out[i] = best;
}

// one CTA per centroid, one thread per dimension: sum of the members in row order, then * (1 / count)
static __global__ void ivf_centroid_update_kernel(const float* __restrict__ rows, int dim, const uint32_t* __restrict__ offsets,
                                           const uint32_t* __restrict__ members, float* __restrict__ centroids) {
    const int c = blockIdx.x;
    const uint32_t lo = offsets[c], hi = offsets[c + 1];
    if (hi == lo) return;  // empty cluster: the host applies the reference's re-seeding rule
    const float inv = 1.0f / (float)(hi - lo);
    for (int d = threadIdx.x; d < dim; d += blockDim.x) {
        float s = 0.0f;
        for (uint32_t m = lo; m < hi; ++m) s = s + __ldg(rows + (size_t)members[m] * dim + d);
        centroids[(size_t)c * dim + d] = s * inv;
    }
}

static __global__ void ivf_gather_rows_kernel(const float* __restrict__ rows, int dim, const uint32_t* __restrict__ ids, uint32_t n_ids,
                                       float* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint64_t)n_ids * dim) return;
    const uint32_t r = (uint32_t)(i / dim);
    const int d = (int)(i - (uint64_t)r * dim);
    out[i] = __ldg(rows + (size_t)ids[r] * dim + d);
}

// min_ranks[i] = min(min_ranks[i], rank(sample[i], centroid)); then the LAST index of maximal min_rank
// (Iterator::max_by keeps the last of equal maxima) is written to *best.  One CTA of 1024 threads.
static __global__ void __launch_bounds__(1024) ivf_farthest_kernel(const float* __restrict__ sample, uint32_t n, int dim,
                                                            const float* __restrict__ centroid, int metric,
                                                            float* __restrict__ min_ranks, uint32_t* __restrict__ best) {
    __shared__ uint64_t s_key[32];
    const bool vec = (dim & 3) == 0;
    const bool asc = metric_ascending(metric);
    uint64_t key = 0;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const float raw = compute_distance<true>(metric, sample + (size_t)i * dim, centroid, dim, vec);
        const float rank = asc ? raw : -raw;
        float m = min_ranks[i];
        if (rank < m) {
            m = rank;
            min_ranks[i] = m;
        }
        const uint64_t k2 = ((uint64_t)f32_orderable(m + 0.0f) << 32) | i;
        key = k2 > key ? k2 : key;
    }
    key = warp_max_u64(key);
    if ((threadIdx.x & 31) == 0) s_key[threadIdx.x >> 5] = key;
    __syncthreads();
    if (threadIdx.x < 32) {
        key = s_key[threadIdx.x];
        key = warp_max_u64(key);
        if (threadIdx.x == 0) *best = (uint32_t)key;
    }
}

}  // namespace lb
