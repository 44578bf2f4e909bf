// lynse_b200.cu — C ABI of liblynse_b200.so (see include/lynse_b200.h).
//
// Host side of the B200-native distance + top-k path: owns the HBM-resident
// corpus and its per-metric side structures, picks the search plan, launches
// the kernels of lb_scan.cuh / lb_tc.cuh on one stream per index, and moves
// queries / results between host and device.  No PyTorch, no CPU fallback.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <chrono>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "lb_host.cuh"
#include "lb_metrics.cuh"
#include "lb_scan.cuh"
#include "lb_scan2.cuh"
#include "lb_ivf.cuh"

namespace lb {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int status, const std::string& msg) {
    g_last_error = msg;
    return status;
}

}  // namespace lb

using namespace lb;

namespace lb {


// VectorStore::append_encoded_bytes (vector_store.rs:379-445): an append joins the last segment when it fits
// inside the segment target, otherwise it opens a new one; it is never split.
static void account_segment(lb_index* idx, uint64_t n_new) {
    uint64_t bytes = n_new * row_bytes(idx);
    uint64_t target = std::max<uint64_t>(idx->seg_target, row_bytes(idx));
    const bool forced = idx->force_new_segment;
    idx->force_new_segment = false;
    if (!forced && !idx->segments.empty() && idx->segments.back() * row_bytes(idx) + bytes <= target)
        idx->segments.back() += n_new;
    else
        idx->segments.push_back(n_new);
}

static int grow_rows(lb_index* idx, uint64_t n_total) {
    return idx->rows.ensure((size_t)n_total * row_bytes(idx), true, idx->stream);
}

int refresh_small_segments(lb_index* idx) {
    std::vector<uint32_t> ranges;
    uint64_t base = 0;
    size_t sig = idx->segments.size() * 1315423911u;
    for (uint64_t r : idx->segments) {
        if (r < 4096) {
            ranges.push_back((uint32_t)base);
            ranges.push_back((uint32_t)(base + r));
        }
        base += r;
        sig = sig * 31 + (size_t)r;
    }
    if (sig == idx->small_seg_sig) return LB_OK;
    idx->n_small = (int)(ranges.size() / 2);
    if (!ranges.empty()) {
        LB_TRY(idx->small_seg.ensure(ranges.size() * 4));
        LB_CUDA_TRY(cudaMemcpyAsync(idx->small_seg.p, ranges.data(), ranges.size() * 4, cudaMemcpyHostToDevice, idx->stream));
        LB_CUDA_TRY(cudaStreamSynchronize(idx->stream));
    }
    idx->small_seg_sig = sig;
    return LB_OK;
}

// f32 <-> binary16 rows (LB_F16 indexes; the host hands over values that are already binary16-exact)
static __global__ void f32_to_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) out[i] = __float2half_rn(in[i]);
}
static __global__ void f16_to_f32_kernel(const __half* __restrict__ in, float* __restrict__ out, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) out[i] = __half2float(in[i]);
}
static inline bool dense_rows(const lb_index* idx) { return idx->dtype == LB_F32 || idx->dtype == LB_F16; }

// ---- side structures --------------------------------------------------------------------------------------
static int ensure_packed(lb_index* idx) {
    if (!dense_rows(idx)) return LB_OK;
    if (idx->packed_rows == idx->n) return LB_OK;
    int nw = (idx->dim + 63) / 64;
    LB_TRY(idx->packed.ensure((size_t)idx->n * nw * 8, true, idx->stream));
    uint64_t first = idx->packed_rows, cnt = idx->n - first;
    const int warps = 8;
    if (idx->dtype == LB_F16)
        pack_binary_kernel<<<(unsigned)ceil_div(cnt, warps), warps * 32, 0, idx->stream>>>(
            idx->rows.as<__half>() + first * idx->dim, cnt, (int)idx->dim, nw, 0.5f, idx->packed.as<uint64_t>() + first * nw);
    else
        pack_binary_kernel<<<(unsigned)ceil_div(cnt, warps), warps * 32, 0, idx->stream>>>(
            idx->rows.as<float>() + first * idx->dim, cnt, (int)idx->dim, nw, 0.5f, idx->packed.as<uint64_t>() + first * nw);
    LB_CUDA_TRY(cudaGetLastError());
    idx->packed_rows = idx->n;
    return LB_OK;
}

static int ensure_js_stats(lb_index* idx) {
    if (idx->js_rows == idx->n) return LB_OK;
    LB_TRY(idx->js_stats.ensure((size_t)idx->n * 8, true, idx->stream));
    uint64_t first = idx->js_rows, cnt = idx->n - first;
    if (idx->dtype == LB_F16)
        row_stats_kernel<<<(unsigned)ceil_div(cnt, 128), 128, 0, idx->stream>>>(idx->rows.as<__half>() + first * idx->dim, cnt,
                                                                                (int)idx->dim, idx->js_stats.as<float>() + 2 * first);
    else
        row_stats_kernel<<<(unsigned)ceil_div(cnt, 128), 128, 0, idx->stream>>>(idx->rows.as<float>() + first * idx->dim, cnt,
                                                                                (int)idx->dim, idx->js_stats.as<float>() + 2 * first);
    LB_CUDA_TRY(cudaGetLastError());
    idx->js_rows = idx->n;
    return LB_OK;
}

static int ensure_mass_stats(lb_index* idx) {
    if (idx->mass_rows == idx->n) return LB_OK;
    LB_TRY(idx->mass_stats.ensure((size_t)idx->n * 8, true, idx->stream));
    const uint64_t first = idx->mass_rows, cnt = idx->n - first;
    if (idx->dtype == LB_F16)
        row_mass_kernel<<<(unsigned)ceil_div(cnt, 128), 128, 0, idx->stream>>>(idx->rows.as<__half>() + first * idx->dim, cnt, (int)idx->dim,
                                                                               idx->mass_stats.as<double>() + first);
    else
        row_mass_kernel<<<(unsigned)ceil_div(cnt, 128), 128, 0, idx->stream>>>(idx->rows.as<float>() + first * idx->dim, cnt, (int)idx->dim,
                                                                               idx->mass_stats.as<double>() + first);
    LB_CUDA_TRY(cudaGetLastError());
    idx->mass_rows = idx->n;
    return LB_OK;
}


// ---- search on device-resident queries ---------------------------------------------------------------------------------
// d_queries: f32 [nq][dim] (LB_F32 index) or u64 [nq][n_words] (LB_PACKED_U64 index); results [nq][k], k <= n.
enum { SCORE_FLAT = 0, SCORE_PAIRWISE = 1, SCORE_F16_ROWS = 2 };

// A tensor-core pass keeps floor(clusters / query groups) slots of one cluster per query group busy; when the number of
// query groups (256 queries each) divides the 74 cluster slots badly — 16 groups: 64 of 74 — the batch is cut in equal
// sub-batches that divide them better (2 x 8 groups: 72 of 74), each streaming the corpus once.
static int tc_query_split(const lb_index* idx, int nq) {
    const int groups = (nq + 255) / 256, clusters = idx->sm_count / 2;
    if (groups <= 4 || tc_env_int("LYNSE_B200_TC_SPLIT", 1) == 0) return 1;
    int best = 1;
    double best_util = 0.0;
    for (int s = 1; s <= 4; ++s) {
        const int g = (groups + s - 1) / s;
        if (g > clusters) continue;
        // sub-batches of g groups: clusters / g slots of g clusters; the last sub-batch may be smaller
        const double util = (double)(clusters / g * g) / clusters * ((double)groups / (double)(g * s));
        if (util > best_util * 1.03) {
            best_util = util;
            best = s;
        }
    }
    return best;
}

// defer_tc_check: a tensor-core plan leaves its certification flags unread (the caller runs tc_finish after whatever
// it enqueues behind the search); implies no synchronisation here.
static int search_device_impl(lb_index* idx, int metric, const void* d_queries, int nq, int k, const uint64_t* d_allow,
                              uint32_t* d_rows, float* d_dists, uint32_t* d_counts, bool sync_at_end = true, bool defer_tc_check = false) {
    idx->stats = lb_search_stats{};
    if (idx->timing) cudaEventRecord(idx->ev[2], idx->stream);
    int kernels = 0;
    float ms_dom = 0;
    // the tensor-core plan is a candidate; building its shadow (a no-op once built) may still rule it out (non-finite rows)
    const bool tc_candidate = dense_rows(idx) && !metric_binary(metric) && idx->plan == LB_PLAN_AUTO && idx->score_mode == SCORE_FLAT &&
                              tc_supported(idx, metric) && k <= 256 && idx->n >= 64;
    bool use_tc = tc_candidate &&
                  // a row filter rides along as a mask on the hit bits; with few allowed rows the shortlists cannot fill
                  // and certification would send everything to the exact scan anyway
                  (d_allow == nullptr || idx->allow_count >= (uint64_t)std::max(1024, 32 * k)) &&
                  // a handful of queries over a small corpus is a latency case: the exact scan is two launches, the tensor
                  // plan three plus a lazily built shadow (100k x 128, one query: 85 against 131 us of device time)
                  !(nq <= 4 && (uint64_t)idx->n * idx->dim * 4 < (256ull << 20));
    if (use_tc) {
        LB_TRY(ensure_shadow(idx, shadow_kind_for(metric)));
        use_tc = !idx->shadow[shadow_kind_for(metric)].disabled;
    }
    if (idx->dtype == LB_PACKED_U64 || metric_binary(metric)) {
        // FlatMmap::search binary branch (flat_mmap.rs:839-845): packed rows, packed queries
        const uint64_t* words;
        const uint64_t* qwords;
        int nw;
        if (idx->dtype == LB_PACKED_U64) {
            if (!metric_binary(metric)) return fail(LB_INVALID_ARGUMENT, "a packed index only serves hamming/jaccard/tanimoto/dice");
            words = idx->rows.as<uint64_t>();
            qwords = reinterpret_cast<const uint64_t*>(d_queries);
            nw = idx->n_words;
        } else {
            LB_TRY(ensure_packed(idx));
            nw = (idx->dim + 63) / 64;
            LB_TRY(idx->w_qwords.ensure((size_t)nq * nw * 8));
            const int warps = 8;
            pack_binary_kernel<<<(nq + warps - 1) / warps, warps * 32, 0, idx->stream>>>(
                reinterpret_cast<const float*>(d_queries), (uint64_t)nq, (int)idx->dim, nw, 0.5f, idx->w_qwords.as<uint64_t>());
            LB_CUDA_TRY(cudaGetLastError());
            ++kernels;
            words = idx->packed.as<uint64_t>();
            qwords = idx->w_qwords.as<uint64_t>();
        }
        if (idx->plan == LB_PLAN_AUTO && idx->score_mode == SCORE_FLAT && tc_bits_supported(idx, metric, nw, nq, k) &&
            (d_allow == nullptr || idx->allow_count >= (uint64_t)std::max(1024, 32 * k))) {
            // a large batch: the popcounts are a {0,1} contraction on the tensor cores (exact: integer accumulators),
            // shortlist -> exact counts on the packed rows -> certified, as for the dense metrics
            idx->stats.kernels_launched = kernels;
            const int split = tc_query_split(idx, nq);
            if (split == 1) {
                LB_TRY(run_tc_bits(idx, metric, words, nw, qwords, nq, k, d_rows, d_dists, d_counts, d_allow, defer_tc_check));
            } else {
                const int per = ((nq + split - 1) / split + 255) / 256 * 256;
                uint32_t fallbacks = 0;
                float ms_sum = 0;
                for (int q0 = 0; q0 < nq; q0 += per) {
                    const int nqs = std::min(per, nq - q0);
                    LB_TRY(run_tc_bits(idx, metric, words, nw, qwords + (size_t)q0 * nw, nqs, k, d_rows + (size_t)q0 * k, d_dists + (size_t)q0 * k,
                                       d_counts + q0, d_allow, false));
                    fallbacks += idx->stats.n_fallback;
                    ms_sum += idx->stats.ms_dominant;
                }
                idx->stats.n_fallback = fallbacks;
                idx->stats.ms_dominant = ms_sum;
                idx->stats.algorithmic_flops = 2ull * (uint64_t)nq * idx->n * (uint64_t)nw * 64;
            }
            kernels = idx->stats.kernels_launched;
            ms_dom = idx->stats.ms_dominant;
        } else {
            ScanRequest r;
            r.words = words;
            r.n_rows = idx->n;
            r.n_words = nw;
            r.qwords = qwords;
            r.nq = nq;
            r.k = k;
            r.metric = metric;
            r.allow_bits = d_allow;
            r.out_rows = d_rows;
            r.out_dists = d_dists;
            r.out_counts = d_counts;
            LB_TRY(run_scan(idx, r, &kernels, &ms_dom));
            idx->stats.plan_used = 2;
            idx->stats.algorithmic_bytes = (uint64_t)idx->n * nw * 8;
        }
    } else if (use_tc) {
        const int split = tc_query_split(idx, nq);
        if (split == 1) {
            LB_TRY(run_tc(idx, metric, reinterpret_cast<const float*>(d_queries), nq, k, d_rows, d_dists, d_counts, nullptr, d_allow, defer_tc_check));
        } else {
            const int per = ((nq + split - 1) / split + 255) / 256 * 256;
            uint32_t fallbacks = 0;
            float ms_sum = 0;
            for (int q0 = 0; q0 < nq; q0 += per) {
                const int nqs = std::min(per, nq - q0);
                LB_TRY(run_tc(idx, metric, reinterpret_cast<const float*>(d_queries) + (size_t)q0 * idx->dim, nqs, k, d_rows + (size_t)q0 * k,
                              d_dists + (size_t)q0 * k, d_counts + q0, nullptr, d_allow, false));
                fallbacks += idx->stats.n_fallback;
                ms_sum += idx->stats.ms_dominant;
            }
            idx->stats.n_fallback = fallbacks;
            idx->stats.ms_dominant = ms_sum;
            idx->stats.algorithmic_flops = 2ull * (uint64_t)nq * idx->n * idx->dim;
        }
        kernels = idx->stats.kernels_launched;
        ms_dom = idx->stats.ms_dominant;
    } else {
        LB_TRY(refresh_small_segments(idx));
        ScanRequest r;
        if (idx->dtype == LB_F16) r.corpus_h = idx->rows.as<__half>();
        else r.corpus = idx->rows.as<float>();
        r.n_rows = idx->n;
        r.dim = (int)idx->dim;
        r.queries = reinterpret_cast<const float*>(d_queries);
        r.nq = nq;
        r.k = k;
        r.metric = metric;
        r.allow_bits = d_allow;
        r.small_seg = idx->small_seg.as<uint32_t>();
        r.n_small = idx->n_small;
        r.ip_single = idx->score_mode == SCORE_PAIRWISE ? 1 : 0;
        r.f16_rows = idx->score_mode == SCORE_F16_ROWS ? 1 : 0;
        r.out_rows = d_rows;
        r.out_dists = d_dists;
        r.out_counts = d_counts;
        if (metric == LB_WASSERSTEIN && !r.f16_rows) {
            LB_TRY(ensure_mass_stats(idx));
            r.row_mass = idx->mass_stats.as<double>();
        }
        std::vector<uint32_t> unhandled;
        if (metric == LB_JENSEN_SHANNON && idx->score_mode == SCORE_FLAT) {
            // FlatMmap::search Jensen-Shannon branch (the per-pair scoring modes call jensen_shannon_distance directly) (flat_mmap.rs:912-921, :926-1111)
            LB_TRY(ensure_js_stats(idx));
            LB_TRY(idx->w_qstats.ensure((size_t)nq * 8));
            LB_TRY(idx->w_nq.ensure((size_t)nq * idx->dim * 4));
            LB_TRY(idx->w_flags.ensure((size_t)nq * 4 + 16));
            row_stats_kernel<<<(nq + 127) / 128, 128, 0, idx->stream>>>(reinterpret_cast<const float*>(d_queries), (uint64_t)nq,
                                                                        (int)idx->dim, idx->w_qstats.as<float>());
            js_prepare_queries_kernel<<<nq, 128, 0, idx->stream>>>(reinterpret_cast<const float*>(d_queries), nq, (int)idx->dim,
                                                                   idx->w_qstats.as<float>(), idx->w_nq.as<float>(),
                                                                   idx->w_flags.as<uint32_t>() + 4);
            LB_CUDA_TRY(cudaGetLastError());
            kernels += 2;
            unhandled.resize(nq);
            LB_CUDA_TRY(cudaMemcpyAsync(unhandled.data(), idx->w_flags.as<uint32_t>() + 4, (size_t)nq * 4, cudaMemcpyDeviceToHost,
                                        idx->stream));
            LB_CUDA_TRY(cudaStreamSynchronize(idx->stream));
            r.queries = idx->w_nq.as<float>();
            r.row_stats = idx->js_stats.as<float>();
            r.query_stats = idx->w_qstats.as<float>();
            r.sqrt_scores = 1;
        }
        LB_TRY(run_scan(idx, r, &kernels, &ms_dom));
        if (metric == LB_JENSEN_SHANNON && idx->score_mode == SCORE_FLAT) {
            // queries the cached path cannot serve fall back to the direct kernel (prepare_jensen_shannon_query -> None)
            std::vector<uint32_t> qmap;
            for (int q = 0; q < nq; ++q)
                if (unhandled[q]) qmap.push_back((uint32_t)q);
            if (!qmap.empty()) {
                const int ns = (int)qmap.size();
                LB_TRY(idx->w_sub_q.ensure((size_t)ns * idx->dim * 4));
                LB_TRY(idx->w_qmap.ensure((size_t)ns * 4));
                LB_CUDA_TRY(cudaMemcpyAsync(idx->w_qmap.p, qmap.data(), (size_t)ns * 4, cudaMemcpyHostToDevice, idx->stream));
                for (int i = 0; i < ns; ++i)
                    LB_CUDA_TRY(cudaMemcpyAsync(idx->w_sub_q.as<float>() + (size_t)i * idx->dim,
                                                reinterpret_cast<const float*>(d_queries) + (size_t)qmap[i] * idx->dim,
                                                (size_t)idx->dim * 4, cudaMemcpyDeviceToDevice, idx->stream));
                ScanRequest r2 = r;
                r2.queries = idx->w_sub_q.as<float>();
                r2.nq = ns;
                r2.row_stats = nullptr;
                r2.query_stats = nullptr;
                r2.sqrt_scores = 0;
                r2.qmap = idx->w_qmap.as<uint32_t>();
                LB_TRY(run_scan(idx, r2, &kernels, nullptr));
            }
        }
        idx->stats.plan_used = 0;
        idx->stats.algorithmic_bytes = (uint64_t)idx->n * row_bytes(idx);
    }
    if (idx->timing) cudaEventRecord(idx->ev[3], idx->stream);
    if ((sync_at_end || idx->timing) && !idx->pending_tc.active) LB_CUDA_TRY(cudaStreamSynchronize(idx->stream));  // (the host path syncs after its copies)
    idx->stats.kernels_launched = kernels;
    idx->stats.ms_dominant = ms_dom;
    if (idx->timing && !idx->pending_tc.active) {
        float ms = 0;
        cudaEventElapsedTime(&ms, idx->ev[2], idx->ev[3]);
        idx->stats.ms_total = ms;
    }
    return LB_OK;
}

static int check_metric(int metric) {
    if (metric < 0 || metric >= LB_METRIC_COUNT) return fail(LB_INVALID_ARGUMENT, "Unknown metric: " + std::to_string(metric));
    return LB_OK;
}

constexpr int QUERY_BATCH = 4096;
constexpr int MAX_K = 2048;

}  // namespace lb

// ============================================ C ABI ===================================================
extern "C" {

const char* lb_last_error(void) { return g_last_error.c_str(); }
const char* lb_version(void) { return "lynse_b200 0.1.0 (sm_100a)"; }

int lb_device_count(int* out) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *out = 0;
        return fail(LB_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    }
    *out = n;
    return LB_OK;
}

int lb_device_info(int device, char* name, int cap, uint64_t* total_bytes, uint64_t* free_bytes, int* sm_count) {
    cudaDeviceProp prop;
    LB_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (name && cap > 0) {
        strncpy(name, prop.name, (size_t)cap - 1);
        name[cap - 1] = 0;
    }
    DeviceGuard g(device);
    size_t fr = 0, tot = 0;
    LB_CUDA_TRY(cudaMemGetInfo(&fr, &tot));
    if (total_bytes) *total_bytes = tot;
    if (free_bytes) *free_bytes = fr;
    if (sm_count) *sm_count = prop.multiProcessorCount;
    return LB_OK;
}

int lb_index_create(lb_index** out, uint32_t dim, int dtype, int device) {
    if (!out) return fail(LB_INVALID_ARGUMENT, "out is null");
    if (dim == 0) return fail(LB_INVALID_ARGUMENT, "dimension must be positive");
    if (dtype != LB_F32 && dtype != LB_PACKED_U64 && dtype != LB_F16) return fail(LB_INVALID_ARGUMENT, "unknown dtype");
    int ndev = 0;
    LB_CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(LB_INVALID_ARGUMENT, "no such CUDA device: " + std::to_string(device));
    cudaDeviceProp prop;
    LB_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(LB_CUDA, std::string("liblynse_b200 is built for sm_100a only; device is ") + prop.name + " (sm_" +
                                 std::to_string(prop.major) + std::to_string(prop.minor) + ")");
    DeviceGuard g(device);
    lb_index* idx = new lb_index();
    idx->device = device;
    idx->dim = dim;
    idx->dtype = dtype;
    idx->n_words = (int)((dim + 63) / 64);
    idx->sm_count = prop.multiProcessorCount;
    cudaError_t e = cudaStreamCreateWithFlags(&idx->stream, cudaStreamNonBlocking);
    for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreate(&idx->ev[i]);
    for (int i = 0; i < 8 && e == cudaSuccess; ++i) e = cudaEventCreate(&idx->user_ev[i]);
    if (e != cudaSuccess) {
        delete idx;
        return fail(LB_CUDA, std::string("stream/event creation: ") + cudaGetErrorString(e));
    }
    *out = idx;
    return LB_OK;
}

void lb_index_destroy(lb_index* idx) {
    if (!idx) return;
    {
        DeviceGuard g(idx->device);
        cudaStreamSynchronize(idx->stream);
        DevBuf* bufs[] = {&idx->rows, &idx->packed, &idx->js_stats, &idx->mass_stats, &idx->max_norm, &idx->small_seg, &idx->w_queries,
                          &idx->w_qwords, &idx->w_allow, &idx->w_lists, &idx->w_counts, &idx->w_thr, &idx->w_out_rows,
                          &idx->w_out_dists, &idx->w_out_counts, &idx->w_out, &idx->w_qb, &idx->w_qnorm, &idx->w_cand_score,
                          &idx->w_cand_row, &idx->w_cand_thr, &idx->w_flags, &idx->w_qstats, &idx->w_nq, &idx->w_sub_q,
                          &idx->w_qmap, &idx->shadow[0].buf, &idx->shadow[1].buf, &idx->shadow[2].buf,
                          &idx->w_send, &idx->w_recv, &idx->w_g_rows, &idx->w_g_dists, &idx->w_g_counts, &idx->w_progress, &idx->w_prof, &idx->w_gfloor};
        for (DevBuf* b : bufs) b->release();
        idx->h_in.release();
        idx->h_out.release();
        idx->h_tails.release();
        for (int i = 0; i < 4; ++i)
            if (idx->ev[i]) cudaEventDestroy(idx->ev[i]);
        for (int i = 0; i < 8; ++i)
            if (idx->user_ev[i]) cudaEventDestroy(idx->user_ev[i]);
        cudaStreamDestroy(idx->stream);
    }
    delete idx;
}

int lb_index_reserve(lb_index* idx, uint64_t n_rows) {
    if (!idx) return fail(LB_INVALID_ARGUMENT, "index is null");
    std::lock_guard<std::mutex> lock(idx->mu);
    DeviceGuard g(idx->device);
    if (n_rows >= 0xFFFFFFFFull) return fail(LB_INVALID_ARGUMENT, "an index holds fewer than 2^32-1 rows");
    return grow_rows(idx, n_rows);
}

int lb_index_new_segment(lb_index* idx) {
    if (!idx) return fail(LB_INVALID_ARGUMENT, "index is null");
    std::lock_guard<std::mutex> lock(idx->mu);
    idx->force_new_segment = true;
    return LB_OK;
}

int lb_index_set_segment_target(lb_index* idx, uint64_t bytes) {
    if (!idx) return fail(LB_INVALID_ARGUMENT, "index is null");
    std::lock_guard<std::mutex> lock(idx->mu);
    idx->seg_target = bytes ? bytes : 256ull * 1024 * 1024;
    return LB_OK;
}

static int append_common(lb_index* idx, const void* host, uint64_t n) {
    if (n == 0) return LB_OK;
    if (idx->n + n >= 0xFFFFFFFFull) return fail(LB_INVALID_ARGUMENT, "an index holds fewer than 2^32-1 rows");
    LB_TRY(grow_rows(idx, idx->n + n));
    size_t rb = row_bytes(idx);
    LB_CUDA_TRY(cudaMemcpyAsync(reinterpret_cast<char*>(idx->rows.p) + idx->n * rb, host, n * rb, cudaMemcpyHostToDevice, idx->stream));
    LB_CUDA_TRY(cudaStreamSynchronize(idx->stream));
    account_segment(idx, n);
    idx->n += n;
    return LB_OK;
}

int lb_index_append_f32(lb_index* idx, const float* rows, uint64_t n) {
    if (!idx) return fail(LB_INVALID_ARGUMENT, "index is null");
    if (!dense_rows(idx)) return fail(LB_INVALID_ARGUMENT, "index does not store f32 rows");
    if (!rows && n) return fail(LB_INVALID_ARGUMENT, "rows is null");
    std::lock_guard<std::mutex> lock(idx->mu);
    DeviceGuard g(idx->device);
    if (idx->dtype == LB_F32) return append_common(idx, rows, n);
    // binary16 index: the values go through an f32 staging buffer and are narrowed on the device (round to nearest even,
    // src/storage/dtype.rs:60-67; exact for values that already are binary16)
    if (n == 0) return LB_OK;
    if (idx->n + n >= 0xFFFFFFFFull) return fail(LB_INVALID_ARGUMENT, "an index holds fewer than 2^32-1 rows");
    LB_TRY(grow_rows(idx, idx->n + n));
    const uint64_t chunk_rows = std::max<uint64_t>(1, (64ull << 20) / ((uint64_t)idx->dim * 4));
    for (uint64_t done = 0; done < n; done += chunk_rows) {
        const uint64_t m = std::min(chunk_rows, n - done), ne = m * idx->dim;
        LB_TRY(idx->w_queries.ensure(ne * 4));
        LB_CUDA_TRY(cudaMemcpyAsync(idx->w_queries.p, rows + done * idx->dim, ne * 4, cudaMemcpyHostToDevice, idx->stream));
        f32_to_f16_kernel<<<(unsigned)idx->sm_count * 8, 256, 0, idx->stream>>>(idx->w_queries.as<float>(),
                                                                              idx->rows.as<__half>() + (idx->n + done) * idx->dim, ne);
        LB_CUDA_TRY(cudaGetLastError());
        LB_CUDA_TRY(cudaStreamSynchronize(idx->stream));
    }
    account_segment(idx, n);
    idx->n += n;
    return LB_OK;
}

int lb_index_append_f16(lb_index* idx, const uint16_t* rows, uint64_t n) {
    if (!idx) return fail(LB_INVALID_ARGUMENT, "index is null");
    if (idx->dtype != LB_F16) return fail(LB_INVALID_ARGUMENT, "index does not store binary16 rows");
    if (!rows && n) return fail(LB_INVALID_ARGUMENT, "rows is null");
    std::lock_guard<std::mutex> lock(idx->mu);
    DeviceGuard g(idx->device);
    return append_common(idx, rows, n);
}

int lb_index_append_packed(lb_index* idx, const uint64_t* words, uint64_t n) {
    if (!idx) return fail(LB_INVALID_ARGUMENT, "index is null");
    if (idx->dtype != LB_PACKED_U64) return fail(LB_INVALID_ARGUMENT, "index does not store packed rows");
    if (!words && n) return fail(LB_INVALID_ARGUMENT, "words is null");
    std::lock_guard<std::mutex> lock(idx->mu);
    DeviceGuard g(idx->device);
    return append_common(idx, words, n);
}

int lb_index_append_synthetic(lb_index* idx, uint64_t n, uint64_t seed, uint64_t row_offset) {
    if (!idx) return fail(LB_INVALID_ARGUMENT, "index is null");
    std::lock_guard<std::mutex> lock(idx->mu);
    DeviceGuard g(idx->device);
    if (n == 0) return LB_OK;
    if (idx->n + n >= 0xFFFFFFFFull) return fail(LB_INVALID_ARGUMENT, "an index holds fewer than 2^32-1 rows");
    LB_TRY(grow_rows(idx, idx->n + n));
    const unsigned blocks = (unsigned)idx->sm_count * 8;
    if (idx->dtype == LB_F32) {
        uint64_t ne = n * idx->dim;
        synth_f32_kernel<<<blocks, 256, 0, idx->stream>>>(idx->rows.as<float>() + idx->n * idx->dim, ne, seed, row_offset * idx->dim);
    } else if (idx->dtype == LB_F16) {
        // the f32 synthetic values narrowed to binary16 (host: synthetic.rows_f32(...).astype(float16)), in pieces of 64 MiB
        const uint64_t chunk_rows = std::max<uint64_t>(1, (64ull << 20) / ((uint64_t)idx->dim * 4));
        for (uint64_t done = 0; done < n; done += chunk_rows) {
            const uint64_t m = std::min(chunk_rows, n - done), ne = m * idx->dim;
            LB_TRY(idx->w_queries.ensure(ne * 4));
            synth_f32_kernel<<<blocks, 256, 0, idx->stream>>>(idx->w_queries.as<float>(), ne, seed, (row_offset + done) * idx->dim);
            f32_to_f16_kernel<<<blocks, 256, 0, idx->stream>>>(idx->w_queries.as<float>(), idx->rows.as<__half>() + (idx->n + done) * idx->dim, ne);
        }
    } else {
        uint64_t ne = n * (uint64_t)idx->n_words;
        synth_u64_kernel<<<blocks, 256, 0, idx->stream>>>(idx->rows.as<uint64_t>() + idx->n * idx->n_words, ne, seed,
                                                         row_offset * (uint64_t)idx->n_words);
    }
    LB_CUDA_TRY(cudaGetLastError());
    LB_CUDA_TRY(cudaStreamSynchronize(idx->stream));
    account_segment(idx, n);
    idx->n += n;
    return LB_OK;
}

uint64_t lb_index_len(const lb_index* idx) { return idx ? idx->n : 0; }
uint32_t lb_index_dim(const lb_index* idx) { return idx ? idx->dim : 0; }

int lb_index_segments(const lb_index* idx, uint64_t* rows_out, int cap, int* n_segments) {
    if (!idx) return fail(LB_INVALID_ARGUMENT, "index is null");
    if (n_segments) *n_segments = (int)idx->segments.size();
    for (int i = 0; i < cap && i < (int)idx->segments.size(); ++i) rows_out[i] = idx->segments[i];
    return LB_OK;
}

int lb_index_read_rows_f32(lb_index* idx, uint64_t first, uint64_t n, float* out) {
    if (!idx || !out) return fail(LB_INVALID_ARGUMENT, "null argument");
    if (!dense_rows(idx)) return fail(LB_INVALID_ARGUMENT, "index does not store f32 rows");
    std::lock_guard<std::mutex> lock(idx->mu);
    DeviceGuard g(idx->device);
    if (first + n > idx->n) return fail(LB_INVALID_ARGUMENT, "row range out of bounds");
    if (idx->dtype == LB_F16) {  // decoded on the device, in pieces of 64 MiB
        const uint64_t chunk_rows = std::max<uint64_t>(1, (64ull << 20) / ((uint64_t)idx->dim * 4));
        for (uint64_t done = 0; done < n; done += chunk_rows) {
            const uint64_t m = std::min(chunk_rows, n - done), ne = m * idx->dim;
            LB_TRY(idx->w_sub_q.ensure(ne * 4));
            f16_to_f32_kernel<<<(unsigned)idx->sm_count * 8, 256, 0, idx->stream>>>(idx->rows.as<__half>() + (first + done) * idx->dim,
                                                                                  idx->w_sub_q.as<float>(), ne);
            LB_CUDA_TRY(cudaGetLastError());
            LB_CUDA_TRY(cudaMemcpyAsync(out + done * idx->dim, idx->w_sub_q.p, ne * 4, cudaMemcpyDeviceToHost, idx->stream));
            LB_CUDA_TRY(cudaStreamSynchronize(idx->stream));
        }
        return LB_OK;
    }
    LB_CUDA_TRY(cudaMemcpyAsync(out, idx->rows.as<float>() + first * idx->dim, n * idx->dim * 4, cudaMemcpyDeviceToHost, idx->stream));
    LB_CUDA_TRY(cudaStreamSynchronize(idx->stream));
    return LB_OK;
}

int lb_index_prepare(lb_index* idx, int metric) {
    if (!idx) return fail(LB_INVALID_ARGUMENT, "index is null");
    LB_TRY(check_metric(metric));
    std::lock_guard<std::mutex> lock(idx->mu);
    DeviceGuard g(idx->device);
    if (idx->n == 0) return LB_OK;
    if (metric_binary(metric)) {
        LB_TRY(ensure_packed(idx));
    } else if (idx->dtype == LB_PACKED_U64) {
        return fail(LB_INVALID_ARGUMENT, "a packed index only serves hamming/jaccard/tanimoto/dice");
    } else if (metric == LB_JENSEN_SHANNON) {
        LB_TRY(ensure_js_stats(idx));
    } else if (idx->plan == LB_PLAN_AUTO && tc_supported(idx, metric)) {
        LB_TRY(ensure_shadow(idx, shadow_kind_for(metric)));
    }
    LB_CUDA_TRY(cudaStreamSynchronize(idx->stream));
    return LB_OK;
}

int lb_index_set_plan(lb_index* idx, int plan) {
    if (!idx) return fail(LB_INVALID_ARGUMENT, "index is null");
    if (plan != LB_PLAN_AUTO && plan != LB_PLAN_EXACT) return fail(LB_INVALID_ARGUMENT, "unknown plan");
    std::lock_guard<std::mutex> lock(idx->mu);
    idx->plan = plan;
    return LB_OK;
}

int lb_index_set_timing(lb_index* idx, int enabled) {
    if (!idx) return fail(LB_INVALID_ARGUMENT, "index is null");
    std::lock_guard<std::mutex> lock(idx->mu);
    idx->timing = enabled != 0;
    return LB_OK;
}

int lb_index_last_stats(const lb_index* idx, lb_search_stats* out) {
    if (!idx || !out) return fail(LB_INVALID_ARGUMENT, "null argument");
    *out = idx->stats;
    return LB_OK;
}

struct ScoreModeScope {  // the caller holds idx->mu
    lb_index* idx;
    ScoreModeScope(lb_index* i, int mode) : idx(i) { idx->score_mode = mode; }
    ~ScoreModeScope() { idx->score_mode = SCORE_FLAT; }
};

static int search_host_common(lb_index* idx, int score_mode, int metric, const void* queries, size_t query_row_bytes, uint32_t nq, uint32_t k,
                              const uint64_t* allow_bits, uint64_t allow_words, uint32_t* out_rows, float* out_dists,
                              uint32_t* out_counts) {
    LB_TRY(check_metric(metric));
    if (nq && !queries) return fail(LB_INVALID_ARGUMENT, "queries is null");
    if (k > (uint32_t)MAX_K) return fail(LB_UNSUPPORTED, "k above 2048 is not supported");
    if (metric == LB_HAVERSINE && idx->dim != 2) return fail(LB_INVALID_ARGUMENT, "haversine requires dimension 2");
    std::lock_guard<std::mutex> lock(idx->mu);
    ScoreModeScope mode_scope(idx, score_mode);
    DeviceGuard g(idx->device);
    const uint32_t kk = (uint32_t)std::min<uint64_t>(k, idx->n);  // k.min(n) (flat_mmap.rs:836)
    for (uint32_t q = 0; q < nq; ++q) out_counts[q] = 0;
    for (size_t i = 0; i < (size_t)nq * k; ++i) {
        out_rows[i] = ROW_NONE;
        out_dists[i] = NAN;
    }
    if (nq == 0 || kk == 0) return LB_OK;  // k == 0 or empty index -> empty result (flat_mmap.rs:833-835)
    const uint64_t* d_allow = nullptr;
    if (allow_bits) {
        uint64_t need = (idx->n + 63) / 64;
        if (allow_words < need) return fail(LB_INVALID_ARGUMENT, "row filter is shorter than the index");
        LB_TRY(idx->w_allow.ensure(need * 8));
        LB_CUDA_TRY(cudaMemcpyAsync(idx->w_allow.p, allow_bits, need * 8, cudaMemcpyHostToDevice, idx->stream));
        d_allow = idx->w_allow.as<uint64_t>();
        uint64_t allowed = 0;  // plan selection only (bits past the last row do not matter at this precision)
        for (uint64_t w = 0; w < need; ++w) allowed += (uint64_t)__builtin_popcountll(allow_bits[w]);
        idx->allow_count = allowed;
    }
    lb_search_stats acc{};
    constexpr size_t STAGE_LIMIT = 256 * 1024;  // batches up to this many bytes go through pinned staging
    for (uint32_t q0 = 0; q0 < nq; q0 += QUERY_BATCH) {
        const int nb = (int)std::min<uint32_t>(QUERY_BATCH, nq - q0);
        const size_t in_bytes = (size_t)nb * query_row_bytes;
        const size_t rows_bytes = (size_t)nb * kk * 4, out_bytes = 2 * rows_bytes + (size_t)nb * 4;
        LB_TRY(idx->w_queries.ensure(in_bytes));
        LB_TRY(idx->w_out.ensure(out_bytes));
        uint32_t* d_rows = idx->w_out.as<uint32_t>();
        float* d_dists = reinterpret_cast<float*>(idx->w_out.as<char>() + rows_bytes);
        uint32_t* d_counts = reinterpret_cast<uint32_t*>(idx->w_out.as<char>() + 2 * rows_bytes);
        const char* src = reinterpret_cast<const char*>(queries) + (size_t)q0 * query_row_bytes;
        if (in_bytes <= STAGE_LIMIT) {
            LB_TRY(idx->h_in.ensure(in_bytes));
            memcpy(idx->h_in.p, src, in_bytes);
            src = reinterpret_cast<const char*>(idx->h_in.p);
        }
        LB_CUDA_TRY(cudaMemcpyAsync(idx->w_queries.p, src, in_bytes, cudaMemcpyHostToDevice, idx->stream));
        // a tensor-core plan leaves its certification flags unread: they come back with the results (one wait per batch),
        // and only a batch with uncertified queries — re-run by the exact scan inside tc_finish — copies its results again
        LB_TRY(search_device_impl(idx, metric, idx->w_queries.p, nb, (int)kk, d_allow, d_rows, d_dists, d_counts, false, true));
        const bool flags_pending = idx->pending_tc.active;
        uint32_t* flag_head = nullptr;
        if (flags_pending) {
            LB_TRY(idx->h_tails.ensure(16));
            flag_head = reinterpret_cast<uint32_t*>(idx->h_tails.p);
        }
        auto finish_flags = [&](bool* changed) -> int {
            *changed = false;
            if (!flags_pending) return LB_OK;
            return tc_finish(idx, changed, flag_head);
        };
        if (out_bytes <= STAGE_LIMIT) {
            LB_TRY(idx->h_out.ensure(out_bytes));
            LB_CUDA_TRY(cudaMemcpyAsync(idx->h_out.p, idx->w_out.p, out_bytes, cudaMemcpyDeviceToHost, idx->stream));
            if (flags_pending) LB_CUDA_TRY(cudaMemcpyAsync(flag_head, idx->w_flags.p, 16, cudaMemcpyDeviceToHost, idx->stream));
            LB_CUDA_TRY(cudaStreamSynchronize(idx->stream));
            bool changed = false;
            LB_TRY(finish_flags(&changed));
            if (changed) {
                LB_CUDA_TRY(cudaMemcpyAsync(idx->h_out.p, idx->w_out.p, out_bytes, cudaMemcpyDeviceToHost, idx->stream));
                LB_CUDA_TRY(cudaStreamSynchronize(idx->stream));
            }
            const char* h = reinterpret_cast<const char*>(idx->h_out.p);
            for (int q = 0; q < nb; ++q) {
                memcpy(out_rows + (size_t)(q0 + q) * k, h + (size_t)q * kk * 4, (size_t)kk * 4);
                memcpy(out_dists + (size_t)(q0 + q) * k, h + rows_bytes + (size_t)q * kk * 4, (size_t)kk * 4);
            }
            memcpy(out_counts + q0, h + 2 * rows_bytes, (size_t)nb * 4);
        } else {
            LB_CUDA_TRY(cudaMemcpy2DAsync(out_rows + (size_t)q0 * k, (size_t)k * 4, d_rows, (size_t)kk * 4, (size_t)kk * 4, nb,
                                          cudaMemcpyDeviceToHost, idx->stream));
            LB_CUDA_TRY(cudaMemcpy2DAsync(out_dists + (size_t)q0 * k, (size_t)k * 4, d_dists, (size_t)kk * 4, (size_t)kk * 4, nb,
                                          cudaMemcpyDeviceToHost, idx->stream));
            LB_CUDA_TRY(cudaMemcpyAsync(out_counts + q0, d_counts, (size_t)nb * 4, cudaMemcpyDeviceToHost, idx->stream));
            if (flags_pending) LB_CUDA_TRY(cudaMemcpyAsync(flag_head, idx->w_flags.p, 16, cudaMemcpyDeviceToHost, idx->stream));
            LB_CUDA_TRY(cudaStreamSynchronize(idx->stream));
            bool changed = false;
            LB_TRY(finish_flags(&changed));
            if (changed) {
                LB_CUDA_TRY(cudaMemcpy2DAsync(out_rows + (size_t)q0 * k, (size_t)k * 4, d_rows, (size_t)kk * 4, (size_t)kk * 4, nb,
                                              cudaMemcpyDeviceToHost, idx->stream));
                LB_CUDA_TRY(cudaMemcpy2DAsync(out_dists + (size_t)q0 * k, (size_t)k * 4, d_dists, (size_t)kk * 4, (size_t)kk * 4, nb,
                                              cudaMemcpyDeviceToHost, idx->stream));
                LB_CUDA_TRY(cudaMemcpyAsync(out_counts + q0, d_counts, (size_t)nb * 4, cudaMemcpyDeviceToHost, idx->stream));
                LB_CUDA_TRY(cudaStreamSynchronize(idx->stream));
            }
        }
        acc.plan_used = idx->stats.plan_used;
        acc.n_fallback += idx->stats.n_fallback;
        acc.n_partitions = idx->stats.n_partitions;
        acc.kernels_launched += idx->stats.kernels_launched;
        acc.ms_dominant += idx->stats.ms_dominant;
        acc.ms_total += idx->stats.ms_total;
        acc.algorithmic_bytes += idx->stats.algorithmic_bytes;
        acc.algorithmic_flops += idx->stats.algorithmic_flops;
        acc.coarse_operand = idx->stats.coarse_operand;
        acc.coarse_hit_mode = idx->stats.coarse_hit_mode;
        acc.coarse_sm_mhz = idx->stats.coarse_sm_mhz;
    }
    idx->stats = acc;
    return LB_OK;
}

int lb_index_search(lb_index* idx, int metric, const float* queries, uint32_t nq, uint32_t k, const uint64_t* allow_bits,
                    uint64_t allow_words, uint32_t* out_rows, float* out_dists, uint32_t* out_counts) {
    if (!idx) return fail(LB_INVALID_ARGUMENT, "index is null");
    if (!dense_rows(idx)) return fail(LB_INVALID_ARGUMENT, "use lb_index_search_packed for a packed index");
    if (!out_rows || !out_dists || !out_counts) return fail(LB_INVALID_ARGUMENT, "output buffer is null");
    return search_host_common(idx, SCORE_FLAT, metric, queries, (size_t)idx->dim * 4, nq, k, allow_bits, allow_words, out_rows,
                              out_dists, out_counts);
}

int lb_index_search_pairwise(lb_index* idx, int metric, const float* queries, uint32_t nq, uint32_t k, const uint64_t* allow_bits,
                             uint64_t allow_words, uint32_t* out_rows, float* out_dists, uint32_t* out_counts) {
    if (!idx) return fail(LB_INVALID_ARGUMENT, "index is null");
    if (!dense_rows(idx)) return fail(LB_INVALID_ARGUMENT, "pairwise search needs f32 rows");
    if (!out_rows || !out_dists || !out_counts) return fail(LB_INVALID_ARGUMENT, "output buffer is null");
    return search_host_common(idx, SCORE_PAIRWISE, metric, queries, (size_t)idx->dim * 4, nq, k, allow_bits, allow_words, out_rows,
                              out_dists, out_counts);
}

int lb_index_search_f16_rows(lb_index* idx, int metric, const float* queries, uint32_t nq, uint32_t k, const uint64_t* allow_bits,
                             uint64_t allow_words, uint32_t* out_rows, float* out_dists, uint32_t* out_counts) {
    if (!idx) return fail(LB_INVALID_ARGUMENT, "index is null");
    if (!dense_rows(idx)) return fail(LB_INVALID_ARGUMENT, "the float16-row search needs (decoded) f32 rows");
    if (!out_rows || !out_dists || !out_counts) return fail(LB_INVALID_ARGUMENT, "output buffer is null");
    // the binary metrics go through the packed cache whatever the storage dtype (flat_mmap.rs:839-845, :504-510)
    const int mode = (metric >= 0 && metric < LB_METRIC_COUNT && metric_binary(metric)) ? SCORE_FLAT : SCORE_F16_ROWS;
    return search_host_common(idx, mode, metric, queries, (size_t)idx->dim * 4, nq, k, allow_bits, allow_words, out_rows, out_dists,
                              out_counts);
}

int lb_index_search_packed(lb_index* idx, int metric, const uint64_t* query_words, uint32_t nq, uint32_t k, uint32_t* out_rows,
                           float* out_dists, uint32_t* out_counts) {
    if (!idx) return fail(LB_INVALID_ARGUMENT, "index is null");
    if (idx->dtype != LB_PACKED_U64) return fail(LB_INVALID_ARGUMENT, "index does not store packed rows");
    if (!out_rows || !out_dists || !out_counts) return fail(LB_INVALID_ARGUMENT, "output buffer is null");
    if (metric >= 0 && metric < LB_METRIC_COUNT && !metric_binary(metric))
        return fail(LB_INVALID_ARGUMENT, "a packed index only serves hamming/jaccard/tanimoto/dice");
    return search_host_common(idx, SCORE_FLAT, metric, query_words, (size_t)idx->n_words * 8, nq, k, nullptr, 0, out_rows, out_dists,
                              out_counts);
}

int lb_index_search_device(lb_index* idx, int metric, const void* d_queries, uint32_t nq, uint32_t k, uint32_t* d_out_rows,
                           float* d_out_dists, uint32_t* d_out_counts) {
    if (!idx) return fail(LB_INVALID_ARGUMENT, "index is null");
    LB_TRY(check_metric(metric));
    if (!d_queries || !d_out_rows || !d_out_dists || !d_out_counts) return fail(LB_INVALID_ARGUMENT, "null device pointer");
    if (k == 0 || k > (uint32_t)MAX_K || k > idx->n) return fail(LB_INVALID_ARGUMENT, "k must be in [1, min(n, 2048)] for the device-resident entry point");
    if (nq == 0 || nq > (uint32_t)QUERY_BATCH) return fail(LB_INVALID_ARGUMENT, "nq must be in [1, 4096] for the device-resident entry point");
    std::lock_guard<std::mutex> lock(idx->mu);
    DeviceGuard g(idx->device);
    return search_device_impl(idx, metric, d_queries, (int)nq, (int)k, nullptr, d_out_rows, d_out_dists, d_out_counts);
}

// ---- stateless operators ------------------------------------------------------------------------------------------------
int lb_compute_distance(const float* a, const float* b, uint32_t dim, int metric, float* out) {
    LB_TRY(check_metric(metric));
    if (!a || !b || !out) return fail(LB_INVALID_ARGUMENT, "null argument");
    if (metric == LB_HAVERSINE && dim != 2)
        return fail(LB_INVALID_ARGUMENT, "haversine requires two values: longitude and latitude in degrees");
    float* d = nullptr;
    LB_CUDA_TRY(cudaMalloc(&d, ((size_t)2 * dim + 4) * 4));
    cudaError_t e = cudaMemcpy(d, a, (size_t)dim * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d + dim, b, (size_t)dim * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        // b is read through 16-byte loads when dim % 4 == 0: keep both operands 16-byte aligned
        pair_distance_kernel<<<1, 32>>>(d, d + dim, (int)dim, metric, d + 2 * (size_t)dim);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out, d + 2 * (size_t)dim, 4, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(LB_CUDA, std::string("lb_compute_distance: ") + cudaGetErrorString(e));
    return LB_OK;
}

int lb_top_k_search(const float* query, const float* candidates, uint64_t n, uint32_t dim, uint32_t k, int metric, uint32_t* ids,
                    float* dists, uint32_t* out_count) {
    LB_TRY(check_metric(metric));
    if (!out_count) return fail(LB_INVALID_ARGUMENT, "out_count is null");
    *out_count = 0;
    if (dim == 0) return fail(LB_INVALID_ARGUMENT, "dimension must be positive");
    if (n == 0 || k == 0) return LB_OK;  // distance/mod.rs:381-384
    if (!query || !candidates || !ids || !dists) return fail(LB_INVALID_ARGUMENT, "null argument");
    if (metric == LB_HAVERSINE && dim != 2)
        return fail(LB_INVALID_ARGUMENT, "haversine requires two values: longitude and latitude in degrees");
    if (n >= 0xFFFFFFFFull) return fail(LB_INVALID_ARGUMENT, "too many candidates");
    int device = 0;
    LB_CUDA_TRY(cudaGetDevice(&device));
    lb_index* idx = nullptr;
    LB_TRY(lb_index_create(&idx, dim, LB_F32, device));
    int st = lb_index_append_f32(idx, candidates, n);
    if (st == LB_OK) {
        std::lock_guard<std::mutex> lock(idx->mu);
        DeviceGuard g(idx->device);
        const uint32_t kk = (uint32_t)std::min<uint64_t>(std::min<uint64_t>(k, n), MAX_K);
        st = (k > (uint32_t)MAX_K && n > (uint64_t)MAX_K) ? fail(LB_UNSUPPORTED, "k above 2048 is not supported") : LB_OK;
        if (st == LB_OK) st = idx->w_queries.ensure((size_t)dim * 4);
        if (st == LB_OK) st = idx->w_out_rows.ensure((size_t)kk * 4);
        if (st == LB_OK) st = idx->w_out_dists.ensure((size_t)kk * 4);
        if (st == LB_OK) st = idx->w_out_counts.ensure(4);
        if (st == LB_OK) {
            cudaError_t e = cudaMemcpyAsync(idx->w_queries.p, query, (size_t)dim * 4, cudaMemcpyHostToDevice, idx->stream);
            if (e != cudaSuccess) st = fail(LB_CUDA, cudaGetErrorString(e));
        }
        if (st == LB_OK) {
            ScanRequest r;
            r.corpus = idx->rows.as<float>();
            r.n_rows = n;
            r.dim = (int)dim;
            r.queries = idx->w_queries.as<float>();
            r.nq = 1;
            r.k = (int)kk;
            r.metric = metric;
            r.ip_single = 1;  // compute_distance_f32 -> inner_product_f32 (single-row kernel) for every candidate
            r.out_rows = idx->w_out_rows.as<uint32_t>();
            r.out_dists = idx->w_out_dists.as<float>();
            r.out_counts = idx->w_out_counts.as<uint32_t>();
            st = run_scan(idx, r, nullptr, nullptr);
        }
        if (st == LB_OK) {
            cudaError_t e = cudaMemcpyAsync(ids, idx->w_out_rows.p, (size_t)kk * 4, cudaMemcpyDeviceToHost, idx->stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(dists, idx->w_out_dists.p, (size_t)kk * 4, cudaMemcpyDeviceToHost, idx->stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(out_count, idx->w_out_counts.p, 4, cudaMemcpyDeviceToHost, idx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(idx->stream);
            if (e != cudaSuccess) st = fail(LB_CUDA, std::string("lb_top_k_search: ") + cudaGetErrorString(e));
        }
    }
    std::string keep = g_last_error;
    lb_index_destroy(idx);
    g_last_error = keep;
    return st;
}

// ---- memory helpers ---------------------------------------------------------------------------------------------------------
int lb_device_malloc(int device, uint64_t bytes, void** out) {
    DeviceGuard g(device);
    LB_CUDA_TRY(cudaMalloc(out, bytes));
    return LB_OK;
}
int lb_device_free(int device, void* p) {
    DeviceGuard g(device);
    LB_CUDA_TRY(cudaFree(p));
    return LB_OK;
}
int lb_host_malloc(uint64_t bytes, void** out) {
    LB_CUDA_TRY(cudaMallocHost(out, bytes));
    return LB_OK;
}
int lb_host_free(void* p) {
    LB_CUDA_TRY(cudaFreeHost(p));
    return LB_OK;
}
int lb_memcpy_h2d(int device, void* dst, const void* src, uint64_t bytes) {
    DeviceGuard g(device);
    LB_CUDA_TRY(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return LB_OK;
}
int lb_memcpy_d2h(int device, void* dst, const void* src, uint64_t bytes) {
    DeviceGuard g(device);
    LB_CUDA_TRY(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return LB_OK;
}
int lb_device_synchronize(int device) {
    DeviceGuard g(device);
    LB_CUDA_TRY(cudaDeviceSynchronize());
    return LB_OK;
}
int lb_device_memset(int device, void* dst, int value, uint64_t bytes) {
    DeviceGuard g(device);
    LB_CUDA_TRY(cudaMemset(dst, value, bytes));
    return LB_OK;
}

}  // extern "C"

// ---- IVF: k-means build + probe / gather / score search (src/index/ivf.rs, src/index/kmeans.rs) ---------------------------------
struct lb_ivf {
    lb_index* idx = nullptr;     // the rows (not owned)
    lb_index* cidx = nullptr;    // centroids as a tiny index (owned): centroid ranking is one more exact scan
    int metric = 0, routing = 0;
    uint32_t nc = 0;
    uint64_t n = 0;              // rows covered by the lists
    std::vector<float> centroids;
    std::vector<uint32_t> assignments, offsets, members;
    std::vector<uint32_t> routing_dims;  // standalone IVF_FLAT inner-product routing (lb_ivf_flat_search), built on first use
    std::once_flag routing_once;         // ... exactly once, whichever thread searches first
    DevBuf d_ids, d_q, d_qw, d_rows, d_dists, d_counts, d_subset;
    DevBuf d_members, d_seg, d_out, d_allow;  // inverted lists in HBM; per-query (src, dst, len) copy descriptors; packed results
    HostBuf h_seg, h_out;                     // pinned staging of the descriptors and the results
};

namespace lb {
namespace {
struct FastRng {  // kmeans.rs:21-48
    uint64_t s;
    double next_f64() {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        return (double)(s >> 33) / (double)(1ull << 31);
    }
};
__global__ void ivf_copy_indexed_kernel(const float* src, int dim, const uint32_t* index, float* dst) {
    const uint32_t r = *index;
    for (int d = threadIdx.x; d < dim; d += blockDim.x) dst[d] = src[(size_t)r * dim + d];
}
void build_lists(lb_ivf* ivf) {  // inverted_lists_from_assignments (kmeans.rs:317-345): members in row order
    ivf->offsets.assign(ivf->nc + 1, 0);
    for (uint32_t c : ivf->assignments) ivf->offsets[c + 1]++;
    for (uint32_t c = 0; c < ivf->nc; ++c) ivf->offsets[c + 1] += ivf->offsets[c];
    ivf->members.resize(ivf->assignments.size());
    std::vector<uint32_t> cur(ivf->offsets.begin(), ivf->offsets.end() - 1);
    for (uint32_t i = 0; i < (uint32_t)ivf->assignments.size(); ++i) ivf->members[cur[ivf->assignments[i]]++] = i;
}
int ivf_finish(lb_ivf* ivf) {  // centroid index for the routing scan
    LB_TRY(lb_index_create(&ivf->cidx, ivf->idx->dim, LB_F32, ivf->idx->device));
    LB_TRY(lb_index_append_f32(ivf->cidx, ivf->centroids.data(), ivf->nc));
    build_lists(ivf);
    {   // the inverted lists live in HBM: a search copies the probed lists device-side
        DeviceGuard g(ivf->idx->device);
        LB_TRY(ivf->d_members.ensure(ivf->members.size() * 4));
        LB_CUDA_TRY(cudaMemcpy(ivf->d_members.p, ivf->members.data(), ivf->members.size() * 4, cudaMemcpyHostToDevice));
    }
    return LB_OK;
}
}  // namespace
}  // namespace lb

extern "C" {

void lb_ivf_destroy(lb_ivf* ivf) {
    if (!ivf) return;
    if (ivf->idx) {
        DeviceGuard g(ivf->idx->device);
        DevBuf* bufs[] = {&ivf->d_ids, &ivf->d_q, &ivf->d_qw, &ivf->d_rows, &ivf->d_dists, &ivf->d_counts, &ivf->d_subset,
                          &ivf->d_members, &ivf->d_seg, &ivf->d_out, &ivf->d_allow};
        for (DevBuf* b : bufs) b->release();
        ivf->h_seg.release();
        ivf->h_out.release();
    }
    if (ivf->cidx) lb_index_destroy(ivf->cidx);
    delete ivf;
}

int lb_ivf_create(lb_index* idx, int metric, const float* centroids, uint32_t n_centroids, const uint32_t* assignments, lb_ivf** out) {
    if (!idx || !out || !centroids || !assignments) return fail(LB_INVALID_ARGUMENT, "null argument");
    LB_TRY(check_metric(metric));
    if (idx->dtype != LB_F32) return fail(LB_INVALID_ARGUMENT, "IVF is built over f32 rows");
    if (n_centroids == 0 || idx->n == 0) return fail(LB_INVALID_ARGUMENT, "IVF needs rows and centroids");
    lb_ivf* ivf = new lb_ivf();
    ivf->idx = idx;
    ivf->metric = metric;
    ivf->routing = metric_binary(metric) ? LB_L2 : metric;  // ivf.rs:80-87
    ivf->nc = n_centroids;
    ivf->n = idx->n;
    ivf->centroids.assign(centroids, centroids + (size_t)n_centroids * idx->dim);
    ivf->assignments.assign(assignments, assignments + idx->n);
    for (uint32_t a : ivf->assignments)
        if (a >= n_centroids) {
            delete ivf;
            return fail(LB_INVALID_ARGUMENT, "assignment out of range");
        }
    int st = ivf_finish(ivf);
    if (st != LB_OK) {
        lb_ivf_destroy(ivf);
        return st;
    }
    *out = ivf;
    return LB_OK;
}

int lb_ivf_train(lb_index* idx, int metric, uint32_t n_clusters, uint32_t max_iter, lb_ivf** out) {
    if (!idx || !out) return fail(LB_INVALID_ARGUMENT, "null argument");
    LB_TRY(check_metric(metric));
    if (idx->dtype != LB_F32) return fail(LB_INVALID_ARGUMENT, "IVF is built over f32 rows");
    if (idx->n == 0 || n_clusters == 0) return fail(LB_INVALID_ARGUMENT, "IVF needs rows and at least one cluster");
    if (metric == LB_HAVERSINE && idx->dim != 2) return fail(LB_INVALID_ARGUMENT, "haversine requires dimension 2");
    const int routing = metric_binary(metric) ? LB_L2 : metric;
    const uint64_t n = idx->n;
    const int dim = (int)idx->dim;
    const uint32_t nc = (uint32_t)std::min<uint64_t>(n_clusters, n);
    lb_ivf* ivf = new lb_ivf();
    ivf->idx = idx;
    ivf->metric = metric;
    ivf->routing = routing;
    ivf->nc = nc;
    ivf->n = n;
    int st = LB_OK;
    {
        std::lock_guard<std::mutex> lock(idx->mu);
        DeviceGuard g(idx->device);
        cudaStream_t sm = idx->stream;
        DevBuf d_cent, d_sample, d_sidx, d_minr, d_best, d_assign, d_off, d_mem;
        auto body = [&]() -> int {
            // ---- kmeans_pp_init_metric: farthest-point seeding on a seeded sample (kmeans.rs:141-196)
            FastRng rng{42};
            uint64_t sample_n = std::min<uint64_t>(n, std::min<uint64_t>(std::max<uint64_t>((uint64_t)nc * 32, 2048), 10000));
            std::vector<uint32_t> sidx;
            if (sample_n >= n) {
                sidx.resize(n);
                for (uint64_t i = 0; i < n; ++i) sidx[i] = (uint32_t)i;
            } else {
                std::vector<uint32_t> all(n);
                for (uint64_t i = 0; i < n; ++i) all[i] = (uint32_t)i;
                for (uint64_t i = 0; i < sample_n; ++i) {
                    uint64_t j = i + std::min<uint64_t>((uint64_t)(rng.next_f64() * (double)(n - i)), n - i - 1);
                    std::swap(all[i], all[j]);
                }
                sidx.assign(all.begin(), all.begin() + sample_n);
            }
            sample_n = sidx.size();
            LB_TRY(d_cent.ensure((size_t)nc * dim * 4));
            LB_TRY(d_sample.ensure((size_t)sample_n * dim * 4));
            LB_TRY(d_sidx.ensure((size_t)sample_n * 4));
            LB_TRY(d_minr.ensure((size_t)sample_n * 4));
            LB_TRY(d_best.ensure(4));
            LB_CUDA_TRY(cudaMemcpyAsync(d_sidx.p, sidx.data(), (size_t)sample_n * 4, cudaMemcpyHostToDevice, sm));
            ivf_gather_rows_kernel<<<(unsigned)ceil_div(sample_n * dim, 256), 256, 0, sm>>>(idx->rows.as<float>(), dim, d_sidx.as<uint32_t>(),
                                                                                          (uint32_t)sample_n, d_sample.as<float>());
            std::vector<float> big(sample_n, 3.402823466e+38f);
            LB_CUDA_TRY(cudaMemcpyAsync(d_minr.p, big.data(), (size_t)sample_n * 4, cudaMemcpyHostToDevice, sm));
            const uint32_t first = (uint32_t)((uint64_t)(rng.next_f64() * (double)sample_n) % sample_n);
            LB_CUDA_TRY(cudaMemcpyAsync(d_cent.p, d_sample.as<float>() + (size_t)first * dim, (size_t)dim * 4, cudaMemcpyDeviceToDevice, sm));
            for (uint32_t c = 1; c < nc; ++c) {
                ivf_farthest_kernel<<<1, 1024, 0, sm>>>(d_sample.as<float>(), (uint32_t)sample_n, dim, d_cent.as<float>() + (size_t)(c - 1) * dim,
                                                       routing, d_minr.as<float>(), d_best.as<uint32_t>());
                ivf_copy_indexed_kernel<<<1, 256, 0, sm>>>(d_sample.as<float>(), dim, d_best.as<uint32_t>(), d_cent.as<float>() + (size_t)c * dim);
            }
            LB_CUDA_TRY(cudaGetLastError());
            // ---- Lloyd iterations (kmeans.rs:96-131)
            LB_TRY(d_assign.ensure((size_t)n * 4));
            LB_TRY(d_off.ensure((size_t)(nc + 1) * 4));
            LB_TRY(d_mem.ensure((size_t)n * 4));
            ivf->assignments.assign(n, 0xFFFFFFFFu);
            ivf->centroids.resize((size_t)nc * dim);
            std::vector<uint32_t> fresh(n);
            std::vector<float> old_c((size_t)nc * dim);
            for (uint32_t it = 0; it < max_iter; ++it) {
                ivf_assign_kernel<<<(unsigned)ceil_div(n, 128), 128, 0, sm>>>(idx->rows.as<float>(), n, dim, d_cent.as<float>(), (int)nc, routing,
                                                                               d_assign.as<uint32_t>());
                LB_CUDA_TRY(cudaGetLastError());
                LB_CUDA_TRY(cudaMemcpyAsync(fresh.data(), d_assign.p, (size_t)n * 4, cudaMemcpyDeviceToHost, sm));
                LB_CUDA_TRY(cudaMemcpyAsync(old_c.data(), d_cent.p, old_c.size() * 4, cudaMemcpyDeviceToHost, sm));
                LB_CUDA_TRY(cudaStreamSynchronize(sm));
                const bool changed = fresh != ivf->assignments;
                ivf->assignments = fresh;
                build_lists(ivf);
                LB_CUDA_TRY(cudaMemcpyAsync(d_off.p, ivf->offsets.data(), (size_t)(nc + 1) * 4, cudaMemcpyHostToDevice, sm));
                LB_CUDA_TRY(cudaMemcpyAsync(d_mem.p, ivf->members.data(), (size_t)n * 4, cudaMemcpyHostToDevice, sm));
                ivf_centroid_update_kernel<<<nc, 256, 0, sm>>>(idx->rows.as<float>(), dim, d_off.as<uint32_t>(), d_mem.as<uint32_t>(), d_cent.as<float>());
                LB_CUDA_TRY(cudaGetLastError());
                // empty clusters are re-seeded from the largest one, in centroid order (kmeans.rs:118-126)
                uint32_t max_c = 0, max_count = 0;
                bool any_empty = false;
                for (uint32_t c = 0; c < nc; ++c) {
                    const uint32_t cnt = ivf->offsets[c + 1] - ivf->offsets[c];
                    if (cnt >= max_count) {  // max_by_key keeps the last maximum
                        max_count = cnt;
                        max_c = c;
                    }
                    any_empty |= cnt == 0;
                }
                if (any_empty && max_count > 1) {
                    std::vector<float> new_c((size_t)nc * dim);
                    LB_CUDA_TRY(cudaMemcpyAsync(new_c.data(), d_cent.p, new_c.size() * 4, cudaMemcpyDeviceToHost, sm));
                    LB_CUDA_TRY(cudaStreamSynchronize(sm));
                    for (uint32_t c = 0; c < nc; ++c) {
                        if (ivf->offsets[c + 1] != ivf->offsets[c]) continue;
                        const float* src = (max_c < c ? new_c.data() : old_c.data()) + (size_t)max_c * dim;
                        for (int d = 0; d < dim; ++d) new_c[(size_t)c * dim + d] = src[d] * (1.0f + 1e-4f * (float)d);
                    }
                    LB_CUDA_TRY(cudaMemcpyAsync(d_cent.p, new_c.data(), new_c.size() * 4, cudaMemcpyHostToDevice, sm));
                    LB_CUDA_TRY(cudaStreamSynchronize(sm));
                }
                if (!changed) break;
            }
            ivf_assign_kernel<<<(unsigned)ceil_div(n, 128), 128, 0, sm>>>(idx->rows.as<float>(), n, dim, d_cent.as<float>(), (int)nc, routing,
                                                                           d_assign.as<uint32_t>());
            LB_CUDA_TRY(cudaGetLastError());
            LB_CUDA_TRY(cudaMemcpyAsync(ivf->assignments.data(), d_assign.p, (size_t)n * 4, cudaMemcpyDeviceToHost, sm));
            LB_CUDA_TRY(cudaMemcpyAsync(ivf->centroids.data(), d_cent.p, ivf->centroids.size() * 4, cudaMemcpyDeviceToHost, sm));
            LB_CUDA_TRY(cudaStreamSynchronize(sm));
            return LB_OK;
        };
        st = body();
        DevBuf* bufs[] = {&d_cent, &d_sample, &d_sidx, &d_minr, &d_best, &d_assign, &d_off, &d_mem};
        for (DevBuf* b : bufs) b->release();
    }
    if (st == LB_OK) st = ivf_finish(ivf);
    if (st != LB_OK) {
        lb_ivf_destroy(ivf);
        return st;
    }
    *out = ivf;
    return LB_OK;
}

int lb_ivf_info(const lb_ivf* ivf, uint32_t* n_centroids, uint64_t* n_rows) {
    if (!ivf) return fail(LB_INVALID_ARGUMENT, "ivf is null");
    if (n_centroids) *n_centroids = ivf->nc;
    if (n_rows) *n_rows = ivf->n;
    return LB_OK;
}
int lb_ivf_centroids(const lb_ivf* ivf, float* out) {
    if (!ivf || !out) return fail(LB_INVALID_ARGUMENT, "null argument");
    memcpy(out, ivf->centroids.data(), ivf->centroids.size() * 4);
    return LB_OK;
}
int lb_ivf_assignments(const lb_ivf* ivf, uint32_t* out) {
    if (!ivf || !out) return fail(LB_INVALID_ARGUMENT, "null argument");
    memcpy(out, ivf->assignments.data(), ivf->assignments.size() * 4);
    return LB_OK;
}

}  // extern "C"

namespace lb {
namespace {
// The `np` best centroids for each query under `metric`, best first (ties keep centroid order, as the reference's
// stable sort does): one exact scan over the centroid index.  `subset` (optional, host) restricts one query's
// ranking to the listed centroids (the inner-product routing shortlist of the standalone index).
int ivf_rank_centroids(lb_ivf* ivf, const float* queries, uint32_t nq, uint32_t np, int metric, const std::vector<uint32_t>* subset,
                       std::vector<uint32_t>& probe) {
    lb_index* c = ivf->cidx;
    const int dim = (int)c->dim;
    std::lock_guard<std::mutex> lock(c->mu);
    DeviceGuard g(c->device);
    LB_TRY(c->w_queries.ensure((size_t)nq * dim * 4));
    LB_TRY(c->w_out_rows.ensure((size_t)nq * np * 4));
    LB_TRY(c->w_out_dists.ensure((size_t)nq * np * 4));
    LB_TRY(c->w_out_counts.ensure((size_t)nq * 4));
    LB_CUDA_TRY(cudaMemcpyAsync(c->w_queries.p, queries, (size_t)nq * dim * 4, cudaMemcpyHostToDevice, c->stream));
    ScanRequest r;
    r.corpus = c->rows.as<float>();
    r.n_rows = ivf->nc;
    r.dim = dim;
    r.queries = c->w_queries.as<float>();
    r.nq = (int)nq;
    r.k = (int)np;
    r.metric = metric;
    r.ip_single = 1;
    r.out_rows = c->w_out_rows.as<uint32_t>();
    r.out_dists = c->w_out_dists.as<float>();
    r.out_counts = c->w_out_counts.as<uint32_t>();
    if (subset) {
        LB_TRY(ivf->d_subset.ensure(subset->size() * 4));  // guarded by the centroid index's mutex
        LB_CUDA_TRY(cudaMemcpyAsync(ivf->d_subset.p, subset->data(), subset->size() * 4, cudaMemcpyHostToDevice, c->stream));
        r.row_ids = ivf->d_subset.as<uint32_t>();
        r.n_rows = subset->size();
    }
    LB_TRY(run_scan(c, r, nullptr, nullptr));
    probe.assign((size_t)nq * np, ROW_NONE);
    LB_CUDA_TRY(cudaMemcpyAsync(probe.data(), c->w_out_rows.p, probe.size() * 4, cudaMemcpyDeviceToHost, c->stream));
    LB_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return LB_OK;
}

// Per query: the probed lists (whole lists, in probe order) become one row list in HBM — copied device-side from the
// resident inverted lists, the host only sends (source, destination, length) per list — the subset filter rides along
// as the scan's allow-bitset, every candidate is scored with compute_distance_f32 / the packed kernels, the k best
// are kept.  `corpus_fallback`: a probe without a single allowed row falls back to the filtered corpus
// (IVFIndex::search) instead of returning nothing (IvfFlatMmap::search).
__global__ void ivf_expand_lists_kernel(const uint32_t* __restrict__ members, const uint32_t* __restrict__ seg /*[n][3]*/,
                                        uint32_t* __restrict__ ids) {
    // grid (lists, slices): slice y of a list is its elements [1024 y, 1024 (y + 1))
    const uint32_t src = seg[3 * blockIdx.x], dst = seg[3 * blockIdx.x + 1], len = seg[3 * blockIdx.x + 2];
    const uint32_t lo = blockIdx.y * 1024u, hi = min(len, lo + 1024u);
    for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) ids[dst + i] = members[src + i];
}

int ivf_scan_probes(lb_ivf* ivf, const float* queries, uint32_t nq, uint32_t k, int metric, const std::vector<uint32_t>& probe, uint32_t np,
                    const uint64_t* allow_bits, bool corpus_fallback, uint32_t* out_rows, float* out_dists, uint32_t* out_counts) {
    lb_index* idx = ivf->idx;
    const int dim = (int)idx->dim;
    std::lock_guard<std::mutex> lock(idx->mu);
    DeviceGuard g(idx->device);
    cudaStream_t sm = idx->stream;
    const bool binary = metric_binary(metric);
    const int nw = (dim + 63) / 64;
    if (binary) LB_TRY(ensure_packed(idx));
    // queries (and their packed form), the filter: once per call
    LB_TRY(ivf->d_q.ensure((size_t)nq * dim * 4));
    LB_CUDA_TRY(cudaMemcpyAsync(ivf->d_q.p, queries, (size_t)nq * dim * 4, cudaMemcpyHostToDevice, sm));
    if (binary) {
        LB_TRY(ivf->d_qw.ensure((size_t)nq * nw * 8));
        pack_binary_kernel<<<(nq + 7) / 8, 256, 0, sm>>>(ivf->d_q.as<float>(), (uint64_t)nq, dim, nw, 0.5f, ivf->d_qw.as<uint64_t>());
        LB_CUDA_TRY(cudaGetLastError());
    }
    const uint64_t* d_allow = nullptr;
    if (allow_bits) {
        const size_t words = (ivf->n + 63) / 64;
        LB_TRY(ivf->d_allow.ensure(words * 8));
        LB_CUDA_TRY(cudaMemcpyAsync(ivf->d_allow.p, allow_bits, words * 8, cudaMemcpyHostToDevice, sm));
        d_allow = ivf->d_allow.as<uint64_t>();
    }
    LB_TRY(ivf->h_seg.ensure((size_t)np * 12));
    LB_TRY(ivf->d_seg.ensure((size_t)np * 12));
    LB_TRY(ivf->d_out.ensure((size_t)k * 8 + 4));
    LB_TRY(ivf->h_out.ensure((size_t)k * 8 + 4));
    uint32_t* seg = reinterpret_cast<uint32_t*>(ivf->h_seg.p);
    uint32_t* d_rows = ivf->d_out.as<uint32_t>();
    float* d_dists = reinterpret_cast<float*>(ivf->d_out.as<char>() + (size_t)k * 4);
    uint32_t* d_count = reinterpret_cast<uint32_t*>(ivf->d_out.as<char>() + (size_t)k * 8);
    auto scan = [&](uint32_t q, const uint32_t* row_ids, uint64_t n_rows) -> int {  // -> pinned h_out
        ScanRequest r;
        r.n_rows = n_rows;
        r.row_ids = row_ids;
        r.nq = 1;
        r.k = (int)std::min<uint64_t>(k, n_rows);
        r.metric = metric;
        r.allow_bits = d_allow;
        r.out_rows = d_rows;
        r.out_dists = d_dists;
        r.out_counts = d_count;
        if (binary) {
            r.words = idx->packed.as<uint64_t>();
            r.n_words = nw;
            r.qwords = ivf->d_qw.as<uint64_t>() + (size_t)q * nw;
        } else {
            r.corpus = idx->rows.as<float>();
            r.dim = dim;
            r.queries = ivf->d_q.as<float>() + (size_t)q * dim;
            r.ip_single = 1;  // compute_distance_f32 -> the single-row IP kernel
        }
        LB_TRY(run_scan(idx, r, nullptr, nullptr));
        LB_CUDA_TRY(cudaMemcpyAsync(ivf->h_out.p, ivf->d_out.p, (size_t)k * 8 + 4, cudaMemcpyDeviceToHost, sm));
        LB_CUDA_TRY(cudaStreamSynchronize(sm));
        return LB_OK;
    };
    const bool trace = tc_env_int("LYNSE_B200_IVF_TRACE", 0) != 0;
    for (uint32_t q = 0; q < nq; ++q) {
        const auto tr0 = std::chrono::steady_clock::now();
        uint32_t total = 0, n_seg = 0, max_len = 0;
        for (uint32_t p = 0; p < np; ++p) {
            const uint32_t c = probe[(size_t)q * np + p];
            if (c >= ivf->nc) continue;
            const uint32_t len = ivf->offsets[c + 1] - ivf->offsets[c];
            if (len == 0) continue;
            seg[3 * n_seg] = ivf->offsets[c];
            seg[3 * n_seg + 1] = total;
            seg[3 * n_seg + 2] = len;
            ++n_seg;
            total += len;
            max_len = std::max(max_len, len);
        }
        uint32_t found = 0;
        if (total > 0) {
            LB_TRY(ivf->d_ids.ensure((size_t)total * 4));
            LB_CUDA_TRY(cudaMemcpyAsync(ivf->d_seg.p, seg, (size_t)n_seg * 12, cudaMemcpyHostToDevice, sm));
            ivf_expand_lists_kernel<<<dim3(n_seg, (max_len + 1023) / 1024), 256, 0, sm>>>(ivf->d_members.as<uint32_t>(), ivf->d_seg.as<uint32_t>(), ivf->d_ids.as<uint32_t>());
            LB_CUDA_TRY(cudaGetLastError());
            const auto tr1 = std::chrono::steady_clock::now();
            LB_TRY(scan(q, ivf->d_ids.as<uint32_t>(), total));
            found = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(ivf->h_out.p) + (size_t)k * 8);
            if (trace && q < 4) {
                const auto tr2 = std::chrono::steady_clock::now();
                fprintf(stderr, "[lynse_b200] ivf query %u: %u rows in %u lists; lists+expand enqueue %.1f us, scan+copy+sync %.1f us\n", q, total, n_seg,
                        std::chrono::duration<double, std::micro>(tr1 - tr0).count(), std::chrono::duration<double, std::micro>(tr2 - tr1).count());
            }
        }
        if (found == 0 && corpus_fallback && ivf->n > 0) {  // nothing allowed among the probed rows: the filtered corpus
            LB_TRY(scan(q, nullptr, ivf->n));
            found = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(ivf->h_out.p) + (size_t)k * 8);
        }
        if (found == 0) continue;
        const char* h = reinterpret_cast<const char*>(ivf->h_out.p);
        memcpy(out_rows + (size_t)q * k, h, (size_t)found * 4);
        memcpy(out_dists + (size_t)q * k, h + (size_t)k * 4, (size_t)found * 4);
        out_counts[q] = found;
    }
    return LB_OK;
}

int ivf_search_prologue(lb_ivf* ivf, const float* queries, uint32_t nq, uint32_t k, uint32_t* out_rows, float* out_dists, uint32_t* out_counts) {
    if (!ivf || !out_rows || !out_dists || !out_counts) return fail(LB_INVALID_ARGUMENT, "null argument");
    if (nq && !queries) return fail(LB_INVALID_ARGUMENT, "queries is null");
    if (ivf->idx->n != ivf->n) return fail(LB_INVALID_ARGUMENT, "the index changed since the IVF lists were built; rebuild the index");
    if (k > (uint32_t)MAX_K) return fail(LB_UNSUPPORTED, "k above 2048 is not supported");
    for (uint32_t q = 0; q < nq; ++q) out_counts[q] = 0;
    for (size_t i = 0; i < (size_t)nq * k; ++i) {
        out_rows[i] = ROW_NONE;
        out_dists[i] = NAN;
    }
    return LB_OK;
}

// select_routing_dims (src/storage/ivf_flat_mmap.rs:312-345): the 16 centroid dimensions of highest variance,
// ascending; empty unless dim >= 64 and there are >= 64 centroids.
std::vector<uint32_t> ivf_routing_dims(const std::vector<float>& centroids, size_t dim, size_t nc) {
    if (dim < 64 || nc < 64) return {};
    std::vector<float> sums(dim, 0.0f), sq(dim, 0.0f);
    for (size_t c = 0; c < nc; ++c)
        for (size_t d = 0; d < dim; ++d) {
            const float v = centroids[c * dim + d];
            sums[d] += v;
            sq[d] += v * v;  // unfused, as rustc emits it: host code is built with -ffp-contract=off
        }
    const float inv_k = 1.0f / (float)nc;
    std::vector<std::pair<float, uint32_t>> dims(dim);
    for (size_t d = 0; d < dim; ++d) {
        const float mean = sums[d] * inv_k;
        dims[d] = {sq[d] * inv_k - mean * mean, (uint32_t)d};
    }
    std::stable_sort(dims.begin(), dims.end(), [](const auto& a, const auto& b) { return a.first > b.first; });
    std::vector<uint32_t> sel(16);
    for (size_t i = 0; i < 16; ++i) sel[i] = dims[i].second;
    std::sort(sel.begin(), sel.end());
    return sel;
}
}  // namespace
}  // namespace lb

extern "C" {

// IVFIndex::search (src/index/ivf.rs:181-348): rank centroids with the routing metric, gather the nprobe nearest
// lists, apply the subset filter (an empty probe falls back to the filtered corpus, never to an unfiltered scan),
// score every candidate, keep the k best.
int lb_ivf_search(lb_ivf* ivf, const float* queries, uint32_t nq, uint32_t k, uint32_t nprobe, const uint64_t* allow_bits,
                  uint64_t allow_words, uint32_t* out_rows, float* out_dists, uint32_t* out_counts) {
    LB_TRY(ivf_search_prologue(ivf, queries, nq, k, out_rows, out_dists, out_counts));
    if (allow_bits && allow_words < (ivf->n + 63) / 64) return fail(LB_INVALID_ARGUMENT, "row filter is shorter than the index");
    if (nq == 0 || k == 0) return LB_OK;
    const uint32_t np = std::min<uint32_t>(std::max<uint32_t>(nprobe, 1), ivf->nc);
    std::vector<uint32_t> probe;
    LB_TRY(ivf_rank_centroids(ivf, queries, nq, np, ivf->routing, nullptr, probe));
    return ivf_scan_probes(ivf, queries, nq, k, ivf->metric, probe, np, allow_bits, true, out_rows, out_dists, out_counts);
}

// IvfFlatMmap::search (src/storage/ivf_flat_mmap.rs:225-300) with find_nearest_centroids (:383-446): the probed
// partitions are the nprobe nearest centroids under the SEARCH metric — for inner product on dim >= 64 with >= 64
// partitions, the nearest of a routing-dimension shortlist — and every row of those partitions is scored with
// compute_distance_f32.  Rows are reported by their position in the build data (the reference's original_ids).
int lb_ivf_flat_search(lb_ivf* ivf, const float* queries, uint32_t nq, uint32_t k, uint32_t nprobe, int metric, uint32_t* out_rows,
                       float* out_dists, uint32_t* out_counts) {
    LB_TRY(ivf_search_prologue(ivf, queries, nq, k, out_rows, out_dists, out_counts));
    LB_TRY(check_metric(metric));
    const size_t dim = ivf->idx->dim;
    if (metric == LB_HAVERSINE && dim != 2) return fail(LB_INVALID_ARGUMENT, "haversine requires dimension 2");
    if (nq == 0 || k == 0) return LB_OK;
    const uint32_t nc = ivf->nc;
    const uint32_t np = std::min<uint32_t>(std::max<uint32_t>(nprobe, 1), nc);
    std::vector<uint32_t> probe;
    if (np >= nc) {
        probe.resize((size_t)nq * np);
        for (uint32_t q = 0; q < nq; ++q)
            for (uint32_t c = 0; c < nc; ++c) probe[(size_t)q * np + c] = c;
    } else if (metric == LB_IP && dim >= 64 && nc >= 64) {
        std::call_once(ivf->routing_once, [&] { ivf->routing_dims = ivf_routing_dims(ivf->centroids, dim, nc); });
        const size_t shortlist = std::min<size_t>(std::max<size_t>(std::min<size_t>((size_t)np * 3, 96), 24), nc);
        probe.assign((size_t)nq * np, ROW_NONE);
        std::vector<std::pair<float, uint32_t>> best(shortlist);
        std::vector<uint32_t> subset, ranked;
        for (uint32_t q = 0; q < nq; ++q) {
            const float* qv = queries + (size_t)q * dim;
            size_t len = 0;
            for (uint32_t c = 0; c < nc; ++c) {
                const float* cv = ivf->centroids.data() + (size_t)c * dim;
                float score = 0.0f;
                for (uint32_t d : ivf->routing_dims) score += qv[d] * cv[d];  // coarse_ip_score: multiply then add, unfused
                if (len < shortlist) {
                    best[len++] = {score, c};
                    continue;
                }
                size_t worst = 0;
                float worst_score = best[0].first;
                for (size_t i = 1; i < shortlist; ++i)
                    if (best[i].first < worst_score) {
                        worst_score = best[i].first;
                        worst = i;
                    }
                if (score > worst_score) best[worst] = {score, c};
            }
            subset.resize(len);
            for (size_t i = 0; i < len; ++i) subset[i] = best[i].second;
            std::sort(subset.begin(), subset.end());  // ranking ties then fall to the lower centroid
            LB_TRY(ivf_rank_centroids(ivf, qv, 1, np, LB_IP, &subset, ranked));
            for (uint32_t p = 0; p < np; ++p) probe[(size_t)q * np + p] = ranked[p];
        }
    } else {
        LB_TRY(ivf_rank_centroids(ivf, queries, nq, np, metric, nullptr, probe));
    }
    return ivf_scan_probes(ivf, queries, nq, k, metric, probe, np, nullptr, false, out_rows, out_dists, out_counts);
}

}  // extern "C"

extern "C" {

// ---- NCCL (resolved at run time so the library loads on boxes without it) ----------------------------------------------------
namespace {
struct NcclId {
    char internal[128];
};
typedef void* NcclComm;
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (api.handle) {
            api.GetUniqueId = reinterpret_cast<int (*)(NcclId*)>(dlsym(api.handle, "ncclGetUniqueId"));
            api.CommInitRank = reinterpret_cast<int (*)(NcclComm*, int, NcclId, int)>(dlsym(api.handle, "ncclCommInitRank"));
            api.AllGather = reinterpret_cast<int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t)>(
                dlsym(api.handle, "ncclAllGather"));
            api.CommDestroy = reinterpret_cast<int (*)(NcclComm)>(dlsym(api.handle, "ncclCommDestroy"));
            api.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(api.handle, "ncclGetErrorString"));
        }
    }
    if (!api.handle || !api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.CommDestroy) return nullptr;
    return &api;
}
int nccl_fail(NcclApi* api, const char* what, int rc) {
    return fail(LB_NCCL, std::string(what) + ": " + (api && api->GetErrorString ? api->GetErrorString(rc) : "nccl error"));
}
}  // namespace

struct lb_comm {
    NcclComm comm = nullptr;
    int device = 0, world = 1, rank = 0;
    cudaStream_t stream = nullptr;
};

int lb_nccl_unique_id(uint8_t* id128) {
    NcclApi* api = nccl_api();
    if (!api) return fail(LB_NCCL, "libnccl.so.2 could not be loaded");
    NcclId id;
    int rc = api->GetUniqueId(&id);
    if (rc != 0) return nccl_fail(api, "ncclGetUniqueId", rc);
    memcpy(id128, id.internal, 128);
    return LB_OK;
}

int lb_comm_create(lb_comm** out, int device, int world_size, int rank, const uint8_t* id128) {
    NcclApi* api = nccl_api();
    if (!api) return fail(LB_NCCL, "libnccl.so.2 could not be loaded");
    if (!out || !id128 || world_size < 1 || rank < 0 || rank >= world_size) return fail(LB_INVALID_ARGUMENT, "bad communicator arguments");
    LB_CUDA_TRY(cudaSetDevice(device));
    lb_comm* c = new lb_comm();
    c->device = device;
    c->world = world_size;
    c->rank = rank;
    NcclId id;
    memcpy(id.internal, id128, 128);
    int rc = api->CommInitRank(&c->comm, world_size, id, rank);
    if (rc != 0) {
        delete c;
        return nccl_fail(api, "ncclCommInitRank", rc);
    }
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        api->CommDestroy(c->comm);
        delete c;
        return fail(LB_CUDA, cudaGetErrorString(e));
    }
    *out = c;
    return LB_OK;
}

void lb_comm_destroy(lb_comm* c) {
    if (!c) return;
    NcclApi* api = nccl_api();
    cudaSetDevice(c->device);
    if (c->stream) {
        cudaStreamSynchronize(c->stream);
        cudaStreamDestroy(c->stream);
    }
    if (api && c->comm) api->CommDestroy(c->comm);
    delete c;
}

static int comm_allgather_on(lb_comm* c, const void* d_send, void* d_recv, uint64_t bytes, cudaStream_t stream) {
    NcclApi* api = nccl_api();
    if (!api || !c) return fail(LB_NCCL, "communicator is not initialised");
    int rc = api->AllGather(d_send, d_recv, (size_t)bytes, /*ncclUint8*/ 1, c->comm, stream);
    if (rc != 0) return nccl_fail(api, "ncclAllGather", rc);
    return LB_OK;
}

int lb_comm_allgather(lb_comm* c, const void* d_send, void* d_recv, uint64_t bytes) {
    if (!c) return fail(LB_NCCL, "communicator is not initialised");
    DeviceGuard g(c->device);
    LB_TRY(comm_allgather_on(c, d_send, d_recv, bytes, c->stream));
    LB_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return LB_OK;
}

int lb_comm_barrier(lb_comm* c) {
    if (!c) return fail(LB_NCCL, "communicator is not initialised");
    DeviceGuard g(c->device);
    static thread_local void* scratch = nullptr;
    if (!scratch) LB_CUDA_TRY(cudaMalloc(&scratch, 8 * 1024));
    if (c->world > 1024) return fail(LB_UNSUPPORTED, "world too large");
    return lb_comm_allgather(c, scratch, reinterpret_cast<char*>(scratch) + 4096, 4);
}

// ---- sharded search: local search -> ncclAllGather -> GPU merge by (score, global row) ------------------------------------------------
namespace lb {
// block layout per rank: rows u32 [nq*k] | dists f32 [nq*k] | counts u32 [nq] | pad to 8 | tail: base u64, uncertified u32, status u32
constexpr size_t SHARD_TAIL_BYTES = 16;
static size_t shard_block_bytes(uint32_t nq, uint32_t k) {
    size_t b = (size_t)nq * k * 8 + (size_t)nq * 4;
    b = (b + 7) & ~(size_t)7;
    return b + SHARD_TAIL_BYTES;
}
// tail of this rank's block: its row base, how many queries of its tensor-core search still wait for the exact-scan
// fallback (flags[1], read on the device: no host round trip before the collective) and its status (non-zero = this
// rank's search failed; every rank sees it after the gather and fails with it)
static __global__ void shard_pack_tail_kernel(unsigned char* block, size_t tail_off, uint64_t row_base, const uint32_t* tc_flags, uint32_t status) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        *reinterpret_cast<uint64_t*>(block + tail_off) = row_base;
        uint32_t* w = reinterpret_cast<uint32_t*>(block + tail_off + 8);
        w[0] = tc_flags != nullptr ? tc_flags[1] : 0u;
        w[1] = status != 0u ? status : (tc_flags != nullptr && tc_flags[0] != 0u ? (uint32_t)LB_INTERNAL : 0u);
    }
}
// one CTA per query; G*k <= 4096.  Blocks arrive in rank order and ranks hold ascending row ranges (asserted by the
// caller through the gathered bases), so (score, shard, position) is (score, global row).
static __global__ void __launch_bounds__(256) merge_shards_kernel(const unsigned char* gathered, int G, size_t block_bytes, int nq, int k,
                                                                  int M, int asc, uint64_t* out_rows, float* out_dists,
                                                                  uint32_t* out_counts) {
    extern __shared__ __align__(16) unsigned char smem_shard[];
    uint64_t* s = reinterpret_cast<uint64_t*>(smem_shard);
    const int q = blockIdx.x;
    const size_t dists_off = (size_t)nq * k * 4, counts_off = (size_t)nq * k * 8;
    const size_t base_off = block_bytes - SHARD_TAIL_BYTES;
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
        uint64_t key = KEY_NONE;
        if (i < G * k) {
            int g = i / k, j = i - g * k;
            const unsigned char* blk = gathered + (size_t)g * block_bytes;
            uint32_t cnt = reinterpret_cast<const uint32_t*>(blk + counts_off)[q];
            if ((uint32_t)j < cnt) {
                float d = reinterpret_cast<const float*>(blk + dists_off)[(size_t)q * k + j];
                // each shard's list is already ordered by (score, local row) and shards are ascending row ranges,
                // so (score, shard, position) is (score, global row)
                key = asc ? make_key<true>(d, (uint32_t)i) : make_key<false>(d, (uint32_t)i);
            }
        }
        s[i] = key;
    }
    bitonic_sort_u64(s, M);
    __shared__ uint32_t n_valid;
    if (threadIdx.x == 0) n_valid = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        uint64_t key = s[i];
        uint64_t row = ~0ull;
        float d = __int_as_float(0x7fc00000);
        if (key != KEY_NONE) {
            int slot = (int)key_row(key);
            int g = slot / k, j = slot - g * k;
            const unsigned char* blk = gathered + (size_t)g * block_bytes;
            uint64_t base = *reinterpret_cast<const uint64_t*>(blk + base_off);
            row = base + reinterpret_cast<const uint32_t*>(blk)[(size_t)q * k + j];
            d = reinterpret_cast<const float*>(blk + dists_off)[(size_t)q * k + j];
            atomicAdd(&n_valid, 1u);
        }
        out_rows[(size_t)q * k + i] = row;
        out_dists[(size_t)q * k + i] = d;
    }
    __syncthreads();
    if (threadIdx.x == 0) out_counts[q] = n_valid;
}

// Everything is enqueued on the index stream without a host round trip: search (certification flags left on the
// device), tail, all-gather, merge; then ONE synchronisation reads the gathered tails.  Only when some rank reports
// uncertified queries (every rank sees the same tails, so every rank takes the same branch) do the ranks run their
// exact-scan fallbacks and repeat gather + merge.  A rank whose local search failed still takes part in the collective,
// with its status in the tail, so all ranks fail together instead of hanging in ncclAllGather.
static int sharded_search_device_impl(lb_comm* comm, lb_index* idx, int metric, const void* d_queries, uint32_t nq, uint32_t k,
                                      uint64_t row_base, uint64_t* d_out_rows, float* d_out_dists, uint32_t* d_out_counts,
                                      const uint64_t* d_allow = nullptr) {
    const int G = comm ? comm->world : 1;
    if ((uint64_t)G * k > 4096) return fail(LB_UNSUPPORTED, "world_size * k must not exceed 4096");
    const auto tr_begin = std::chrono::steady_clock::now();
    const size_t bb = shard_block_bytes(nq, k);
    LB_TRY(idx->w_send.ensure(bb));
    LB_TRY(idx->w_recv.ensure(bb * G));
    LB_TRY(idx->h_tails.ensure((size_t)G * SHARD_TAIL_BYTES + 16));   // + this rank's flag head (read with the same wait)
    unsigned char* send = idx->w_send.as<unsigned char>();
    uint32_t* s_rows = reinterpret_cast<uint32_t*>(send);
    float* s_dists = reinterpret_cast<float*>(send + (size_t)nq * k * 4);
    uint32_t* s_counts = reinterpret_cast<uint32_t*>(send + (size_t)nq * k * 8);
    int local_status = k > idx->n ? fail(LB_INVALID_ARGUMENT, "k exceeds the rows of this shard") : LB_OK;
    if (local_status == LB_OK)
        local_status = search_device_impl(idx, metric, d_queries, (int)nq, (int)k, d_allow, s_rows, s_dists, s_counts, false, true);
    const std::string local_error = local_status != LB_OK ? std::string(lb_last_error()) : std::string();
    const lb_search_stats local = idx->stats;
    const int M = std::max(2, next_pow2(G * (int)k));
    const unsigned char* gathered = G > 1 ? idx->w_recv.as<unsigned char>() : send;
    auto gather_and_merge = [&]() -> int {
        shard_pack_tail_kernel<<<1, 32, 0, idx->stream>>>(send, bb - SHARD_TAIL_BYTES, row_base,
                                                         idx->pending_tc.active ? idx->w_flags.as<uint32_t>() : nullptr, (uint32_t)local_status);
        LB_CUDA_TRY(cudaGetLastError());
        if (G > 1) LB_TRY(comm_allgather_on(comm, send, idx->w_recv.p, bb, idx->stream));
        merge_shards_kernel<<<nq, 256, (size_t)M * 8, idx->stream>>>(gathered, G, bb, (int)nq, (int)k, M, metric_ascending(metric) ? 1 : 0,
                                                                   d_out_rows, d_out_dists, d_out_counts);
        LB_CUDA_TRY(cudaGetLastError());
        LB_CUDA_TRY(cudaMemcpy2DAsync(idx->h_tails.p, SHARD_TAIL_BYTES, gathered + bb - SHARD_TAIL_BYTES, bb, SHARD_TAIL_BYTES, (size_t)G,
                                      cudaMemcpyDeviceToHost, idx->stream));
        return LB_OK;
    };
    LB_TRY(gather_and_merge());
    // this rank's certification flags ride along with the tails: one wait, one wake-up per step
    uint32_t* flag_head = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(idx->h_tails.p) + (size_t)G * SHARD_TAIL_BYTES);
    const bool flags_pending = idx->pending_tc.active;
    if (flags_pending)
        LB_CUDA_TRY(cudaMemcpyAsync(flag_head, idx->w_flags.p, 16, cudaMemcpyDeviceToHost, idx->stream));
    const auto tr_enq = std::chrono::steady_clock::now();
    LB_CUDA_TRY(cudaStreamSynchronize(idx->stream));
    const auto tr_sync = std::chrono::steady_clock::now();
    const unsigned char* tails = reinterpret_cast<const unsigned char*>(idx->h_tails.p);
    uint32_t any_uncertified = 0, any_status = 0;
    uint64_t prev_base = 0;
    bool ascending = true;
    for (int g = 0; g < G; ++g) {
        uint64_t base;
        uint32_t w[2];
        memcpy(&base, tails + (size_t)g * SHARD_TAIL_BYTES, 8);
        memcpy(w, tails + (size_t)g * SHARD_TAIL_BYTES + 8, 8);
        any_uncertified += w[0];
        if (w[1] != 0 && any_status == 0) any_status = w[1];
        if (g > 0 && base < prev_base) ascending = false;
        prev_base = base;
    }
    if (any_status != 0) {
        idx->pending_tc.active = false;
        if (local_status != LB_OK) return fail(local_status, local_error);
        return fail((int)any_status, "a peer rank's shard search failed");
    }
    if (!ascending) {
        idx->pending_tc.active = false;
        return fail(LB_INVALID_ARGUMENT, "row bases must ascend with the rank (the merge breaks ties by shard order)");
    }
    int extra_kernels = 0;
    lb_search_stats fin = local;
    if (idx->pending_tc.active) {
        // reads this rank's flags (stream already idle) and runs its fallback when it has one
        idx->stats = local;
        LB_TRY(tc_finish(idx, nullptr, flags_pending ? flag_head : nullptr));
        fin = idx->stats;
    }
    if (any_uncertified != 0) {
        LB_TRY(gather_and_merge());
        LB_CUDA_TRY(cudaStreamSynchronize(idx->stream));
        extra_kernels = 2 + (G > 1 ? 1 : 0);
    }
    idx->stats = fin;
    idx->stats.kernels_launched += 2 + (G > 1 ? 1 : 0) + extra_kernels;
    if (tc_env_int("LYNSE_B200_TC_TRACE", 0) != 0) {
        const auto tr_end = std::chrono::steady_clock::now();
        auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
        fprintf(stderr, "[lynse_b200] sharded step (host clock): enqueue %.1f us, wait for the device %.1f us, flags + fallback %.1f us\n",
                us(tr_begin, tr_enq), us(tr_enq, tr_sync), us(tr_sync, tr_end));
    }
    return LB_OK;
}
}  // namespace lb

static int sharded_check(lb_comm* comm, lb_index* idx, int metric, uint32_t nq, uint32_t k) {
    if (!idx) return fail(LB_INVALID_ARGUMENT, "index is null");
    LB_TRY(check_metric(metric));
    if (comm && comm->device != idx->device) return fail(LB_INVALID_ARGUMENT, "communicator and index live on different devices");
    // rank-uniform preconditions only: anything that depends on this rank's shard is reported through the collective
    if (k == 0 || k > (uint32_t)MAX_K) return fail(LB_INVALID_ARGUMENT, "k must be in [1, 2048] for a sharded search");
    if (nq == 0 || nq > (uint32_t)QUERY_BATCH) return fail(LB_INVALID_ARGUMENT, "nq must be in [1, 4096] for a sharded search");
    return LB_OK;
}

int lb_sharded_search_device(lb_comm* comm, lb_index* idx, int metric, const void* d_queries, uint32_t nq, uint32_t k,
                             uint64_t row_base, uint64_t* d_out_rows, float* d_out_dists, uint32_t* d_out_counts) {
    LB_TRY(sharded_check(comm, idx, metric, nq, k));
    if (!d_queries || !d_out_rows || !d_out_dists || !d_out_counts) return fail(LB_INVALID_ARGUMENT, "null device pointer");
    std::lock_guard<std::mutex> lock(idx->mu);
    DeviceGuard g(idx->device);
    return sharded_search_device_impl(comm, idx, metric, d_queries, nq, k, row_base, d_out_rows, d_out_dists, d_out_counts);
}

// A host pointer the device can address directly (pinned, e.g. from lb_host_malloc): the merge kernel writes the
// result there itself, so no copy follows it.
static void* device_alias_of_host(const void* p) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}

static int sharded_search_host(lb_comm* comm, lb_index* idx, int metric, const void* queries, size_t query_row_bytes, uint32_t nq,
                               uint32_t k, uint64_t row_base, const uint64_t* allow_bits, uint64_t allow_words, uint64_t* out_rows,
                               float* out_dists, uint32_t* out_counts) {
    LB_TRY(sharded_check(comm, idx, metric, nq, k));
    if (!queries || !out_rows || !out_dists || !out_counts) return fail(LB_INVALID_ARGUMENT, "null argument");
    std::lock_guard<std::mutex> lock(idx->mu);
    DeviceGuard g(idx->device);
    LB_TRY(idx->w_queries.ensure((size_t)nq * query_row_bytes));
    const uint64_t* d_allow = nullptr;
    if (allow_bits != nullptr) {
        const uint64_t need_words = (idx->n + 63) / 64;
        if (allow_words < need_words) return fail(LB_INVALID_ARGUMENT, "allow_bits is shorter than the shard");
        LB_TRY(idx->w_allow.ensure(need_words * 8));
        LB_CUDA_TRY(cudaMemcpyAsync(idx->w_allow.p, allow_bits, need_words * 8, cudaMemcpyHostToDevice, idx->stream));
        uint64_t cnt = 0;
        for (uint64_t w = 0; w < need_words; ++w) {
            uint64_t x = allow_bits[w];
            if (w + 1 == need_words && (idx->n & 63)) x &= (1ull << (idx->n & 63)) - 1;
            cnt += (uint64_t)__builtin_popcountll(x);
        }
        idx->allow_count = cnt;
        d_allow = idx->w_allow.as<uint64_t>();
    }
    uint64_t* a_rows = reinterpret_cast<uint64_t*>(device_alias_of_host(out_rows));
    float* a_dists = reinterpret_cast<float*>(device_alias_of_host(out_dists));
    uint32_t* a_counts = reinterpret_cast<uint32_t*>(device_alias_of_host(out_counts));
    const bool direct = a_rows && a_dists && a_counts;
    if (!direct) {
        LB_TRY(idx->w_g_rows.ensure((size_t)nq * k * 8));
        LB_TRY(idx->w_g_dists.ensure((size_t)nq * k * 4));
        LB_TRY(idx->w_g_counts.ensure((size_t)nq * 4));
        a_rows = idx->w_g_rows.as<uint64_t>();
        a_dists = idx->w_g_dists.as<float>();
        a_counts = idx->w_g_counts.as<uint32_t>();
    }
    LB_CUDA_TRY(cudaMemcpyAsync(idx->w_queries.p, queries, (size_t)nq * query_row_bytes, cudaMemcpyHostToDevice, idx->stream));
    LB_TRY(sharded_search_device_impl(comm, idx, metric, idx->w_queries.p, nq, k, row_base, a_rows, a_dists, a_counts, d_allow));
    if (!direct) {
        LB_CUDA_TRY(cudaMemcpyAsync(out_rows, idx->w_g_rows.p, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, idx->stream));
        LB_CUDA_TRY(cudaMemcpyAsync(out_dists, idx->w_g_dists.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, idx->stream));
        LB_CUDA_TRY(cudaMemcpyAsync(out_counts, idx->w_g_counts.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, idx->stream));
        LB_CUDA_TRY(cudaStreamSynchronize(idx->stream));
    }
    return LB_OK;
}

int lb_sharded_search(lb_comm* comm, lb_index* idx, int metric, const float* queries, uint32_t nq, uint32_t k, uint64_t row_base,
                      uint64_t* out_rows, float* out_dists, uint32_t* out_counts) {
    if (idx && !dense_rows(idx)) return fail(LB_INVALID_ARGUMENT, "use lb_sharded_search_packed for a packed index");
    return sharded_search_host(comm, idx, metric, queries, idx ? (size_t)idx->dim * 4 : 0, nq, k, row_base, nullptr, 0, out_rows, out_dists, out_counts);
}

int lb_sharded_search_filtered(lb_comm* comm, lb_index* idx, int metric, const float* queries, uint32_t nq, uint32_t k, uint64_t row_base,
                               const uint64_t* allow_bits, uint64_t allow_words, uint64_t* out_rows, float* out_dists, uint32_t* out_counts) {
    if (idx && !dense_rows(idx)) return fail(LB_INVALID_ARGUMENT, "use lb_sharded_search_packed for a packed index");
    return sharded_search_host(comm, idx, metric, queries, idx ? (size_t)idx->dim * 4 : 0, nq, k, row_base, allow_bits, allow_words, out_rows,
                               out_dists, out_counts);
}

int lb_sharded_search_packed(lb_comm* comm, lb_index* idx, int metric, const uint64_t* query_words, uint32_t nq, uint32_t k,
                             uint64_t row_base, uint64_t* out_rows, float* out_dists, uint32_t* out_counts) {
    if (idx && idx->dtype != LB_PACKED_U64) return fail(LB_INVALID_ARGUMENT, "index does not store packed rows");
    return sharded_search_host(comm, idx, metric, query_words, idx ? (size_t)idx->n_words * 8 : 0, nq, k, row_base, nullptr, 0, out_rows, out_dists,
                               out_counts);
}

// Test hook for the shard merge: G host blocks of [nq][k] (u32 local rows, f32 scores, counts) with their row bases ->
// merge_shards_kernel on `device` -> [nq][k] global rows / scores / counts.  No communicator involved: this is the
// kernel the multi-GPU path runs after its all-gather, on blocks the caller supplies.
int lb_merge_shard_blocks(int device, int metric, int n_shards, uint32_t nq, uint32_t k, const uint32_t* rows, const float* dists,
                          const uint32_t* counts, const uint64_t* row_bases, uint64_t* out_rows, float* out_dists, uint32_t* out_counts) {
    LB_TRY(check_metric(metric));
    if (n_shards < 1 || nq == 0 || k == 0 || !rows || !dists || !counts || !row_bases || !out_rows || !out_dists || !out_counts)
        return fail(LB_INVALID_ARGUMENT, "bad arguments");
    if ((uint64_t)n_shards * k > 4096) return fail(LB_UNSUPPORTED, "n_shards * k must not exceed 4096");
    DeviceGuard g(device);
    const size_t bb = shard_block_bytes(nq, k);
    std::vector<unsigned char> host((size_t)n_shards * bb, 0);
    for (int s = 0; s < n_shards; ++s) {
        unsigned char* blk = host.data() + (size_t)s * bb;
        memcpy(blk, rows + (size_t)s * nq * k, (size_t)nq * k * 4);
        memcpy(blk + (size_t)nq * k * 4, dists + (size_t)s * nq * k, (size_t)nq * k * 4);
        memcpy(blk + (size_t)nq * k * 8, counts + (size_t)s * nq, (size_t)nq * 4);
        memcpy(blk + bb - SHARD_TAIL_BYTES, row_bases + s, 8);
    }
    unsigned char* d_blocks = nullptr;
    uint64_t* d_rows = nullptr;
    float* d_dists = nullptr;
    uint32_t* d_counts = nullptr;
    cudaError_t e = cudaMalloc(&d_blocks, host.size());
    if (e == cudaSuccess) e = cudaMalloc(&d_rows, (size_t)nq * k * 8);
    if (e == cudaSuccess) e = cudaMalloc(&d_dists, (size_t)nq * k * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_counts, (size_t)nq * 4);
    if (e == cudaSuccess) e = cudaMemcpy(d_blocks, host.data(), host.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        const int M = std::max(2, next_pow2(n_shards * (int)k));
        merge_shards_kernel<<<nq, 256, (size_t)M * 8>>>(d_blocks, n_shards, bb, (int)nq, (int)k, M, metric_ascending(metric) ? 1 : 0, d_rows, d_dists,
                                                       d_counts);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out_rows, d_rows, (size_t)nq * k * 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(out_dists, d_dists, (size_t)nq * k * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(out_counts, d_counts, (size_t)nq * 4, cudaMemcpyDeviceToHost);
    cudaFree(d_blocks);
    cudaFree(d_rows);
    cudaFree(d_dists);
    cudaFree(d_counts);
    if (e != cudaSuccess) return fail(LB_CUDA, std::string("lb_merge_shard_blocks: ") + cudaGetErrorString(e));
    return LB_OK;
}

}  // extern "C"

extern "C" {

int lb_index_event_record(lb_index* idx, int slot) {
    if (!idx || slot < 0 || slot >= 8) return fail(LB_INVALID_ARGUMENT, "bad event slot");
    DeviceGuard g(idx->device);
    LB_CUDA_TRY(cudaEventRecord(idx->user_ev[slot], idx->stream));
    return LB_OK;
}

int lb_index_event_elapsed_ms(lb_index* idx, int slot_a, int slot_b, float* ms) {
    if (!idx || !ms || slot_a < 0 || slot_a >= 8 || slot_b < 0 || slot_b >= 8) return fail(LB_INVALID_ARGUMENT, "bad event slot");
    DeviceGuard g(idx->device);
    LB_CUDA_TRY(cudaEventSynchronize(idx->user_ev[slot_b]));
    LB_CUDA_TRY(cudaEventElapsedTime(ms, idx->user_ev[slot_a], idx->user_ev[slot_b]));
    return LB_OK;
}


}  // extern "C"
