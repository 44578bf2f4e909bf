// lb_host.cuh — host-side state shared by the translation units of liblynse_b200.so: growable device / pinned
// buffers, the index object, and the entry points each unit exports to the others.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include <cuda_fp16.h>

#include "lb_common.cuh"

namespace lb {


// Growable device buffer
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes, bool keep = false, cudaStream_t stream = 0) {
        if (bytes <= cap) return LB_OK;
        size_t ncap = keep ? std::max(bytes, cap + cap / 2) : bytes;
        void* np = nullptr;
        cudaError_t e = cudaMalloc(&np, ncap);
        if (e != cudaSuccess && ncap != bytes) {
            ncap = bytes;
            e = cudaMalloc(&np, ncap);
        }
        if (e != cudaSuccess)
            return fail(LB_CUDA, std::string("cudaMalloc(") + std::to_string(ncap) + "): " + cudaGetErrorString(e));
        if (keep && p && cap) {
            e = cudaMemcpyAsync(np, p, cap, cudaMemcpyDeviceToDevice, stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
            if (e != cudaSuccess) {
                cudaFree(np);
                return fail(LB_CUDA, std::string("grow copy: ") + cudaGetErrorString(e));
            }
        }
        if (p) cudaFree(p);
        p = np;
        cap = ncap;
        return LB_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
};

// pinned host staging memory (small batches copy through it: a copy from pageable memory is staged and
// synchronised by the driver, ~10 us a call)
struct HostBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return LB_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        const size_t ncap = std::max<size_t>(bytes, 64 * 1024);
        if (cudaHostAlloc(&p, ncap, cudaHostAllocDefault) != cudaSuccess) {
            p = nullptr;
            return fail(LB_CUDA, "cudaHostAlloc of the staging buffer failed");
        }
        cap = ncap;
        return LB_OK;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    HostBuf() = default;
    HostBuf(const HostBuf&) = delete;
    HostBuf& operator=(const HostBuf&) = delete;
    ~HostBuf() { release(); }
};

struct Shadow {
    DevBuf buf;           // tiled, pre-swizzled operand rows (layout: lb_tc.cuh)
    DevBuf side;          // per-row side values, padded to whole tiles: f32 |c|^2 (L2 shadow), u32 popcount (bit shadow)
    DevBuf side2;         // bit shadow: the popcounts as f32 (Jaccard / Dice keys)
    DevBuf stats;         // tc::ShadowStats on the device
    uint64_t rows = 0;    // rows converted so far
    uint64_t cap_tiles = 0;
    int Dp = 0;           // operand row length in 2-byte units (bf16: padded dim; 8-bit: padded dim / 2)
    int operand = -1;     // tc::OperandKind, chosen when the shadow is first built
    float c_scale = 1.0f, c_zero = 0.0f;  // 8-bit operands: value = c_zero + c_scale * u8
    float range_lo = 0.0f, range_hi = 0.0f;  // element range the quantisation was built for
    float cmax = 0.0f, emax = 0.0f;          // host copies of the statistics (diagnostics)
    bool disabled = false;                  // non-finite rows: the plan is not used
    bool l2_bias = false;                   // L2 shadow: |c|^2 as an f32 side value (no room for the three norm columns)
    // [0] one-CTA kernel (both halves of a K block per box), [1] CTA-pair kernel (one half); .full = KPS K blocks per
    // box, .rem = the partial last stage of a tile ((Dp/64) % KPS K blocks)
    CUtensorMap tmap_full[2], tmap_rem[2];
    uint64_t tmap_tiles = 0;
    void* tmap_ptr = nullptr;
    void release() {
        buf.release();
        side.release();
        side2.release();
        stats.release();
        rows = 0;
        cap_tiles = 0;
        tmap_tiles = 0;
        tmap_ptr = nullptr;
        operand = -1;
    }
};

inline int next_pow2(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}
inline int tc_env_int(const char* name, int dflt) {
    const char* env = getenv(name);
    return env && *env ? atoi(env) : dflt;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) costs a few microseconds per call: remember, per kernel and device,
// the largest size already granted and only raise it.
template <class K>
static cudaError_t ensure_dynamic_smem(K kernel, int bytes) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, int> granted;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    int& have = granted[{reinterpret_cast<const void*>(kernel), dev}];
    if (bytes <= have) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) have = bytes;
    return e;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

}  // namespace lb

using lb::DevBuf;
using lb::HostBuf;
using lb::Shadow;

struct lb_index {
    int device = 0;
    uint32_t dim = 0;
    int dtype = LB_F32;
    int n_words = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    DevBuf rows;  // f32 [n][dim] or u64 [n][n_words]
    uint64_t n = 0;
    std::vector<uint64_t> segments;
    uint64_t seg_target = 256ull * 1024 * 1024;
    bool force_new_segment = false;  // the next append opens a segment even if it would fit the last one (lb_index_new_segment)
    // side structures (derived caches, extended incrementally after appends)
    DevBuf packed;
    uint64_t packed_rows = 0;
    DevBuf js_stats;
    uint64_t js_rows = 0;
    DevBuf mass_stats;       // Wasserstein: f64 row sums, rows [0, mass_rows)
    uint64_t mass_rows = 0;
    Shadow shadow[3];        // IP, cosine, L2 (tc::ShadowKind)
    Shadow bits_shadow;      // the packed rows as {0,1} bytes (binary metrics on the tensor cores)
    DevBuf max_norm;  // 3 floats, one per shadow kind
    DevBuf small_seg;
    int n_small = 0;
    size_t small_seg_sig = ~(size_t)0;
    // workspace
    DevBuf w_queries, w_qwords, w_allow, w_lists, w_counts, w_thr, w_out_rows, w_out_dists, w_out_counts, w_qtiles, w_pbest;
    DevBuf w_out;            // host-buffer searches: [rows | dists | counts] of one batch, so that one copy brings them back
    HostBuf h_in, h_out;     // pinned staging for small batches
    HostBuf h_tails;         // sharded search: the gathered block tails (one read per step)
    DevBuf w_qb, w_qnorm, w_cand_score, w_cand_row, w_cand_thr, w_flags, w_qstats, w_nq, w_sub_q, w_qmap;
    int plan = LB_PLAN_AUTO;
    // how the running search scores a pair (set under `mu` for the duration of one host-buffer search):
    // SCORE_FLAT = the FLAT scan's kernels (tensor-core plan allowed), SCORE_PAIRWISE = compute_distance_f32 on the
    // exact scan, SCORE_F16_ROWS = compute_distance_f16 (scalar order) on the exact scan
    int score_mode = 0;
    uint64_t allow_count = 0;  // rows allowed by the filter of the running host-buffer search (set with the filter)
    bool timing = false;
    lb_search_stats stats{};
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t user_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    DevBuf w_send, w_recv, w_g_rows, w_g_dists, w_g_counts, w_progress, w_prof, w_gfloor;
    DevBuf w_qaux, w_qrange, w_hits, w_hit_count;
    // a tensor-core search whose certification flags have not been read back yet (tc_finish)
    struct PendingTc {
        bool active = false;
        int metric = 0, nq = 0, k = 0, bits = 0, n_words = 0;
        const void* d_queries = nullptr;
        const uint64_t* d_allow = nullptr;
        const uint64_t* words = nullptr;
        uint32_t* d_rows = nullptr;
        float* d_dists = nullptr;
        uint32_t* d_counts = nullptr;
        int grid = 0, cluster = 0, n_slots = 0, P = 0;
        bool pair = false;
    } pending_tc;
};

namespace lb {

inline size_t row_bytes(const lb_index* idx) {
    return idx->dtype == LB_F32 ? (size_t)idx->dim * 4 : (idx->dtype == LB_F16 ? (size_t)idx->dim * 2 : (size_t)idx->n_words * 8);
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        int cur = -1;
        cudaGetDevice(&cur);
        if (prev >= 0 && cur != prev) cudaSetDevice(prev);
    }
};

struct ScanRequest {
    const float* corpus = nullptr;
    const __half* corpus_h = nullptr;   // binary16 rows (LB_F16 index) instead of `corpus`
    const uint64_t* words = nullptr;
    uint64_t n_rows = 0;
    int dim = 0, n_words = 0;
    const float* queries = nullptr;
    const uint64_t* qwords = nullptr;
    int nq = 0, k = 0, metric = 0;
    const uint64_t* allow_bits = nullptr;
    const uint32_t* row_ids = nullptr;
    const uint32_t* small_seg = nullptr;
    int n_small = 0;
    int ip_single = 0;
    int f16_rows = 0;  // score with compute_distance_f16's scalar kernels (old exact kernel only)
    const float* row_stats = nullptr;
    const float* query_stats = nullptr;
    const double* row_mass = nullptr;  // Wasserstein: f64 row sums (enables the streaming scan for it)
    int sqrt_scores = 0;
    const uint32_t* qmap = nullptr;
    uint32_t* out_rows = nullptr;
    float* out_dists = nullptr;
    uint32_t* out_counts = nullptr;
};

// lb_scan_plan.cu: exact scan + merge on idx->stream (k <= n_rows, k <= 2048)
int run_scan(lb_index* idx, const ScanRequest& r, int* kernels, float* ms_dom);
// lynse_b200.cu
int refresh_small_segments(lb_index* idx);
// lb_tc_plan.cu: tensor-core coarse pass + exact rescore (IP / cosine / L2)
int shadow_kind_for(int metric);
bool tc_supported(lb_index* idx, int metric);
int ensure_shadow(lb_index* idx, int kind);
// Enqueues prepare -> coarse -> finalize on idx->stream.  defer_check: the certification flags are not read back here;
// the caller must call tc_finish (after whatever it enqueues behind the search) before it trusts the results.
int run_tc(lb_index* idx, int metric, const float* d_queries, int nq, int k, uint32_t* d_rows, float* d_dists, uint32_t* d_counts,
           float* dump, const uint64_t* d_allow = nullptr, bool defer_check = false);
// the binary metrics (Hamming / Jaccard / Tanimoto / Dice) over packed rows as a {0,1} 8-bit contraction
bool tc_bits_supported(lb_index* idx, int metric, int n_words, int nq, int k);
int run_tc_bits(lb_index* idx, int metric, const uint64_t* words, int n_words, const uint64_t* d_qwords, int nq, int k, uint32_t* d_rows,
                float* d_dists, uint32_t* d_counts, const uint64_t* d_allow = nullptr, bool defer_check = false);
// Synchronises the stream, reads the flags of the pending tensor-core search and re-runs uncertified queries with the
// exact scan.  *changed (optional) = results were rewritten (anything enqueued behind the search saw stale results).
// `head`: the first four flag words, already copied to the host behind the search and synchronised (saves the copy + wait)
int tc_finish(lb_index* idx, bool* changed = nullptr, const uint32_t* head = nullptr);

}  // namespace lb
