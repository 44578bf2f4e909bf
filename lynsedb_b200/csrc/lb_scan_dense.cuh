// lb_scan_dense.cuh — launches of the dense exact scans (lb_scan.cuh, lb_scan2.cuh) for one row type: f32 rows, or the
// binary16 rows of a float16 index (the same kernels instantiated with RT = __half; every load decodes, exactly).
// Each row type is instantiated in its own translation unit (lb_scan_plan.cu, lb_scan_f16.cu) to keep the builds short.
#pragma once
#include "lb_host.cuh"
#include "lb_metrics.cuh"
#include "lb_scan.cuh"
#include "lb_scan2.cuh"
#include "lb_scan3.cuh"

namespace lb {

struct ScanPlan {
    int P;
    uint32_t rows_per_part;
};

template <class RT>
int dense_scan_launch(lb_index* idx, const ScanRequest& r, ScanArgs& a, const ScanPlan& sp, bool s2_metric, bool tma_rows, bool tile_rows) {
    if (tile_rows) {
        // a batch of queries over contiguous f32 rows: register tiles of 4 rows x 16 queries, rows resident in shared memory (lb_scan3.cuh)
        if constexpr (std::is_same<RT, float>::value) {
            PFN_encodeTiled enc = get_encode_tiled();
            if (!enc) return fail(LB_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
            const int nw = scan4_warps(r.dim);
            CUtensorMap tmap;
            cuuint64_t gdim[2] = {(cuuint64_t)r.dim, (cuuint64_t)r.n_rows};
            cuuint64_t gstride[1] = {(cuuint64_t)r.dim * sizeof(float)};
            cuuint32_t box[2] = {32u, (cuuint32_t)(scan4_block_rows(nw))};
            cuuint32_t estr[2] = {1, 1};
            CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)const_cast<float*>(r.corpus), gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (cr != CUDA_SUCCESS) return fail(LB_CUDA, "cuTensorMapEncodeTiled (row tiles) failed with CUresult " + std::to_string((int)cr));
            const bool ip2 = r.ip_single || r.n_small > 0;
#define LB_LAUNCH_S4T(M, IP2V, NWV, TQV, QSPV)                                                             \
    do {                                                                                                   \
        using Cfg = S4Cfg<M, IP2V, TQV, QSPV>;                                                             \
        const int n_tiles = (r.nq + Cfg::kTQ - 1) / Cfg::kTQ;                                              \
        LB_TRY(idx->w_qtiles.ensure((size_t)n_tiles * Cfg::tile_floats(r.dim) * 4));                       \
        scan4_query_tiles_kernel<<<n_tiles, 256, 0, idx->stream>>>(r.queries, r.nq, r.dim, Cfg::kTQ, Cfg::kQStride, idx->w_qtiles.as<float>()); \
        a.query_tiles = idx->w_qtiles.as<float>();                                                         \
        a.smem_lists = Cfg::smem_lists(r.nq, r.k) ? 1 : 0;                                                 \
        const size_t smem = Cfg::smem_bytes(NWV, r.dim, r.nq, r.k);                                        \
        LB_CUDA_TRY(ensure_dynamic_smem(scan_tile_kernel<M, IP2V, NWV, TQV, QSPV>, (int)smem));            \
        scan_tile_kernel<M, IP2V, NWV, TQV, QSPV><<<sp.P, NWV * 32, smem, idx->stream>>>(tmap, a);         \
    } while (0)
// up to 256 dims, 64-row blocks: the one-accumulator metrics (IP, L1, Chebyshev) run eight warps, two per row group, each
// with eight of the tile's sixteen queries (16 warps per SM); the two-accumulator ones (L2, cosine, Bray-Curtis; IP over
// small segments) four warps with eight queries per tile — measured at 64 queries over 4M x 256: L1 6.80 -> 6.35 ms,
// Chebyshev 11.3 -> 9.9 with the eight warps, but L2 8.7 -> 10.0 and Bray-Curtis 12.5 -> 14.0 (64 accumulators under the
// 128-register cap of a 256-thread CTA; up to 128 dims they take the eight warps too: L2 11.8 -> 10.7 ms over 8M x 128).
// Up to 512 dims: two warps, 32-row blocks, eight queries per tile.
#define LB_LAUNCH_S4(M, IP2V)                                                            \
    do {                                                                                 \
        constexpr bool one_acc = Scan2Op<M, IP2V>::kState == 8;                          \
        if (nw == 8 && (one_acc || r.dim <= 128)) LB_LAUNCH_S4T(M, IP2V, 8, 16, 2);      \
        else if (nw == 8) LB_LAUNCH_S4T(M, IP2V, (one_acc ? 8 : 4), (one_acc ? 16 : 8), (one_acc ? 2 : 1)); \
        else LB_LAUNCH_S4T(M, IP2V, 2, 8, 1);                                            \
    } while (0)
            switch (r.metric) {
                case LB_IP: if (ip2) LB_LAUNCH_S4(LB_IP, true); else LB_LAUNCH_S4(LB_IP, false); break;
                case LB_L2: LB_LAUNCH_S4(LB_L2, false); break;
                case LB_COSINE: LB_LAUNCH_S4(LB_COSINE, false); break;
                case LB_MANHATTAN: LB_LAUNCH_S4(LB_MANHATTAN, false); break;
                case LB_CHEBYSHEV: LB_LAUNCH_S4(LB_CHEBYSHEV, false); break;
                default: LB_LAUNCH_S4(LB_BRAY_CURTIS, false); break;
            }
#undef LB_LAUNCH_S4
#undef LB_LAUNCH_S4T
        } else {
            return fail(LB_INTERNAL, "the row-tile scan serves f32 rows");
        }
    } else if (s2_metric && !r.f16_rows && tc_env_int("LYNSE_B200_SCAN2", 1) != 0 &&
               (size_t)8 * ((r.dim + 3) & ~3) * 4 + 8 * S2_ROWS * 8 + 256 <= 200 * 1024) {
        // streaming scan: the row is read once per query tile (lb_scan2.cuh)
        const int dim_pad = (r.dim + 3) & ~3;
        const bool ip2 = r.ip_single || r.n_small > 0;
        // contiguous rows of a 16-byte-multiple width go through TMA-staged shared memory
        // (measured on 10M x 768: 4.55 TB/s against 3.74 TB/s at one query, 8.8 against 9.4 ms at four; from eight
        // queries on the pass is bound by the shared-memory reads of the queries and the direct version is as fast)
        const bool use_tma = tma_rows;
        if (use_tma) {
            PFN_encodeTiled enc = get_encode_tiled();
            if (!enc) return fail(LB_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
            CUtensorMap tmap;
            cuuint64_t gdim[2] = {(cuuint64_t)r.dim, (cuuint64_t)r.n_rows};
            cuuint64_t gstride[1] = {(cuuint64_t)r.dim * sizeof(RT)};
            cuuint32_t box[2] = {(cuuint32_t)(128 / sizeof(RT)), (cuuint32_t)S2_ROWS};
            cuuint32_t estr[2] = {1, 1};
            const bool half_rows = !std::is_same<RT, float>::value;
            void* base = half_rows ? (void*)const_cast<__half*>(r.corpus_h) : (void*)const_cast<float*>(r.corpus);
            CUresult cr = enc(&tmap, half_rows ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (cr != CUDA_SUCCESS) return fail(LB_CUDA, "cuTensorMapEncodeTiled (f32 rows) failed with CUresult " + std::to_string((int)cr));
#define LB_LAUNCH_S3(M, IP2V)                                                                                                 \
    do {                                                                                                                      \
        constexpr int tqv = Scan2Op<M, IP2V>::kTQ < S3_TQ ? Scan2Op<M, IP2V>::kTQ : S3_TQ;                                    \
        a.smem_lists = (r.nq <= tqv && r.k <= 256) ? 1 : 0;                                                                   \
        const size_t smem = (size_t)S3Cfg<RT>::kStages * S3_STAGE_BYTES + (size_t)tqv * S2_ROWS * 8 + (size_t)tqv * dim_pad * 4 +  \
                            (size_t)tqv * 32 + 128 + 1024 + (a.smem_lists ? (size_t)tqv * r.k * 8 + tqv * 4 + 16 : 0);         \
        LB_CUDA_TRY(ensure_dynamic_smem(scan_stream_tma_kernel<M, IP2V, RT>, (int)smem)); \
        scan_stream_tma_kernel<M, IP2V, RT><<<sp.P, S2_ROWS, smem, idx->stream>>>(tmap, a);                                       \
    } while (0)
            switch (r.metric) {
                case LB_IP: if (ip2) LB_LAUNCH_S3(LB_IP, true); else LB_LAUNCH_S3(LB_IP, false); break;
                case LB_L2: LB_LAUNCH_S3(LB_L2, false); break;
                case LB_COSINE: LB_LAUNCH_S3(LB_COSINE, false); break;
                case LB_MANHATTAN: LB_LAUNCH_S3(LB_MANHATTAN, false); break;
                case LB_CHEBYSHEV: LB_LAUNCH_S3(LB_CHEBYSHEV, false); break;
                case LB_CANBERRA: LB_LAUNCH_S3(LB_CANBERRA, false); break;
                default: LB_LAUNCH_S3(LB_BRAY_CURTIS, false); break;
            }
#undef LB_LAUNCH_S3
        } else {
#define LB_LAUNCH_S2(M, IP2V)                                                                                             \
    do {                                                                                                                  \
        constexpr int tqv = Scan2Op<M, IP2V>::kTQ;                                                                        \
        a.smem_lists = (r.nq <= tqv && r.k <= 256) ? 1 : 0;                                                               \
        const size_t smem = (size_t)tqv * S2_ROWS * 8 + (size_t)tqv * dim_pad * 4 + (size_t)tqv * 32 + 64 +                \
                            (a.smem_lists ? (size_t)tqv * r.k * 8 + tqv * 4 + 16 : 0);                                     \
        LB_CUDA_TRY(ensure_dynamic_smem(scan_stream_kernel<M, IP2V, RT>, (int)smem)); \
        scan_stream_kernel<M, IP2V, RT><<<sp.P, S2_ROWS, smem, idx->stream>>>(a);                                             \
    } while (0)
        switch (r.metric) {
            case LB_IP: if (ip2) LB_LAUNCH_S2(LB_IP, true); else LB_LAUNCH_S2(LB_IP, false); break;
            case LB_L2: LB_LAUNCH_S2(LB_L2, false); break;
            case LB_COSINE: LB_LAUNCH_S2(LB_COSINE, false); break;
            case LB_MANHATTAN: LB_LAUNCH_S2(LB_MANHATTAN, false); break;
            case LB_CHEBYSHEV: LB_LAUNCH_S2(LB_CHEBYSHEV, false); break;
            case LB_CANBERRA: LB_LAUNCH_S2(LB_CANBERRA, false); break;
            case LB_CORRELATION: LB_LAUNCH_S2(LB_CORRELATION, false); break;
            case LB_HELLINGER: LB_LAUNCH_S2(LB_HELLINGER, false); break;
            case LB_WASSERSTEIN: LB_LAUNCH_S2(LB_WASSERSTEIN, false); break;
            case LB_JENSEN_SHANNON: LB_LAUNCH_S2(LB_JENSEN_SHANNON, false); break;
            default: LB_LAUNCH_S2(LB_BRAY_CURTIS, false); break;
        }
        }
#undef LB_LAUNCH_S2
    } else {
        int dim_pad = (r.dim + 3) & ~3;
        size_t smem = (size_t)SCAN_TQ * SCAN_THREADS * 8 + (size_t)SCAN_TQ * dim_pad * 4;
        if (smem > 200 * 1024) return fail(LB_UNSUPPORTED, "dimension above 2560 is not supported by the exact scan");
        if (metric_ascending(r.metric)) {
            LB_CUDA_TRY(ensure_dynamic_smem(scan_exact_kernel<true, RT>, (int)smem));
            scan_exact_kernel<true, RT><<<sp.P, SCAN_THREADS, smem, idx->stream>>>(a);
        } else {
            LB_CUDA_TRY(ensure_dynamic_smem(scan_exact_kernel<false, RT>, (int)smem));
            scan_exact_kernel<false, RT><<<sp.P, SCAN_THREADS, smem, idx->stream>>>(a);
        }
    }
    return LB_OK;
}

extern template int dense_scan_launch<float>(lb_index*, const ScanRequest&, ScanArgs&, const ScanPlan&, bool, bool, bool);
extern template int dense_scan_launch<__half>(lb_index*, const ScanRequest&, ScanArgs&, const ScanPlan&, bool, bool, bool);

}  // namespace lb
