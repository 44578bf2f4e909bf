// lb_scan_plan.cu — the exact CUDA-core scans (lb_scan.cuh, lb_scan2.cuh, lb_packed.cuh): plan + launches + merge.
#include "lb_scan_dense.cuh"
#include "lb_packed.cuh"

using namespace lb;

namespace lb {

template int dense_scan_launch<float>(lb_index*, const ScanRequest&, ScanArgs&, const ScanPlan&, bool, bool, bool);

// ---- exact scan plan ------------------------------------------------------------------------------------------
static ScanPlan plan_scan(const lb_index* idx, uint64_t n_rows, int nq, int k, int ctas_per_sm = 2) {
    uint64_t P = std::min<uint64_t>(ceil_div(n_rows, SCAN_THREADS), (uint64_t)idx->sm_count * ctas_per_sm);
    // bound the candidate lists to 512 MiB
    uint64_t max_p = std::max<uint64_t>(1, (512ull << 20) / ((uint64_t)nq * k * 8));
    P = std::max<uint64_t>(1, std::min(P, max_p));
    uint64_t rpp = ceil_div(ceil_div(n_rows, P), SCAN_THREADS) * SCAN_THREADS;
    ScanPlan p;
    p.rows_per_part = (uint32_t)rpp;
    p.P = (int)ceil_div(n_rows, rpp);
    return p;
}

// scan + merge on idx->stream (k <= n_rows, k <= 2048)
int run_scan(lb_index* idx, const ScanRequest& r, int* kernels, float* ms_dom) {
    // the TMA-staged f32 scan runs one CTA per SM (its ring takes the shared memory): one partition per SM
    const bool s2_metric = scan2_supported(r.metric) && (r.metric != LB_WASSERSTEIN || r.row_mass != nullptr) &&
                           (r.row_stats != nullptr) == (r.metric == LB_JENSEN_SHANNON);  // Jensen-Shannon: the cached form only
    // (the f64 metrics are bound by arithmetic latency, not by the load pattern: two resident CTAs per SM beat the
    // one-CTA TMA ring for them — Hellinger 2.1 against 1.3 TB/s, Wasserstein 3.4 against 2.4 at one query)
    const bool tma_f32 = !r.words && !r.f16_rows && s2_metric && !scan2_f64(r.metric) && r.metric != LB_JENSEN_SHANNON && tc_env_int("LYNSE_B200_SCAN2", 1) != 0 &&
                         r.row_ids == nullptr && (r.dim & (r.corpus_h != nullptr ? 7 : 3)) == 0 && r.dim >= 8 && r.n_rows >= 4096 && r.nq <= 4 &&
                         tc_env_int("LYNSE_B200_SCAN_TMA", 1) != 0 &&
                         (size_t)S3_NSTAGES * S3_STAGE_BYTES + (size_t)S3_TQ * (S2_ROWS * 8 + ((r.dim + 3) & ~3) * 4 + 256 * 8) + 2048 <= 226 * 1024;
    // a dozen queries and more over contiguous f32 rows of up to 512 dims: the row-tile scan (lb_scan3.cuh).  Below that the
    // streaming scan is HBM-bound anyway; above 512 dims the resident row block leaves too few warps per SM.
    const bool tile_f32 = !r.words && !r.f16_rows && r.corpus != nullptr && r.corpus_h == nullptr && s2_metric && scan4_supported(r.metric) &&
                          r.row_ids == nullptr && (r.dim & 3) == 0 && r.dim >= 8 && r.dim <= 512 && r.n_rows >= 4096 &&
                          r.nq >= tc_env_int("LYNSE_B200_SCAN_TILE_MIN_Q", 12) && tc_env_int("LYNSE_B200_SCAN_TILE", 1) != 0;
    int ctas_per_sm = (tma_f32 && r.corpus_h == nullptr) ? 1 : 2;
    if (tile_f32) ctas_per_sm = std::max(1, std::min(tc_env_int("LYNSE_B200_SCAN_TILE_CTAS", 2), 4));
    ScanPlan sp = plan_scan(idx, r.n_rows, r.nq, r.k, ctas_per_sm);
    size_t nl = (size_t)sp.P * r.nq;
    LB_TRY(idx->w_lists.ensure(nl * r.k * 8));
    LB_TRY(idx->w_counts.ensure(nl * 4));
    LB_TRY(idx->w_thr.ensure(nl * 8));
    LB_CUDA_TRY(cudaMemsetAsync(idx->w_counts.p, 0, nl * 4, idx->stream));
    ScanArgs a{};
    a.corpus = r.corpus;
    a.corpus_h = r.corpus_h;
    const bool f16 = r.corpus_h != nullptr;   // binary16 rows: the kernels are instantiated with RT = __half
    a.words = r.words;
    a.n_rows = (uint32_t)r.n_rows;
    a.dim = r.dim;
    a.n_words = r.n_words;
    a.queries = r.queries;
    a.qwords = r.qwords;
    a.nq = r.nq;
    a.k = r.k;
    a.metric = r.metric;
    a.allow_bits = r.allow_bits;
    a.row_ids = r.row_ids;
    a.small_seg = r.small_seg;
    a.n_small = r.n_small;
    a.ip_single = r.ip_single;
    a.f16_rows = r.f16_rows;
    a.row_mass = r.row_mass;
    a.row_stats = r.row_stats;
    a.query_stats = r.query_stats;
    a.lists = idx->w_lists.as<uint64_t>();
    a.counts = idx->w_counts.as<uint32_t>();
    a.thr = idx->w_thr.as<uint64_t>();
    a.rows_per_part = sp.rows_per_part;
    if (idx->timing) cudaEventRecord(idx->ev[0], idx->stream);
    if (r.words && r.n_words == PK_W && r.row_ids == nullptr && tc_env_int("LYNSE_B200_PACKED_TMA", 1) != 0) {
        // 1024-bit fingerprints: TMA-staged kernel (lb_packed.cuh)
        PFN_encodeTiled enc = get_encode_tiled();
        if (!enc) return fail(LB_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
        CUtensorMap tmap;
        cuuint64_t gdim[2] = {(cuuint64_t)PK_W, (cuuint64_t)r.n_rows};
        cuuint64_t gstride[1] = {(cuuint64_t)PK_W * 8};
        cuuint32_t box[2] = {(cuuint32_t)PK_W, (cuuint32_t)PK_ROWS};
        cuuint32_t estr[2] = {1, 1};
        CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<uint64_t*>(r.words), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return fail(LB_CUDA, "cuTensorMapEncodeTiled (packed rows) failed with CUresult " + std::to_string((int)cr));
        const bool hs = tc_env_int("LYNSE_B200_PACKED_HS", 0) != 0;
        const int mode = r.metric == LB_HAMMING ? 0 : (r.metric == LB_DICE ? 2 : 1);
#define LB_LAUNCH_PK(MODE, HSV)                                                                                              \
    do {                                                                                                                     \
        a.smem_lists = (r.nq <= PK_TQ && r.k <= 256) ? 1 : 0;                                                                \
        const size_t pk_smem = PK_SMEM_BYTES + (a.smem_lists ? (size_t)PK_TQ * r.k * 8 + PK_TQ * 4 + 16 : 0);                \
        LB_CUDA_TRY(ensure_dynamic_smem(scan_packed16_kernel<MODE, HSV>, (int)pk_smem));                                             \
        scan_packed16_kernel<MODE, HSV><<<sp.P, PK_ROWS, pk_smem, idx->stream>>>(tmap, a);                                   \
    } while (0)
        if (mode == 0) { if (hs) LB_LAUNCH_PK(0, true); else LB_LAUNCH_PK(0, false); }
        else if (mode == 1) { if (hs) LB_LAUNCH_PK(1, true); else LB_LAUNCH_PK(1, false); }
        else { if (hs) LB_LAUNCH_PK(2, true); else LB_LAUNCH_PK(2, false); }
#undef LB_LAUNCH_PK
    } else if (r.words) {
        size_t smem = (size_t)SCAN_TQ * SCAN_THREADS * 8 + (size_t)SCAN_TQ * r.n_words * 8;
#define LB_LAUNCH_PACKED(W)                                                                                      \
    do {                                                                                                         \
        LB_CUDA_TRY(ensure_dynamic_smem(scan_packed_kernel<W>, (int)smem)); \
        scan_packed_kernel<W><<<sp.P, SCAN_THREADS, smem, idx->stream>>>(a);                                     \
    } while (0)
        if (smem > 200 * 1024) return fail(LB_UNSUPPORTED, "packed rows wider than 1280 words are not supported");
        switch (r.n_words) {
            case 1: LB_LAUNCH_PACKED(1); break;
            case 2: LB_LAUNCH_PACKED(2); break;
            case 4: LB_LAUNCH_PACKED(4); break;
            case 8: LB_LAUNCH_PACKED(8); break;
            case 16: LB_LAUNCH_PACKED(16); break;
            default: LB_LAUNCH_PACKED(0); break;
        }
#undef LB_LAUNCH_PACKED
    } else if (f16) {
        LB_TRY(dense_scan_launch<__half>(idx, r, a, sp, s2_metric, tma_f32, false));
    } else {
        LB_TRY(dense_scan_launch<float>(idx, r, a, sp, s2_metric, tma_f32 && !tile_f32, tile_f32));
    }
    LB_CUDA_TRY(cudaGetLastError());
    if (idx->timing) cudaEventRecord(idx->ev[1], idx->stream);
    MergeArgs m{};
    m.lists = a.lists;
    m.counts = a.counts;
    m.P = sp.P;
    m.nq = r.nq;
    m.k = r.k;
    m.kp2 = next_pow2(r.k);
    long total = (long)sp.P * r.k;
    m.M = (int)std::min<long>(4096, std::max<long>(2L * m.kp2, next_pow2((int)std::min<long>(4096, m.kp2 + total))));
    m.asc = metric_ascending(r.metric) ? 1 : 0;
    m.sqrt_scores = r.sqrt_scores;
    m.qmap = r.qmap;
    m.out_rows = r.out_rows;
    m.out_dists = r.out_dists;
    m.out_counts = r.out_counts;
    // few queries: few blocks, so each gets 1024 threads (the block-wide bitonic sort is the latency of a single-query call)
    if (r.k <= MERGE_SMALL_MAX_K && total <= MERGE_SMALL_MAX_KEYS)
        merge_lists_small_kernel<<<r.nq, 1024, 0, idx->stream>>>(m);
    else
        merge_lists_kernel<<<r.nq, r.nq <= 64 ? 1024 : 256, (size_t)m.M * 8, idx->stream>>>(m);
    LB_CUDA_TRY(cudaGetLastError());
    if (kernels) *kernels += tile_f32 ? 3 : 2;
    idx->stats.n_partitions = sp.P;
    if (idx->timing && ms_dom) {
        LB_CUDA_TRY(cudaEventSynchronize(idx->ev[1]));
        float ms = 0;
        cudaEventElapsedTime(&ms, idx->ev[0], idx->ev[1]);
        *ms_dom = ms;
    }
    return LB_OK;
}

}  // namespace lb
