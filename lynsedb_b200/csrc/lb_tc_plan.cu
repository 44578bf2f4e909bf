// lb_tc_plan.cu — host side of the tensor-core plan: shadow upkeep, the coarse pass + finalize launches, the
// exact-scan fallback of uncertified queries, and the tcgen05 diagnostics.
#include "lb_host.cuh"
#include "lb_metrics.cuh"
#include "lb_scan.cuh"
#include "lb_tc.cuh"
#include "lb_tc1.cuh"
#include "lb_tc2.cuh"

using namespace lb;

namespace lb {

int shadow_kind_for(int metric) {
    return metric == LB_IP ? tc::SHADOW_IP : (metric == LB_COSINE ? tc::SHADOW_COSINE : tc::SHADOW_L2);
}
static int shadow_dp(const lb_index* idx, int kind) {
    int d = (int)idx->dim + (kind == tc::SHADOW_L2 ? 3 : 0);
    return (d + tc::KBLK - 1) / tc::KBLK * tc::KBLK;
}
bool tc_supported(const lb_index* idx, int metric) {
    if (idx->dtype != LB_F32) return false;
    if (metric != LB_IP && metric != LB_COSINE && metric != LB_L2) return false;
    return shadow_dp(idx, shadow_kind_for(metric)) <= tc::MAX_DP;
}


// corpus rows per accumulator tile: 128 for CTA pairs when the A operand leaves room for two 128-column accumulators
static int tc_rows_per_tile(const lb_index* idx, int kind, bool pair) {
    if (!pair || shadow_dp(idx, kind) > tc::PairCfg<128>::kMaxDp) return 64;
    return tc_env_int("LYNSE_B200_TC_BN", 128) == 64 ? 64 : 128;
}

static int encode_shadow_map(CUtensorMap* out, void* base, int nkb, uint64_t n_tiles, int box_halves, int box_kb) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return fail(LB_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    // (256 bf16 = 512 B, 8 rows of 512 B = one 4 KiB half block, 2 halves, tiles * K blocks)
    cuuint64_t gdim[4] = {256, 8, 2, (cuuint64_t)n_tiles * (cuuint64_t)nkb};
    cuuint64_t gstride[3] = {512, 4096, 8192};
    cuuint32_t box[4] = {256, 8, (cuuint32_t)box_halves, (cuuint32_t)box_kb};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(LB_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return LB_OK;
}

int ensure_shadow(lb_index* idx, int kind) {
    Shadow& sh = idx->shadow[kind];
    const int Dp = shadow_dp(idx, kind);
    const int nkb = Dp / tc::KBLK;
    if (!idx->max_norm.p) {
        LB_TRY(idx->max_norm.ensure(3 * sizeof(float)));
        LB_CUDA_TRY(cudaMemsetAsync(idx->max_norm.p, 0, 3 * sizeof(float), idx->stream));
    }
    const uint64_t need_tiles = (ceil_div(idx->n, tc::BN) + 1) & ~(uint64_t)1;  // even: the 128-row kernel reads tiles in pairs
    if (need_tiles > sh.cap_tiles) {
        // grow with head-room; the tiled image is rebuilt from the f32 rows (a derived structure, like the reference's
        // lazily built caches that are dropped on append, flat_mmap.rs:341)
        const uint64_t cap = std::max<uint64_t>(need_tiles, sh.cap_tiles + sh.cap_tiles / 2);
        const uint64_t reserve_tiles = ceil_div(idx->rows.cap / row_bytes(idx), tc::BN);
        const uint64_t want = std::max(cap, std::min<uint64_t>(reserve_tiles, need_tiles * 4));
        const size_t bytes = (size_t)want * nkb * 2 * tc::HALF_BLOCK_BYTES;
        sh.buf.release();
        LB_TRY(sh.buf.ensure(bytes));
        LB_CUDA_TRY(cudaMemsetAsync(sh.buf.p, 0, bytes, idx->stream));
        sh.cap_tiles = want;
        sh.rows = 0;
    }
    sh.Dp = Dp;
    if (sh.rows < idx->n) {
        uint64_t first = sh.rows, cnt = idx->n - first;
        const int warps = 8;
        tc::build_shadow_kernel<<<(unsigned)ceil_div(cnt, warps), warps * 32, 0, idx->stream>>>(
            idx->rows.as<float>(), first, cnt, (int)idx->dim, Dp, kind, sh.buf.as<unsigned char>(), idx->max_norm.as<float>() + kind);
        LB_CUDA_TRY(cudaGetLastError());
        sh.rows = idx->n;
    }
    if (sh.tmap_tiles != need_tiles || sh.tmap_ptr != sh.buf.p) {
        const int rem = nkb % tc::KPS;
        for (int kernel = 0; kernel < 2; ++kernel) {
            const int halves = kernel == 0 ? 2 : 1;
            LB_TRY(encode_shadow_map(&sh.tmap_full[kernel], sh.buf.p, nkb, need_tiles, halves, tc::KPS));
            LB_TRY(encode_shadow_map(&sh.tmap_rem[kernel], sh.buf.p, nkb, need_tiles, halves, rem ? rem : 1));
        }
        sh.tmap_tiles = need_tiles;
        sh.tmap_ptr = sh.buf.p;
    }
    return LB_OK;
}

// ---- tensor-core plan ----------------------------------------------------------------------------------------
int run_tc(lb_index* idx, int metric, const float* d_queries, int nq, int k, uint32_t* d_rows, float* d_dists,
           uint32_t* d_counts, float* dump, const uint64_t* d_allow) {
    const int kind = shadow_kind_for(metric);
    const int n_mtiles = (nq + tc::BM - 1) / tc::BM;
    // one query tile: one CTA per partition (lb_tc1.cuh); more: CTA pairs (tcgen05 cta_group::2, lb_tc2.cuh)
    const bool pair = n_mtiles >= 2;
    const int cluster = pair ? 2 : 1;
    const int BN = tc_rows_per_tile(idx, kind, pair);
    LB_TRY(ensure_shadow(idx, kind));
    LB_TRY(refresh_small_segments(idx));
    Shadow& sh = idx->shadow[kind];
    const int Dp = sh.Dp;
    const int n_mgroups = (n_mtiles + cluster - 1) / cluster;
    const int nq_pad = n_mgroups * cluster * tc::BM;
    LB_TRY(idx->w_qb.ensure((size_t)nq_pad * Dp * 2));
    LB_TRY(idx->w_qnorm.ensure((size_t)nq * 4));
    {
        const int warps = 8;
        tc::prepare_queries_kernel<<<(nq_pad + warps - 1) / warps, warps * 32, 0, idx->stream>>>(
            d_queries, nq, nq_pad, (int)idx->dim, Dp, kind, idx->w_qb.as<__nv_bfloat16>(), idx->w_qnorm.as<float>());
        LB_CUDA_TRY(cudaGetLastError());
    }
    const uint32_t tiles_total = (uint32_t)ceil_div(idx->n, BN);
    // Slots: groups of n_mgroups co-resident clusters (one per query group) that stream the same row partitions in
    // lockstep, so every shadow tile comes from HBM once and is served to the other query groups of the slot from L2.
    const uint64_t G = (uint64_t)(idx->sm_count / cluster);  // clusters resident at once (one CTA per SM)
    uint64_t n_slots = std::max<uint64_t>(1, G / (uint64_t)n_mgroups);
    n_slots = std::min<uint64_t>(n_slots, tiles_total);
    n_slots = std::min<uint64_t>(n_slots, 4096 / tc::KP);
    // The certification needs the largest partition floor T (the KP-th best coarse score of one partition) to sit
    // well below the k-th best score overall, so the union of the shortlists must reach far past rank k: aim at
    // P*KP >= 32*k candidates (measured on C3, k = 100: P = 36 leaves 1385 of 1024 queries uncertified, P = 72
    // five, P >= 144 none).  LYNSE_B200_TC_PARTS overrides.
    // Large k (no single partition floor is far enough past rank k): a pre-pass over 1/64 of the corpus seeds every
    // query's floor (seed_floor_kernel); the main pass then only needs enough partitions for the true top-k not to
    // crowd into one 16-entry list: P >= 0.75 k.
    const bool seeded = k > tc::KP - 4 && tiles_total >= (uint32_t)(64 * 8) * (uint32_t)n_slots && dump == nullptr &&
                        tc_env_int("LYNSE_B200_TC_SEED", 1) != 0;
    // two epilogue sets (eight epilogue warps per CTA, two shortlists per partition): 128-row tiles of narrow rows at
    // large k, where the epilogue's instruction issue rate bounds the pass (LYNSE_B200_TC_EPI=1 turns it off)
    const bool epi2 = pair && BN == 128 && Dp <= tc::PairCfg<128, 3>::kMaxDp && k > tc::KP - 4 && tc_env_int("LYNSE_B200_TC_EPI", 2) == 2 &&
                      tc_env_int("LYNSE_B200_TC_NACC", 3) == 3;
    const uint64_t L = epi2 ? 2 : 1;
    uint64_t parts_per_slot = 1;
    {
        uint64_t want = seeded ? ((uint64_t)3 * k + 3) / 4 : ((uint64_t)32 * k + tc::KP - 1) / tc::KP;
        want = ceil_div(want, L);  // a partition contributes L shortlists
        const int env_parts = tc_env_int("LYNSE_B200_TC_PARTS", 0);
        if (env_parts > 0) want = (uint64_t)env_parts;
        while (n_slots * parts_per_slot < want && n_slots * (parts_per_slot + 1) * L <= 4096 / tc::KP) ++parts_per_slot;
    }
    uint64_t P = std::min<uint64_t>(n_slots * parts_per_slot, tiles_total);
    const uint32_t tiles_per_part = (uint32_t)ceil_div(tiles_total, P);
    P = ceil_div(tiles_total, tiles_per_part);
    parts_per_slot = ceil_div(P, n_slots);
    const uint64_t P_buf = std::max<uint64_t>(P, n_slots) * L;
    LB_TRY(idx->w_cand_score.ensure((size_t)nq * P_buf * tc::KP * 4));
    LB_TRY(idx->w_cand_row.ensure((size_t)nq * P_buf * tc::KP * 4));
    LB_TRY(idx->w_cand_thr.ensure((size_t)nq * P_buf * 4));
    LB_TRY(idx->w_flags.ensure((size_t)nq * 8 + 16));
    uint32_t* flags = idx->w_flags.as<uint32_t>();  // [0]=kernel error, [1]=n_uncertified, [4..]=per-query flags, then gthr[nq]
    LB_CUDA_TRY(cudaMemsetAsync(flags, 0, 16, idx->stream));
    LB_CUDA_TRY(cudaMemsetAsync(flags + 4 + nq, 0, (size_t)nq * 4, idx->stream));

    tc::TcArgs a{};
    a.qb = idx->w_qb.as<__nv_bfloat16>();
    a.nq = nq;
    a.n_mtiles = n_mtiles;
    a.Dp = Dp;
    a.rem_kb = (Dp / tc::KBLK) % tc::KPS;
    a.n_rows = (uint32_t)idx->n;
    a.tiles_total = tiles_total;
    a.tiles_per_part = tiles_per_part;
    a.P = (int)P;
    a.lists_per_part = (int)L;
    a.allow_bits = d_allow;
    a.cand_score = idx->w_cand_score.as<float>();
    a.cand_row = idx->w_cand_row.as<uint32_t>();
    a.cand_thr = idx->w_cand_thr.as<float>();
    a.gthr = flags + 4 + nq;
    // Shared floors: a published floor must have enough rows above it to be far past rank k (the certification needs
    // the final floor well below the k-th best score).  For k <= KP - 4 the KP-th best score of one partition will do.
    // For larger k a valid floor is the minimum over a group of m = ceil(10 k / KP) partitions (>= 10 k rows above it);
    // that variant is implemented (LYNSE_B200_TC_GROUPS=1) but off: on C3 (k = 100) it only becomes available after
    // the first m partitions (30 % of the pass) and measured 8.9 ms against 7.6 ms without any sharing.
    const int m_req = k <= tc::KP - 4 ? 1 : (10 * k + tc::KP - 1) / tc::KP;
    a.share_floor = ((int)P >= m_req && (m_req == 1 || tc_env_int("LYNSE_B200_TC_GROUPS", 0) != 0)) ? 1 : 0;
    if (seeded) a.share_floor = 2;
    a.floor_group = m_req;
    a.gfloor = nullptr;
    if (a.share_floor == 1 && m_req > 1) {
        LB_TRY(idx->w_gfloor.ensure((size_t)nq * P * 4));
        fill_f32_kernel<<<(unsigned)std::min<uint64_t>(ceil_div((uint64_t)nq * P, 256), 1024), 256, 0, idx->stream>>>(
            idx->w_gfloor.as<float>(), (uint64_t)nq * P, -INFINITY);
        LB_CUDA_TRY(cudaGetLastError());
        a.gfloor = idx->w_gfloor.as<float>();
    }
    a.error_flag = flags;
    a.dump = dump;
    a.n_slots = (int)n_slots;
    a.parts_per_slot = (int)parts_per_slot;
    a.window = tc_env_int("LYNSE_B200_TC_WINDOW", 16);
    a.prefetch_tiles = tc_env_int("LYNSE_B200_TC_PREFETCH", 0);
    a.debug_mode = tc_env_int("LYNSE_B200_TC_DEBUG", 0);
    // optional warm-up sample (LYNSE_B200_TC_SAMPLE tiles, scanned by every CTA before its partitions; off by default:
    // with the compact slow path the open gate at the start of a partition no longer stalls the tensor pipe)
    a.sample_tiles = 0;
    if (a.share_floor && tiles_per_part >= 1024) a.sample_tiles = std::min<int>(tc_env_int("LYNSE_B200_TC_SAMPLE", 0), (int)tiles_total);
    a.prof = nullptr;
    const bool want_prof = getenv("LYNSE_B200_TC_PROF") != nullptr;
    if (want_prof) {
        LB_TRY(idx->w_prof.ensure((size_t)idx->sm_count * 8 * 8));
        LB_CUDA_TRY(cudaMemsetAsync(idx->w_prof.p, 0, (size_t)idx->sm_count * 8 * 8, idx->stream));
        a.prof = idx->w_prof.as<unsigned long long>();
    }
    a.progress = nullptr;
    if (n_mgroups > 1 && a.window > 0) {
        LB_TRY(idx->w_progress.ensure((size_t)n_slots * tc::PROGRESS_STRIDE * 4));
        LB_CUDA_TRY(cudaMemsetAsync(idx->w_progress.p, 0, (size_t)n_slots * tc::PROGRESS_STRIDE * 4, idx->stream));
        a.progress = idx->w_progress.as<uint32_t>();
    }
    const int grid = (int)n_slots * n_mgroups * cluster;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(tc::NUM_THREADS);
    cfg.dynamicSmemBytes = tc::SMEM_BYTES;
    cfg.stream = idx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    auto launch_coarse = [&](const tc::TcArgs& args) -> int {
        if (epi2) {
            cudaLaunchConfig_t cfg2 = cfg;
            cfg2.blockDim = dim3(64 + 128 * 2);
            LB_CUDA_TRY(ensure_dynamic_smem(tc::coarse_pair_kernel<128, 3, 2>, (int)tc::SMEM_BYTES));
            LB_CUDA_TRY(cudaLaunchKernelEx(&cfg2, tc::coarse_pair_kernel<128, 3, 2>, sh.tmap_full[0], sh.tmap_rem[0], args));
        } else if (pair && BN == 128 && Dp <= tc::PairCfg<128, 3>::kMaxDp && tc_env_int("LYNSE_B200_TC_NACC", 3) == 3) {
            // narrow rows leave TMEM room for a third accumulator tile (see PairCfg)
            LB_CUDA_TRY(ensure_dynamic_smem(tc::coarse_pair_kernel<128, 3>, (int)tc::SMEM_BYTES));
            LB_CUDA_TRY(cudaLaunchKernelEx(&cfg, tc::coarse_pair_kernel<128, 3>, sh.tmap_full[0], sh.tmap_rem[0], args));
        } else if (pair && BN == 128) {
            LB_CUDA_TRY(ensure_dynamic_smem(tc::coarse_pair_kernel<128, 2>, (int)tc::SMEM_BYTES));
            LB_CUDA_TRY(cudaLaunchKernelEx(&cfg, tc::coarse_pair_kernel<128, 2>, sh.tmap_full[0], sh.tmap_rem[0], args));
        } else if (pair) {
            LB_CUDA_TRY(ensure_dynamic_smem(tc::coarse_pair_kernel<64, 2>, (int)tc::SMEM_BYTES));
            LB_CUDA_TRY(cudaLaunchKernelEx(&cfg, tc::coarse_pair_kernel<64, 2>, sh.tmap_full[1], sh.tmap_rem[1], args));
        } else {
            LB_CUDA_TRY(ensure_dynamic_smem(tc::coarse_single_kernel, (int)tc::SMEM_BYTES));
            LB_CUDA_TRY(cudaLaunchKernelEx(&cfg, tc::coarse_single_kernel, sh.tmap_full[0], sh.tmap_rem[0], args));
        }
        return LB_OK;
    };
    if (idx->timing) cudaEventRecord(idx->ev[0], idx->stream);
    if (seeded) {
        // pre-pass: the first 1/64 of the tiles, one partition per slot, private floors; then the seed
        tc::TcArgs sa = a;
        const uint32_t S = std::max<uint32_t>(tiles_total / 64, (uint32_t)n_slots * 8);
        sa.tiles_total = S;
        sa.n_rows = (uint32_t)std::min<uint64_t>(idx->n, (uint64_t)S * BN);
        sa.tiles_per_part = (uint32_t)ceil_div(S, n_slots);
        sa.P = (int)ceil_div(S, sa.tiles_per_part);
        sa.parts_per_slot = 1;
        sa.share_floor = 0;
        sa.gfloor = nullptr;
        sa.sample_tiles = 0;
        LB_TRY(launch_coarse(sa));
        const int s_lists = sa.P * (int)L;
        const int sm = next_pow2(s_lists * tc::KP);
        // aim at ~10 k rows of the whole corpus above the seeded floor
        int r = (int)ceil_div((uint64_t)10 * k * S, tiles_total);
        r = std::max(4, std::min(r, s_lists * tc::KP / 2));
        LB_CUDA_TRY(ensure_dynamic_smem(tc::seed_floor_kernel, sm * 8));
        tc::seed_floor_kernel<<<nq, 256, (size_t)sm * 8, idx->stream>>>(sa.cand_score, sa.cand_row, s_lists, sm, r, a.gthr);
        LB_CUDA_TRY(cudaGetLastError());
        if (a.progress) LB_CUDA_TRY(cudaMemsetAsync(idx->w_progress.p, 0, (size_t)n_slots * tc::PROGRESS_STRIDE * 4, idx->stream));
        idx->stats.kernels_launched += 2;
    }
    LB_TRY(launch_coarse(a));
    LB_CUDA_TRY(cudaGetLastError());
    if (idx->timing) cudaEventRecord(idx->ev[1], idx->stream);

    tc::FinArgs f{};
    f.cand_score = a.cand_score;
    f.cand_row = a.cand_row;
    f.cand_thr = a.cand_thr;
    f.P = (int)(P * L);
    f.M1 = next_pow2((int)(P * L) * tc::KP);
    f.R = std::min(1024, std::max(128, next_pow2(4 * k)));
    f.corpus = idx->rows.as<float>();
    f.dim = (int)idx->dim;
    f.queries = d_queries;
    f.qnorm = idx->w_qnorm.as<float>();
    f.max_norm = idx->max_norm.as<float>() + kind;
    f.nq = nq;
    f.k = k;
    f.metric = metric;
    f.eps_rel = (0.00390625f * 1.01f + (float)Dp * 4.76837158e-7f) * 1.0001f;
    f.small_seg = idx->small_seg.as<uint32_t>();
    f.n_small = idx->n_small;
    f.out_rows = d_rows;
    f.out_dists = d_dists;
    f.out_counts = d_counts;
    f.uncertified = flags + 4;
    f.n_uncertified = flags + 1;
    const int fin_threads = nq <= 64 ? 1024 : 256;  // few queries: few blocks, so each gets 1024 threads
    size_t fsmem = (size_t)(f.M1 + f.R) * 8 + (size_t)((idx->dim + 3) & ~3u) * 4 + (size_t)(fin_threads / 8) * tc::FIN_COLS * 4;  // + row buffers
    if (metric_ascending(metric)) {
        LB_CUDA_TRY(ensure_dynamic_smem(tc::finalize_kernel<true>, (int)fsmem));
        tc::finalize_kernel<true><<<nq, fin_threads, fsmem, idx->stream>>>(f);
    } else {
        LB_CUDA_TRY(ensure_dynamic_smem(tc::finalize_kernel<false>, (int)fsmem));
        tc::finalize_kernel<false><<<nq, fin_threads, fsmem, idx->stream>>>(f);
    }
    LB_CUDA_TRY(cudaGetLastError());
    idx->stats.kernels_launched += 3;
    idx->stats.n_partitions = (uint32_t)P;
    idx->stats.plan_used = 1;
    idx->stats.algorithmic_bytes = (uint64_t)idx->n * Dp * 2;
    idx->stats.algorithmic_flops = 2ull * (uint64_t)nq * idx->n * idx->dim;

    uint32_t head[4] = {0, 0, 0, 0};
    LB_CUDA_TRY(cudaMemcpyAsync(head, flags, 16, cudaMemcpyDeviceToHost, idx->stream));
    LB_CUDA_TRY(cudaStreamSynchronize(idx->stream));
    if (idx->timing) {
        float ms = 0;
        cudaEventElapsedTime(&ms, idx->ev[0], idx->ev[1]);
        idx->stats.ms_dominant = ms;
    }
    if (want_prof) {
        std::vector<unsigned long long> pr((size_t)idx->sm_count * 8);
        LB_CUDA_TRY(cudaMemcpy(pr.data(), idx->w_prof.p, pr.size() * 8, cudaMemcpyDeviceToHost));
        double sum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        int n_lead = 0, n_cta = 0;
        if (pair) {
            double sc[5] = {0, 0, 0, 0, 0};
            double mx = 0;
            for (int b = 1; b < grid && b < idx->sm_count; b += 2) {
                for (int i = 0; i < 5; ++i) sc[i] += (double)pr[(size_t)b * 8 + i];
                mx = std::max(mx, (double)pr[(size_t)b * 8 + 3]);
                for (int i = 0; i < 6; ++i) pr[(size_t)b * 8 + i] = 0;
            }
            if (sc[4] > 0)
                fprintf(stderr, "[lynse_b200] scan (one warp per odd CTA): %.0f cycles/tile, slow tiles %.3f/tile at %.0f cycles each, longest %.0f\n",
                        sc[0] / sc[4], sc[2] / sc[4], sc[2] > 0 ? sc[1] / sc[2] : 0.0, mx);
        }
        for (int b = 0; b < grid && b < idx->sm_count; ++b) {
            if (pr[(size_t)b * 8] > 0) {
                ++n_lead;
                for (int i = 0; i < 6; ++i) sum[i] += (double)pr[(size_t)b * 8 + i];
            }
            if (pr[(size_t)b * 8 + 6] + pr[(size_t)b * 8 + 7] > 0) {
                ++n_cta;
                sum[6] += (double)pr[(size_t)b * 8 + 6];
                sum[7] += (double)pr[(size_t)b * 8 + 7];
            }
        }
        if (n_lead > 0 && n_cta > 0 && sum[5] > 0)
            fprintf(stderr,
                    "[lynse_b200] per tile (cycles): mma loop %.0f, wait tempty %.0f (%.2f waits/tile), wait full %.0f (%.2f waits/tile); "
                    "epilogue wait tfull %.0f, read+release %.0f\n",
                    sum[0] / sum[5], sum[1] / sum[5], sum[3] / sum[5], sum[2] / sum[5], sum[4] / sum[5],
                    sum[6] / n_cta / (sum[5] / n_lead), sum[7] / n_cta / (sum[5] / n_lead));
    }
    if (getenv("LYNSE_B200_TC_TRACE") && head[3] > 0)
        fprintf(stderr, "[lynse_b200] coarse kernel: %.3f ms, %.0f SM MHz, grid %d, cluster %d, slots %d, P %d\n", head[3] * 1e-6,
                (double)head[2] * 16.0 / (double)head[3] * 1e3, grid, cluster, (int)n_slots, (int)P);
    if (head[0] != 0)
        return fail(LB_INTERNAL, "tensor-core coarse kernel: barrier wait timed out (code " + std::to_string(head[0]) + ")");
    idx->stats.n_fallback = head[1];
    if (head[1] > 0) {
        // Re-run the uncertified queries with the exact scan and overwrite their result slots.
        std::vector<uint32_t> fl(nq);
        LB_CUDA_TRY(cudaMemcpy(fl.data(), flags + 4, (size_t)nq * 4, cudaMemcpyDeviceToHost));
        std::vector<uint32_t> qmap;
        for (int q = 0; q < nq; ++q)
            if (fl[q]) qmap.push_back((uint32_t)q);
        const int ns = (int)qmap.size();
        LB_TRY(idx->w_sub_q.ensure((size_t)ns * idx->dim * 4));
        LB_TRY(idx->w_qmap.ensure((size_t)ns * 4));
        LB_CUDA_TRY(cudaMemcpyAsync(idx->w_qmap.p, qmap.data(), (size_t)ns * 4, cudaMemcpyHostToDevice, idx->stream));
        for (int i = 0; i < ns; ++i)
            LB_CUDA_TRY(cudaMemcpyAsync(idx->w_sub_q.as<float>() + (size_t)i * idx->dim, d_queries + (size_t)qmap[i] * idx->dim,
                                        (size_t)idx->dim * 4, cudaMemcpyDeviceToDevice, idx->stream));
        ScanRequest r;
        r.corpus = idx->rows.as<float>();
        r.n_rows = idx->n;
        r.dim = (int)idx->dim;
        r.queries = idx->w_sub_q.as<float>();
        r.nq = ns;
        r.k = k;
        r.metric = metric;
        r.small_seg = idx->small_seg.as<uint32_t>();
        r.n_small = idx->n_small;
        r.allow_bits = d_allow;
        r.qmap = idx->w_qmap.as<uint32_t>();
        r.out_rows = d_rows;
        r.out_dists = d_dists;
        r.out_counts = d_counts;
        int kern = 0;
        LB_TRY(run_scan(idx, r, &kern, nullptr));
        idx->stats.kernels_launched += kern;
        LB_CUDA_TRY(cudaStreamSynchronize(idx->stream));
    }
    return LB_OK;
}

}  // namespace lb

extern "C" {

// ---- diagnostics ------------------------------------------------------------------------------------------------------------------
int lb_debug_tc_scores(const float* queries, uint32_t nq, const float* rows, uint32_t n, uint32_t dim, float* out) {
    if (!queries || !rows || !out || nq == 0 || n == 0) return fail(LB_INVALID_ARGUMENT, "bad arguments");
    int device = 0;
    LB_CUDA_TRY(cudaGetDevice(&device));
    lb_index* idx = nullptr;
    LB_TRY(lb_index_create(&idx, dim, LB_F32, device));
    int st = lb_index_append_f32(idx, rows, n);
    float* dump = nullptr;
    if (st == LB_OK && !tc_supported(idx, LB_IP)) st = fail(LB_UNSUPPORTED, "dimension too large for the tensor-core path");
    if (st == LB_OK) {
        std::lock_guard<std::mutex> lock(idx->mu);
        DeviceGuard g(idx->device);
        int n_mtiles = ((int)nq + tc::BM - 1) / tc::BM;
        n_mtiles = (n_mtiles + 1) & ~1;  // room for the padded query tile of a 2-CTA cluster
        const int dbg_bn = tc_rows_per_tile(idx, tc::SHADOW_IP, (int)nq > tc::BM);
        const size_t ld = (size_t)ceil_div(n, dbg_bn) * dbg_bn;
        const size_t dump_elems = (size_t)n_mtiles * tc::BM * ld;
        const int k = (int)std::min<uint32_t>(n, 10);
        cudaError_t e = cudaMalloc(&dump, dump_elems * 4);
        if (e == cudaSuccess) e = cudaMemset(dump, 0xFF, dump_elems * 4);
        if (e != cudaSuccess) st = fail(LB_CUDA, cudaGetErrorString(e));
        if (st == LB_OK) st = idx->w_queries.ensure((size_t)nq * dim * 4);
        if (st == LB_OK) st = idx->w_out_rows.ensure((size_t)nq * k * 4);
        if (st == LB_OK) st = idx->w_out_dists.ensure((size_t)nq * k * 4);
        if (st == LB_OK) st = idx->w_out_counts.ensure((size_t)nq * 4);
        if (st == LB_OK) {
            e = cudaMemcpyAsync(idx->w_queries.p, queries, (size_t)nq * dim * 4, cudaMemcpyHostToDevice, idx->stream);
            if (e != cudaSuccess) st = fail(LB_CUDA, cudaGetErrorString(e));
        }
        if (st == LB_OK)
            st = run_tc(idx, LB_IP, idx->w_queries.as<float>(), (int)nq, k, idx->w_out_rows.as<uint32_t>(),
                        idx->w_out_dists.as<float>(), idx->w_out_counts.as<uint32_t>(), dump);
        if (st == LB_OK) {
            e = cudaMemcpy2D(out, (size_t)n * 4, dump, ld * 4, (size_t)n * 4, nq, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) st = fail(LB_CUDA, cudaGetErrorString(e));
        }
    }
    if (dump) cudaFree(dump);
    std::string keep = lb_last_error();
    lb_index_destroy(idx);
    set_error(keep);
    return st;
}

int lb_debug_mma_rate(int n, int n_acc, int iters, int a_in_tmem, int grid, uint64_t* cycles_total, uint64_t* cycles_issue) {
    if (iters < 16 || grid < 1) return fail(LB_INVALID_ARGUMENT, "bad probe arguments");
    unsigned long long* d = nullptr;
    LB_CUDA_TRY(cudaMalloc(&d, (size_t)grid * 16));
    const size_t smem = 49152 + 64 + 1024;
    cudaError_t e = cudaSuccess;
    bool found = false;
#define LB_PROBE(NN, NA, TSV)                                                                                         \
    if (!found && n == NN && n_acc == NA && (a_in_tmem != 0) == TSV) {                                              \
        found = true;                                                                                                \
        e = cudaFuncSetAttribute(tc::mma_rate_kernel<NN, NA, TSV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e == cudaSuccess) {                                                                                      \
            tc::mma_rate_kernel<NN, NA, TSV><<<grid, 64, smem>>>(iters / 16, tc_env_int("LYNSE_B200_PROBE_COMMIT", 0), d);                                     \
            e = cudaDeviceSynchronize();                                                                             \
        }                                                                                                            \
    }
    LB_PROBE(64, 1, true)
    LB_PROBE(64, 2, true)
    LB_PROBE(128, 1, true)
    LB_PROBE(64, 1, false)
    LB_PROBE(64, 2, false)
    LB_PROBE(64, 4, false)
    LB_PROBE(128, 1, false)
    LB_PROBE(128, 2, false)
    LB_PROBE(256, 1, false)
    LB_PROBE(256, 2, false)
#undef LB_PROBE
    if (!found) {
        cudaFree(d);
        return fail(LB_INVALID_ARGUMENT, "probe shape not instantiated");
    }
    std::vector<unsigned long long> h((size_t)grid * 2);
    if (e == cudaSuccess) e = cudaMemcpy(h.data(), d, h.size() * 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(LB_CUDA, std::string("mma probe: ") + cudaGetErrorString(e));
    unsigned long long mt = 0, mi = 0;
    for (int i = 0; i < grid; ++i) {
        mt = std::max(mt, h[2 * i]);
        mi = std::max(mi, h[2 * i + 1]);
    }
    *cycles_total = mt;
    *cycles_issue = mi;
    return LB_OK;
}

}  // extern "C"
