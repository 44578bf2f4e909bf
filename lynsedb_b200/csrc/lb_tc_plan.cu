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

constexpr int TC_MAX_ROW_BYTES = 2 * tc::MAX_DP;   // A operand: row_bytes / 4 <= 384 TMEM columns
constexpr uint32_t TC_HIT_TOTAL = 8192;            // hit mode: entries of a query's hit regions together (each region 64..1024)
constexpr int TC_FIN_MAX = 4096;                   // candidates finalize can sort per query

int shadow_kind_for(int metric) {
    return metric == LB_IP ? tc::SHADOW_IP : (metric == LB_COSINE ? tc::SHADOW_COSINE : tc::SHADOW_L2);
}
// operand row length in bytes (a multiple of the 128-byte swizzle row)
static int operand_row_bytes(int dim, int operand) {
    const int b = operand == tc::OPERAND_BF16 ? dim * 2 : dim;
    return (b + 127) / 128 * 128;
}
// L2 shadow: |c|^2 travels as three extra bf16 columns when the padded row has (or can get) room for them, so the
// contraction itself yields 2 q.c - |c|^2 and the epilogue ranks raw accumulators; rows of 766..768 dims (and
// LYNSE_B200_TC_L2_BIAS=1) keep the norm as an f32 side value that the epilogue subtracts (CM_F32_BIAS).
static bool l2_uses_bias(int dim) {
    return operand_row_bytes(dim + 3, tc::OPERAND_BF16) > TC_MAX_ROW_BYTES || tc_env_int("LYNSE_B200_TC_L2_BIAS", 0) != 0;
}
static int shadow_row_bytes(const lb_index* idx, int kind, int operand) {
    const int extra = (kind == tc::SHADOW_L2 && operand == tc::OPERAND_BF16 && !l2_uses_bias((int)idx->dim)) ? 3 : 0;
    return operand_row_bytes((int)idx->dim + extra, operand);
}
// LYNSE_B200_TC_OPERAND = bf16 | u8 | auto (default).  auto: 8-bit operands for IP / cosine when the measured
// quantisation error of the corpus is within 2x of what bf16 rounding would have cost (uniform-ish data: 1.3x; heavy
// tails: 5x and more, which would only send queries to the exact-scan fallback); L2 keeps bf16, whose epilogue budget
// (one subtraction per score) is what bounds narrow rows.
static int operand_policy() {
    const char* env = getenv("LYNSE_B200_TC_OPERAND");
    if (env && (!strcmp(env, "bf16") || !strcmp(env, "0"))) return 0;
    if (env && (!strcmp(env, "u8") || !strcmp(env, "i8") || !strcmp(env, "1"))) return 1;
    return 2;
}
static int first_operand(const lb_index* idx, int kind) {
    const int pol = operand_policy();
    if (kind == tc::SHADOW_L2) return tc::OPERAND_BF16;
    if (pol == 0 && shadow_row_bytes(idx, kind, tc::OPERAND_BF16) <= TC_MAX_ROW_BYTES) return tc::OPERAND_BF16;
    return tc::OPERAND_U8;
}
bool tc_supported(lb_index* idx, int metric) {
    if (idx->dtype != LB_F32 && idx->dtype != LB_F16) return false;
    if (metric != LB_IP && metric != LB_COSINE && metric != LB_L2) return false;
    const int kind = shadow_kind_for(metric);
    const Shadow& sh = idx->shadow[kind];
    if (sh.disabled) return false;
    const int operand = sh.operand >= 0 ? sh.operand : first_operand(idx, kind);
    return shadow_row_bytes(idx, kind, operand) <= TC_MAX_ROW_BYTES;
}

static int encode_shadow_map(CUtensorMap* out, void* base, int nkb, uint64_t n_tiles, int box_halves, int box_kb) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return fail(LB_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    // (512 B = 4 rows of one K block, 8 of those = one 4 KiB half block, 2 halves, tiles * K blocks); plain bytes
    cuuint64_t gdim[4] = {256, 8, 2, (cuuint64_t)n_tiles * (cuuint64_t)nkb};
    cuuint64_t gstride[3] = {512, 4096, 8192};
    cuuint32_t box[4] = {256, 8, (cuuint32_t)box_halves, (cuuint32_t)box_kb};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(LB_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return LB_OK;
}

static int reset_stats(lb_index* idx, Shadow& sh) {
    LB_TRY(sh.stats.ensure(sizeof(tc::ShadowStats)));
    tc::ShadowStats z{};
    z.vmin_ord = 0xFFFFFFFFu;
    LB_CUDA_TRY(cudaMemcpyAsync(sh.stats.p, &z, sizeof(z), cudaMemcpyHostToDevice, idx->stream));
    LB_CUDA_TRY(cudaStreamSynchronize(idx->stream));  // z is a stack object
    return LB_OK;
}
static int read_stats(lb_index* idx, Shadow& sh, tc::ShadowStats* out) {
    LB_CUDA_TRY(cudaMemcpyAsync(out, sh.stats.p, sizeof(*out), cudaMemcpyDeviceToHost, idx->stream));
    LB_CUDA_TRY(cudaStreamSynchronize(idx->stream));
    return LB_OK;
}
static float bits_f32(uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}

// (Re)allocates the tiled image for `n` rows of `row_bytes` and, when asked, the side array(s) filled with pad values.
static int shadow_reserve(lb_index* idx, Shadow& sh, uint64_t n, uint64_t reserve_rows, int row_bytes, int n_side, uint32_t pad0, uint32_t pad1) {
    const int nkb = row_bytes / 128;
    const uint64_t need_tiles = (ceil_div(n, tc::BN) + 1) & ~(uint64_t)1;  // even: the 128-row kernels read tiles in pairs
    if (need_tiles > sh.cap_tiles) {
        // grow with head-room; the tiled image is rebuilt from the rows (a derived structure, like the reference's
        // lazily built caches that are dropped on append, flat_mmap.rs:341)
        const uint64_t cap = std::max<uint64_t>(need_tiles, sh.cap_tiles + sh.cap_tiles / 2);
        const uint64_t reserve_tiles = ceil_div(reserve_rows, tc::BN);
        const uint64_t want = std::max(cap, std::min<uint64_t>(reserve_tiles, need_tiles * 4));
        const size_t bytes = (size_t)want * nkb * 2 * tc::HALF_BLOCK_BYTES;
        sh.buf.release();
        sh.side.release();
        sh.side2.release();
        LB_TRY(sh.buf.ensure(bytes));
        LB_CUDA_TRY(cudaMemsetAsync(sh.buf.p, 0, bytes, idx->stream));
        const uint64_t side_words = (want + 2) * tc::BN;  // the epilogue prefetches one tile past the last
        if (n_side >= 1) {
            LB_TRY(sh.side.ensure(side_words * 4));
            tc::fill_u32_kernel<<<1024, 256, 0, idx->stream>>>(sh.side.as<uint32_t>(), side_words, pad0);
        }
        if (n_side >= 2) {
            LB_TRY(sh.side2.ensure(side_words * 4));
            tc::fill_u32_kernel<<<1024, 256, 0, idx->stream>>>(sh.side2.as<uint32_t>(), side_words, pad1);
        }
        LB_CUDA_TRY(cudaGetLastError());
        sh.cap_tiles = want;
        sh.rows = 0;
    }
    sh.Dp = row_bytes / 2;
    return LB_OK;
}
static int shadow_maps(Shadow& sh, uint64_t n) {
    const int nkb = sh.Dp / tc::KBLK;
    const uint64_t need_tiles = (ceil_div(n, tc::BN) + 1) & ~(uint64_t)1;
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

int ensure_shadow(lb_index* idx, int kind) {
    Shadow& sh = idx->shadow[kind];
    if (sh.disabled) return LB_OK;
    if (sh.operand < 0) sh.operand = first_operand(idx, kind);
    const int warps = 8;
    for (int attempt = 0; attempt < 2; ++attempt) {
        const int rb = shadow_row_bytes(idx, kind, sh.operand);
        if (rb > TC_MAX_ROW_BYTES) return fail(LB_UNSUPPORTED, "dimension too large for the tensor-core plan");
        if (!sh.stats.p) LB_TRY(reset_stats(idx, sh));
        if (sh.rows == 0) sh.l2_bias = kind == tc::SHADOW_L2 && l2_uses_bias((int)idx->dim);
        LB_TRY(shadow_reserve(idx, sh, idx->n, idx->rows.cap / row_bytes(idx), rb, sh.l2_bias ? 1 : 0, 0x7F800000u /* +inf */, 0u));
        if (sh.rows < idx->n) {
            uint64_t first = sh.rows, cnt = idx->n - first;
            tc::ShadowStats st{};
            const bool f16 = idx->dtype == LB_F16;   // binary16 rows: the same kernels with RT = __half
            const unsigned grid_rows = (unsigned)ceil_div(cnt, warps);
            if (sh.operand == tc::OPERAND_U8) {
                // pass 1: element range of the new rows; a wider range than the image was quantised for rebuilds it
                if (f16)
                    tc::shadow_range_kernel<<<grid_rows, warps * 32, 0, idx->stream>>>(idx->rows.as<__half>(), first, cnt, (int)idx->dim, kind,
                                                                                      sh.stats.as<tc::ShadowStats>());
                else
                    tc::shadow_range_kernel<<<grid_rows, warps * 32, 0, idx->stream>>>(idx->rows.as<float>(), first, cnt, (int)idx->dim, kind,
                                                                                      sh.stats.as<tc::ShadowStats>());
                LB_CUDA_TRY(cudaGetLastError());
                LB_TRY(read_stats(idx, sh, &st));
                if (st.nonfinite) {
                    sh.disabled = true;
                    return LB_OK;
                }
                const float lo = f32_from_orderable(st.vmin_ord), hi = f32_from_orderable(st.vmax_ord);
                if (first == 0 || lo < sh.range_lo || hi > sh.range_hi) {
                    sh.range_lo = lo;
                    sh.range_hi = hi;
                    sh.c_zero = lo;
                    sh.c_scale = hi > lo ? (hi - lo) / 255.0f : 1.0f;
                    first = 0;
                    cnt = idx->n;
                    // the error statistics restart with the new quantisation
                    LB_CUDA_TRY(cudaMemsetAsync(&sh.stats.as<tc::ShadowStats>()->emax_bits, 0, 4, idx->stream));
                }
                const unsigned grid_all = (unsigned)ceil_div(cnt, warps);
                if (f16)
                    tc::build_shadow_kernel<tc::OPERAND_U8><<<grid_all, warps * 32, 0, idx->stream>>>(
                        idx->rows.as<__half>(), first, cnt, (int)idx->dim, rb, kind, sh.buf.as<unsigned char>(), nullptr,
                        sh.stats.as<tc::ShadowStats>(), sh.c_scale, sh.c_zero);
                else
                    tc::build_shadow_kernel<tc::OPERAND_U8><<<grid_all, warps * 32, 0, idx->stream>>>(
                        idx->rows.as<float>(), first, cnt, (int)idx->dim, rb, kind, sh.buf.as<unsigned char>(), nullptr,
                        sh.stats.as<tc::ShadowStats>(), sh.c_scale, sh.c_zero);
            } else if (f16) {
                tc::build_shadow_kernel<tc::OPERAND_BF16><<<grid_rows, warps * 32, 0, idx->stream>>>(
                    idx->rows.as<__half>(), first, cnt, (int)idx->dim, rb, kind, sh.buf.as<unsigned char>(), sh.l2_bias ? sh.side.as<float>() : nullptr,
                    sh.stats.as<tc::ShadowStats>(), 1.0f, 0.0f);
            } else {
                tc::build_shadow_kernel<tc::OPERAND_BF16><<<grid_rows, warps * 32, 0, idx->stream>>>(
                    idx->rows.as<float>(), first, cnt, (int)idx->dim, rb, kind, sh.buf.as<unsigned char>(), sh.l2_bias ? sh.side.as<float>() : nullptr,
                    sh.stats.as<tc::ShadowStats>(), 1.0f, 0.0f);
            }
            LB_CUDA_TRY(cudaGetLastError());
            LB_TRY(read_stats(idx, sh, &st));
            sh.rows = idx->n;
            sh.cmax = bits_f32(st.cmax_bits);
            sh.emax = bits_f32(st.emax_bits);
            if (st.nonfinite) {
                sh.disabled = true;
                return LB_OK;
            }
            if (sh.operand == tc::OPERAND_U8 && operand_policy() == 2 && shadow_row_bytes(idx, kind, tc::OPERAND_BF16) <= TC_MAX_ROW_BYTES &&
                sh.emax > 2.0f * bits_f32(st.e16max_bits)) {
                // 8-bit quantisation is too coarse for this corpus (heavy tails): use the bf16 operand instead
                if (getenv("LYNSE_B200_TC_TRACE"))
                    fprintf(stderr, "[lynse_b200] shadow %d: u8 error %.4g against bf16 %.4g -> bf16 operand\n", kind, sh.emax, bits_f32(st.e16max_bits));
                sh.release();
                sh.operand = tc::OPERAND_BF16;
                continue;
            }
            if (getenv("LYNSE_B200_TC_TRACE"))
                fprintf(stderr, "[lynse_b200] shadow %d: operand %s, %llu rows x %d B, max |c'| %.6g, max |c' - c~| %.6g (bf16 would be %.6g)\n", kind,
                        sh.operand == tc::OPERAND_U8 ? "u8" : "bf16", (unsigned long long)idx->n, rb, sh.cmax, sh.emax, bits_f32(st.e16max_bits));
        }
        break;
    }
    return shadow_maps(sh, idx->n);
}

// ---- the packed rows as {0,1} bytes ---------------------------------------------------------------------------
static int bits_row_bytes(int n_words) { return (n_words * 64 + 127) / 128 * 128; }
// Worth it from a few tens of queries on (the image is 8x the packed rows; below that the popcount scan is HBM-bound on
// 1/8 of the bytes), when the image fits in free memory.  LYNSE_B200_BITS_TC=0 turns the plan off.
bool tc_bits_supported(lb_index* idx, int metric, int n_words, int nq, int k) {
    if (!metric_binary(metric) || k > 256 || idx->n < 4096) return false;
    if (tc_env_int("LYNSE_B200_BITS_TC", 1) == 0) return false;
    if (bits_row_bytes(n_words) > TC_MAX_ROW_BYTES) return false;
    if (nq < tc_env_int("LYNSE_B200_BITS_TC_MIN_Q", 32)) return false;
    if (idx->bits_shadow.disabled) return false;
    if (idx->bits_shadow.rows == idx->n && idx->bits_shadow.buf.p) return true;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return false;
    const size_t need = (size_t)(ceil_div(idx->n, tc::BN) + 2) * tc::BN * (size_t)bits_row_bytes(n_words) + (size_t)idx->n * 8 + idx->bits_shadow.buf.cap;
    return need + (2ull << 30) < free_b + idx->bits_shadow.buf.cap;
}
static int ensure_bits_shadow(lb_index* idx, const uint64_t* words, int n_words) {
    Shadow& sh = idx->bits_shadow;
    const int rb = bits_row_bytes(n_words);
    sh.operand = tc::OPERAND_U8;
    float big = 1e30f;
    uint32_t big_bits;
    memcpy(&big_bits, &big, 4);
    LB_TRY(shadow_reserve(idx, sh, idx->n, idx->rows.cap / row_bytes(idx), rb, 2, 0x3FFFFFFFu, big_bits));
    if (sh.rows < idx->n) {
        const uint64_t first = sh.rows, cnt = idx->n - first;
        const int warps = 8;
        tc::build_bits_shadow_kernel<<<(unsigned)ceil_div(cnt, warps), warps * 32, 0, idx->stream>>>(
            words, first, cnt, n_words, rb, sh.buf.as<unsigned char>(), sh.side.as<uint32_t>(), sh.side2.as<float>());
        LB_CUDA_TRY(cudaGetLastError());
        sh.rows = idx->n;
    }
    return shadow_maps(sh, idx->n);
}

// ---- coarse pass ------------------------------------------------------------------------------------------------
struct CoarseJob {
    Shadow* sh = nullptr;
    int mode = tc::CM_F32;
    const unsigned char* qb = nullptr;      // prepared A operand rows
    const float* qaux = nullptr;
    const uint32_t* bias = nullptr;
    const uint32_t* idesc_extra = nullptr;
    int nq = 0, k = 0;
    int n_ksteps = 0;      // MMA K steps (32 operand bytes) that hold data
    const uint64_t* d_allow = nullptr;
    float* dump = nullptr;
    // results of the planning, for finalize
    int n_lists = 0;       // shortlists per query (list mode)
    bool hit_mode = false;
    uint32_t hit_cap = 0;
    uint32_t* flags = nullptr;   // [0]=kernel error, [1]=n_uncertified, [2..3] clock probe, [4..4+nq) per-query flags, then gthr[nq]
    uint32_t* gthr = nullptr;
};
static bool mode_int_key(int mode) { return mode == tc::CM_I32 || mode == tc::CM_I32_HAMMING; }

// kernel shapes: 0 = one CTA per query tile; CTA pairs: 1 = <64 rows, 2 accumulators>, 2 = <128, 2>, 3 = <128, 3>,
// 4 = <128, 3, two epilogue sets> (hit mode only: without the shortlist registers the eight epilogue warps fit)
template <int MODE, bool HITS>
static int launch_coarse_mode(int cfg_id, cudaLaunchConfig_t cfg, const Shadow& sh, const tc::TcArgs& args) {
    switch (cfg_id) {
        case 0:
            LB_CUDA_TRY(ensure_dynamic_smem(tc::coarse_single_kernel<MODE, HITS>, (int)tc::SMEM_BYTES));
            LB_CUDA_TRY(cudaLaunchKernelEx(&cfg, tc::coarse_single_kernel<MODE, HITS>, sh.tmap_full[0], sh.tmap_rem[0], args));
            break;
        case 1:
            LB_CUDA_TRY(ensure_dynamic_smem(tc::coarse_pair_kernel<64, 2, 1, MODE, HITS>, (int)tc::SMEM_BYTES));
            LB_CUDA_TRY(cudaLaunchKernelEx(&cfg, tc::coarse_pair_kernel<64, 2, 1, MODE, HITS>, sh.tmap_full[1], sh.tmap_rem[1], args));
            break;
        case 2:
            LB_CUDA_TRY(ensure_dynamic_smem(tc::coarse_pair_kernel<128, 2, 1, MODE, HITS>, (int)tc::SMEM_BYTES));
            LB_CUDA_TRY(cudaLaunchKernelEx(&cfg, tc::coarse_pair_kernel<128, 2, 1, MODE, HITS>, sh.tmap_full[0], sh.tmap_rem[0], args));
            break;
        case 3:
            LB_CUDA_TRY(ensure_dynamic_smem(tc::coarse_pair_kernel<128, 3, 1, MODE, HITS>, (int)tc::SMEM_BYTES));
            LB_CUDA_TRY(cudaLaunchKernelEx(&cfg, tc::coarse_pair_kernel<128, 3, 1, MODE, HITS>, sh.tmap_full[0], sh.tmap_rem[0], args));
            break;
        case 6:
        case 7:
            // list mode with helper warps (lb_tc2.cuh, HELP_): 6 = two accumulator tiles, 7 = three
            if constexpr (!HITS && (MODE == tc::CM_I32 || MODE == tc::CM_F32)) {
                cfg.blockDim = dim3(64 + 128 * 2);
                cfg.dynamicSmemBytes = tc::SMEM_BYTES_HELP;
                if (cfg_id == 6) {
                    LB_CUDA_TRY(ensure_dynamic_smem(tc::coarse_pair_kernel<128, 2, 1, MODE, false, true>, (int)tc::SMEM_BYTES_HELP));
                    LB_CUDA_TRY(cudaLaunchKernelEx(&cfg, tc::coarse_pair_kernel<128, 2, 1, MODE, false, true>, sh.tmap_full[0], sh.tmap_rem[0], args));
                } else {
                    LB_CUDA_TRY(ensure_dynamic_smem(tc::coarse_pair_kernel<128, 3, 1, MODE, false, true>, (int)tc::SMEM_BYTES_HELP));
                    LB_CUDA_TRY(cudaLaunchKernelEx(&cfg, tc::coarse_pair_kernel<128, 3, 1, MODE, false, true>, sh.tmap_full[0], sh.tmap_rem[0], args));
                }
            } else {
                return fail(LB_INTERNAL, "helper warps are a list-mode kernel shape for keys without side values");
            }
            break;
        default:
            if (HITS) {
                cfg.blockDim = dim3(64 + 128 * 2);
                LB_CUDA_TRY(ensure_dynamic_smem(tc::coarse_pair_kernel<128, 3, 2, MODE, true>, (int)tc::SMEM_BYTES));
                LB_CUDA_TRY(cudaLaunchKernelEx(&cfg, tc::coarse_pair_kernel<128, 3, 2, MODE, true>, sh.tmap_full[0], sh.tmap_rem[0], args));
            } else {
                return fail(LB_INTERNAL, "two epilogue sets are a hit-mode kernel shape");
            }
            break;
    }
    return LB_OK;
}
static int launch_coarse_any(int mode, bool hits, int cfg_id, const cudaLaunchConfig_t& cfg, const Shadow& sh, const tc::TcArgs& args) {
#define LB_MODE_CASE(M)                                                                       \
    case M:                                                                                   \
        return hits ? launch_coarse_mode<M, true>(cfg_id, cfg, sh, args) : launch_coarse_mode<M, false>(cfg_id, cfg, sh, args);
    switch (mode) {
        LB_MODE_CASE(tc::CM_F32)
        LB_MODE_CASE(tc::CM_F32_BIAS)
        LB_MODE_CASE(tc::CM_I32)
        LB_MODE_CASE(tc::CM_I32_HAMMING)
        default:
            return hits ? launch_coarse_mode<tc::CM_RATIO, true>(cfg_id, cfg, sh, args) : launch_coarse_mode<tc::CM_RATIO, false>(cfg_id, cfg, sh, args);
    }
#undef LB_MODE_CASE
}

// rows per accumulator tile of the kernel coarse_pass picks for operand rows of row_b bytes
static int plan_bn(int row_b, bool pair) {
    return pair && row_b / 4 <= tc::PairCfg<128, 2>::kDCol && tc_env_int("LYNSE_B200_TC_BN", 128) != 64 ? 128 : 64;
}

// Plans the partitions, seeds the floors when k is large, and launches the coarse kernel(s) on idx->stream.
static int coarse_pass(lb_index* idx, CoarseJob& job) {
    Shadow& sh = *job.sh;
    const int nq = job.nq, k = job.k;
    const int Dp = sh.Dp;
    const int row_b = 2 * Dp;
    const int n_mtiles = (nq + tc::BM - 1) / tc::BM;
    // one query tile: one CTA per partition (lb_tc1.cuh); more: CTA pairs (tcgen05 cta_group::2, lb_tc2.cuh)
    const bool pair = n_mtiles >= 2;
    const int cluster = pair ? 2 : 1;
    // kernel shape: 128-row accumulator tiles when the A operand (row_b / 4 TMEM columns) leaves room for two of them,
    // a third accumulator tile when it leaves room for three, two epilogue warp sets for narrow rows at large k
    int cfg_id = 0;
    int BN = 64;
    if (pair) {
        const bool bn128 = plan_bn(row_b, true) == 128;
        const bool nacc3 = bn128 && row_b / 4 <= tc::PairCfg<128, 3>::kDCol && tc_env_int("LYNSE_B200_TC_NACC", 3) == 3;
        cfg_id = nacc3 ? 3 : (bn128 ? 2 : 1);
        BN = bn128 ? 128 : 64;
    }
    const int n_mgroups = (n_mtiles + cluster - 1) / cluster;
    const uint32_t tiles_total = (uint32_t)ceil_div(idx->n, BN);
    // Slots: groups of n_mgroups co-resident clusters (one per query group) that stream the same row partitions in
    // lockstep, so every shadow tile comes from HBM once and is served to the other query groups of the slot from L2.
    const uint64_t G = (uint64_t)(idx->sm_count / cluster);  // clusters resident at once (one CTA per SM)
    uint64_t n_slots = std::max<uint64_t>(1, G / (uint64_t)n_mgroups);
    n_slots = std::min<uint64_t>(n_slots, tiles_total);
    n_slots = std::min<uint64_t>(n_slots, 4096 / tc::KP);
    // Large k (no single partition floor is far enough past rank k): a pre-pass over 1/64 of the corpus seeds every
    // query's floor (seed_floor_kernel) at about rank max(10 k, 512) of the corpus, and the main pass appends every row
    // above that floor to the query's hit buffer (hit mode): no shortlist upkeep in the accumulator hand-off.
    // (LYNSE_B200_TC_SEED_SMALLK=1 sends small k the same way.  Measured on C2: the main pass gets ~5 % faster — 0.12
    // appends per tile and warp instead of 0.16 shortlist insertions — but the pre-pass costs more than that, at 10M rows
    // (5.03 against 4.86 ms per step) and on a 1.25M-row shard (0.97 against 0.94 ms): off.)
    const bool seeded = (k > tc::KP - 4 || tc_env_int("LYNSE_B200_TC_SEED_SMALLK", 0) != 0) &&
                        tiles_total >= (uint32_t)(64 * 8) * (uint32_t)n_slots && job.dump == nullptr && tc_env_int("LYNSE_B200_TC_SEED", 1) != 0;
    const bool hit_mode = seeded && tc_env_int("LYNSE_B200_TC_HITS", 1) != 0;
    // narrow rows in hit mode: two sets of epilogue warps, each scanning one 64-row half of every tile into its own hit region
    if (hit_mode && cfg_id == 3 && tc_env_int("LYNSE_B200_TC_EPI", 2) == 2) cfg_id = 4;
    const uint64_t L = cfg_id == 4 ? 2 : 1;  // shortlists / hit regions per partition
    // List mode: the certification needs the largest partition floor T (the KP-th best coarse key of one partition) to sit
    // well below the k-th best score overall, so the union of the shortlists must reach far past rank k: aim at
    // P*KP >= 24*k candidates (measured on C3, k = 100: P = 36 leaves 1385 of 1024 queries uncertified, P = 72
    // five, P >= 144 none; on C2, k = 10, one partition per slot — P = 18, 288 candidates — certifies every query and
    // halves finalize's sort: 5.12 -> 5.05 ms per step at 10M rows, 0.867 -> 0.844 ms on a 1.25M-row shard).  With a seeded floor the lists only need room for the true top-k: P >= 0.75 k.
    // LYNSE_B200_TC_PARTS overrides.
    uint64_t parts_per_slot = 1;
    if (!hit_mode) {
        uint64_t want = seeded ? ((uint64_t)3 * k + 3) / 4 : ((uint64_t)24 * k + tc::KP - 1) / tc::KP;
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
    // Everything a search needs zeroed lives in ONE buffer, cleared with one memset (each separate clear is a launch in
    // front of the coarse kernel): [head 4 words | per-query flags nq | shared floors nq | pad] [second-best exchange
    // nq x PBEST_STRIDE] [lockstep progress n_slots x PROGRESS_STRIDE]
    const size_t z_flags = (((size_t)nq * 8 + 16) + 255) & ~(size_t)255;
    const size_t z_pbest = (size_t)nq * tc::PBEST_STRIDE * 4;
    const size_t z_progress = (size_t)n_slots * tc::PROGRESS_STRIDE * 4;
    LB_TRY(idx->w_flags.ensure(z_flags + z_pbest + z_progress));
    uint32_t* flags = idx->w_flags.as<uint32_t>();
    uint32_t* z_pbest_p = reinterpret_cast<uint32_t*>(idx->w_flags.as<char>() + z_flags);
    uint32_t* z_progress_p = reinterpret_cast<uint32_t*>(idx->w_flags.as<char>() + z_flags + z_pbest);
    LB_CUDA_TRY(cudaMemsetAsync(flags, 0, z_flags + z_pbest + z_progress, idx->stream));
    // hit regions: one per (query, partition, epilogue set); a region that is never visited (short corpus) must read 0
    const uint32_t hit_cap = (uint32_t)std::min<uint64_t>(1024, std::max<uint64_t>(64, TC_HIT_TOTAL / (P * L)));
    if (hit_mode) {
        LB_TRY(idx->w_hits.ensure((size_t)nq * P * L * hit_cap * 8));
        LB_TRY(idx->w_hit_count.ensure((size_t)nq * P * L * 4));
        LB_CUDA_TRY(cudaMemsetAsync(idx->w_hit_count.p, 0, (size_t)nq * P * L * 4, idx->stream));
    }

    tc::TcArgs a{};
    a.qb = job.qb;
    a.nq = nq;
    a.n_mtiles = n_mtiles;
    a.Dp = Dp;
    a.rem_kb = (Dp / tc::KBLK) % tc::KPS;
    a.n_ksteps = job.n_ksteps;
    a.n_rows = (uint32_t)idx->n;
    a.tiles_total = tiles_total;
    a.tiles_per_part = tiles_per_part;
    a.P = (int)P;
    a.lists_per_part = (int)L;
    a.allow_bits = job.d_allow;
    a.bias = job.bias;
    a.qaux = job.qaux;
    a.idesc_extra = job.idesc_extra;
    a.cand_key = idx->w_cand_score.as<uint32_t>();
    a.cand_row = idx->w_cand_row.as<uint32_t>();
    a.cand_thr = idx->w_cand_thr.as<uint32_t>();
    a.gthr = flags + 4 + nq;
    // Shared floors: a published floor must have enough rows above it to be far past rank k (the certification needs
    // the final floor well below the k-th best score).  For k <= KP - 4 the KP-th best key of one partition will do.
    a.share_floor = k <= tc::KP - 4 ? 1 : 0;
    if (seeded) a.share_floor = 2;
    // second-best exchange between the partitions of a query (list mode, every partition in flight at once, few enough of
    // them for one poll): LYNSE_B200_TC_PBEST=0 turns it off
    a.pbest2 = nullptr;
    a.pbest_first = 0;
    // Which of a list's best keys it publishes: the floor (the smallest over the P partitions) has depth x P rows above it.
    // Depth 2 (36 rows at P = 18) sits so close to the k = 10-th score that a batch now and then holds a query or two the
    // 8-bit bound cannot certify — on a 5M-row shard of C2 two of 1024, every step, each re-run by the exact scan: 2.4 -> 5.3
    // ms per step.  Depth 4 (72 rows): none on 1.25M / 2.5M / 5M / 10M rows.
    a.pbest_depth = std::max(2, std::min(4, tc_env_int("LYNSE_B200_TC_PBEST_DEPTH", 4)));
    // ... and only when that is at least 6 k rows (k = 10, P = 18: 72): fewer partitions keep the lists' own floors
    if (a.share_floor == 1 && parts_per_slot == 1 && L == 1 && P >= 8 && P <= (uint64_t)tc::PBEST_STRIDE &&
        (uint64_t)a.pbest_depth * P >= (uint64_t)6 * k && tc_env_int("LYNSE_B200_TC_PBEST", 1) != 0) {
        a.pbest2 = z_pbest_p;
        a.pbest_first = tc_env_int("LYNSE_B200_TC_PBEST", 1) == 2 && P >= (uint64_t)k + 4 ? 1 : 0;
    }
    a.hit_count = nullptr;
    a.hit_buf = nullptr;
    a.hit_cap = 0;
    a.error_flag = flags;
    a.dump = job.dump;
    a.n_slots = (int)n_slots;
    a.parts_per_slot = (int)parts_per_slot;
    a.window = tc_env_int("LYNSE_B200_TC_WINDOW", 16);
    a.poll_mask = tc_env_int("LYNSE_B200_TC_POLL", 7);
    a.debug_mode = tc_env_int("LYNSE_B200_TC_DEBUG", 0);
    a.prof = nullptr;
    if (getenv("LYNSE_B200_TC_PROF") != nullptr) {
        LB_TRY(idx->w_prof.ensure((size_t)idx->sm_count * 8 * 8));
        LB_CUDA_TRY(cudaMemsetAsync(idx->w_prof.p, 0, (size_t)idx->sm_count * 8 * 8, idx->stream));
        a.prof = idx->w_prof.as<unsigned long long>();
    }
    a.progress = nullptr;
    if (n_mgroups > 1 && a.window > 0) {
        a.progress = z_progress_p;
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
    if (idx->timing) cudaEventRecord(idx->ev[0], idx->stream);
    if (seeded) {
        // pre-pass: the first 1/64 of the tiles, one partition per slot, private floors, shortlists; then the seed
        tc::TcArgs sa = a;
        // the sample is as small as it can be for its 8th best key to sit near rank `want_rows` of the corpus:
        // 8 / want_rows of the tiles (1/64 at k = 32, 1/125 at k = 100), at least eight tiles per slot
        const uint64_t want_rows = std::max<uint64_t>((uint64_t)10 * k, 512);
        const uint32_t S = std::max<uint32_t>((uint32_t)std::min<uint64_t>(tiles_total, ceil_div((uint64_t)8 * tiles_total * BN, want_rows * BN)),
                                              (uint32_t)n_slots * 8);
        sa.tiles_total = S;
        sa.n_rows = (uint32_t)std::min<uint64_t>(idx->n, (uint64_t)S * BN);
        sa.tiles_per_part = (uint32_t)ceil_div(S, n_slots);
        sa.P = (int)ceil_div(S, sa.tiles_per_part);
        sa.parts_per_slot = 1;
        sa.share_floor = 0;
        sa.lists_per_part = 1;
        LB_TRY(launch_coarse_any(job.mode, false, cfg_id == 4 ? 3 : cfg_id, cfg, sh, sa));
        const int s_lists = sa.P;
        const int sm = next_pow2(s_lists * tc::KP);
        // aim at ~max(10 k, 512) rows of the whole corpus above the seeded floor
        int r = (int)ceil_div(want_rows * S, tiles_total);
        r = std::max(8, std::min(r, s_lists * tc::KP / 2));
        LB_CUDA_TRY(ensure_dynamic_smem(tc::seed_floor_kernel, sm * 8));
        tc::seed_floor_kernel<<<nq, 256, (size_t)sm * 8, idx->stream>>>(sa.cand_key, sa.cand_row, s_lists, sm, r, mode_int_key(job.mode) ? 1 : 0, a.gthr);
        LB_CUDA_TRY(cudaGetLastError());
        if (a.progress) LB_CUDA_TRY(cudaMemsetAsync(z_progress_p, 0, z_progress, idx->stream));
        idx->stats.kernels_launched += 2;
        if (hit_mode) {
            a.hit_count = idx->w_hit_count.as<uint32_t>();
            a.hit_buf = idx->w_hits.as<uint2>();
            a.hit_cap = hit_cap;
        }
    }
    // list mode, 128-row tiles, one partition per slot, keys without side values: the shortlists move to helper warps
    int main_cfg = cfg_id;
    if (!hit_mode && (cfg_id == 2 || cfg_id == 3) && parts_per_slot == 1 && (job.mode == tc::CM_I32 || job.mode == tc::CM_F32) &&
        job.dump == nullptr && tc_env_int("LYNSE_B200_TC_HELPER", 1) != 0)
        main_cfg = cfg_id == 2 ? 6 : 7;
    // ... with a pre-pass of the second-best exchange when the partitions are long enough for it to be a small share
    // (no row filter: group maxima cannot tell allowed rows from filtered ones)
    a.pre_tiles = 0;
    if (main_cfg >= 6 && a.pbest2 != nullptr && a.pbest_first == 0 && job.d_allow == nullptr) {
        // 1/16 of a partition, at most 32 tiles (measured on a 1.25M-row shard of C2, 543 tiles per partition: kernel
        // 0.685 ms without, 0.663 / 0.652 / 0.642 / 0.637 with 4 / 8 / 16 / 32 tiles)
        const int pre_env = tc_env_int("LYNSE_B200_TC_PRE", -1);
        // (C1, 44 tiles per partition: 0.135 / 0.125 / 0.110 ms with 2 / 4 / 8 tiles — short partitions are all "first tiles")
        const uint32_t pre_auto = std::min<uint32_t>(32, std::max<uint32_t>(tiles_per_part / 16, std::min<uint32_t>(8, tiles_per_part / 4)));
        a.pre_tiles = pre_env >= 0 ? std::min<int>(pre_env, (int)tiles_per_part / 4) : (int)pre_auto;
    }
    LB_TRY(launch_coarse_any(job.mode, hit_mode, main_cfg, cfg, sh, a));
    LB_CUDA_TRY(cudaGetLastError());
    if (idx->timing) cudaEventRecord(idx->ev[1], idx->stream);
    idx->stats.kernels_launched += 1;
    idx->stats.n_partitions = (uint32_t)P;
    job.n_lists = (int)(P * L);
    job.hit_mode = hit_mode;
    job.hit_cap = hit_cap;
    idx->stats.coarse_operand = sh.operand == tc::OPERAND_U8 ? 1u : 0u;
    idx->stats.coarse_hit_mode = hit_mode ? 1u : 0u;
    job.flags = flags;
    job.gthr = a.gthr;
    idx->pending_tc.grid = grid;
    idx->pending_tc.cluster = cluster;
    idx->pending_tc.n_slots = (int)n_slots;
    idx->pending_tc.P = (int)P;
    idx->pending_tc.pair = pair;
    return LB_OK;
}

static void fill_fin_candidates(lb_index* idx, const CoarseJob& job, tc::FinArgs& f) {
    f.cand_key = idx->w_cand_score.as<uint32_t>();
    f.cand_row = idx->w_cand_row.as<uint32_t>();
    f.cand_thr = idx->w_cand_thr.as<uint32_t>();
    f.P = job.n_lists;
    f.hit_count = job.hit_mode ? idx->w_hit_count.as<uint32_t>() : nullptr;
    f.hit_buf = job.hit_mode ? idx->w_hits.as<uint2>() : nullptr;
    f.hit_cap = job.hit_mode ? job.hit_cap : 0;
    f.gthr = job.gthr;
    f.int_key = mode_int_key(job.mode) ? 1 : 0;
    f.M1 = job.hit_mode ? TC_FIN_MAX : next_pow2(job.n_lists * tc::KP);
    f.R = std::min(1024, std::max(128, next_pow2(4 * job.k)));
    if (tc_env_int("LYNSE_B200_FIN_R", 0) > 0) f.R = std::max(next_pow2(job.k), next_pow2(tc_env_int("LYNSE_B200_FIN_R", 0)));  // diagnostics
    // first round: half the budget when that still is at least 2k candidates (LYNSE_B200_FIN_TWO_ROUNDS=0: one round)
    f.R1 = (tc_env_int("LYNSE_B200_FIN_TWO_ROUNDS", 1) != 0 && f.R >= 64 && f.R / 2 >= 2 * job.k) ? f.R / 2 : 0;
    f.nq = job.nq;
    f.k = job.k;
    f.uncertified = job.flags + 4;
    f.n_uncertified = job.flags + 1;
}

// ---- tensor-core plan, dense metrics ---------------------------------------------------------------------------
int run_tc(lb_index* idx, int metric, const float* d_queries, int nq, int k, uint32_t* d_rows, float* d_dists,
           uint32_t* d_counts, float* dump, const uint64_t* d_allow, bool defer_check) {
    const int kind = shadow_kind_for(metric);
    LB_TRY(ensure_shadow(idx, kind));
    LB_TRY(refresh_small_segments(idx));
    Shadow& sh = idx->shadow[kind];
    if (sh.disabled) return fail(LB_UNSUPPORTED, "the corpus holds non-finite values: the tensor-core plan is not available");
    const int row_b = 2 * sh.Dp;
    const int n_mtiles = (nq + tc::BM - 1) / tc::BM;
    const int cluster = n_mtiles >= 2 ? 2 : 1;
    const int nq_pad = (n_mtiles + cluster - 1) / cluster * cluster * tc::BM;
    LB_TRY(idx->w_qb.ensure((size_t)nq_pad * row_b));
    LB_TRY(idx->w_qnorm.ensure((size_t)nq * sizeof(tc::QStat)));
    const int warps = 8;
    CoarseJob job;
    job.sh = &sh;
    if (sh.operand == tc::OPERAND_U8) {
        LB_TRY(idx->w_qrange.ensure((size_t)nq * 8 + 16));
        uint32_t* idesc_extra = reinterpret_cast<uint32_t*>(idx->w_qrange.as<float>() + 2 * (size_t)nq);
        LB_CUDA_TRY(cudaMemsetAsync(idesc_extra, 0, 4, idx->stream));
        tc::query_range_kernel<<<(nq + warps - 1) / warps, warps * 32, 0, idx->stream>>>(d_queries, nq, (int)idx->dim, kind,
                                                                                       idx->w_qrange.as<float>(), idesc_extra);
        tc::quantise_queries_kernel<<<(nq_pad + warps - 1) / warps, warps * 32, 0, idx->stream>>>(
            d_queries, nq, nq_pad, (int)idx->dim, row_b, kind, idx->w_qrange.as<float>(), idesc_extra, idx->w_qb.as<unsigned char>(),
            idx->w_qnorm.as<tc::QStat>());
        LB_CUDA_TRY(cudaGetLastError());
        idx->stats.kernels_launched += 1;
        job.idesc_extra = idesc_extra;
        job.mode = tc::CM_I32;
        job.n_ksteps = ((int)idx->dim + 31) / 32;
    } else {
        const bool norm_cols = kind == tc::SHADOW_L2 && !sh.l2_bias;
        tc::prepare_queries_kernel<<<(nq_pad + warps - 1) / warps, warps * 32, 0, idx->stream>>>(
            d_queries, nq, nq_pad, (int)idx->dim, row_b, kind, norm_cols ? 1 : 0, idx->w_qb.as<unsigned char>(), idx->w_qnorm.as<tc::QStat>());
        LB_CUDA_TRY(cudaGetLastError());
        job.mode = sh.l2_bias ? tc::CM_F32_BIAS : tc::CM_F32;
        job.bias = sh.l2_bias ? sh.side.as<uint32_t>() : nullptr;
        job.n_ksteps = ((int)idx->dim + (norm_cols ? 3 : 0) + 15) / 16;
    }
    job.qb = idx->w_qb.as<unsigned char>();
    job.nq = nq;
    job.k = k;
    job.d_allow = d_allow;
    job.dump = dump;
    LB_TRY(coarse_pass(idx, job));

    tc::FinArgs f{};
    fill_fin_candidates(idx, job, f);
    f.corpus = idx->rows.p;
    f.dim = (int)idx->dim;
    f.queries = d_queries;
    f.qstat = idx->w_qnorm.as<tc::QStat>();
    f.sstat = sh.stats.as<tc::ShadowStats>();
    f.operand = sh.operand;
    f.c_scale = sh.c_scale;
    f.c_zero = sh.c_zero;
    f.metric = metric;
    f.small_seg = idx->small_seg.as<uint32_t>();
    f.n_small = idx->n_small;
    f.out_rows = d_rows;
    f.out_dists = d_dists;
    f.out_counts = d_counts;
    int fin_threads = nq <= 64 ? 1024 : 256;  // few queries: few blocks, so each gets 1024 threads
    if (tc_env_int("LYNSE_B200_FIN_THREADS", 0) > 0) fin_threads = tc_env_int("LYNSE_B200_FIN_THREADS", 0);  // diagnostics
    size_t fsmem = (size_t)(f.M1 + f.R) * 8 + (size_t)((idx->dim + 3) & ~3u) * 4 + (size_t)(fin_threads / 8) * tc::FIN_STRIDE * 4;  // + row buffers
#define LB_LAUNCH_FIN(ASCV, RTV)                                                                 \
    do {                                                                                        \
        LB_CUDA_TRY(ensure_dynamic_smem(tc::finalize_kernel<ASCV, RTV>, (int)fsmem));           \
        tc::finalize_kernel<ASCV, RTV><<<nq, fin_threads, fsmem, idx->stream>>>(f);             \
    } while (0)
    if (idx->dtype == LB_F16) {
        if (metric_ascending(metric)) LB_LAUNCH_FIN(true, __half);
        else LB_LAUNCH_FIN(false, __half);
    } else {
        if (metric_ascending(metric)) LB_LAUNCH_FIN(true, float);
        else LB_LAUNCH_FIN(false, float);
    }
#undef LB_LAUNCH_FIN
    LB_CUDA_TRY(cudaGetLastError());
    idx->stats.kernels_launched += 2;
    idx->stats.plan_used = 1;
    idx->stats.algorithmic_bytes = (uint64_t)idx->n * row_b;
    idx->stats.algorithmic_flops = 2ull * (uint64_t)nq * idx->n * idx->dim;

    lb_index::PendingTc& p = idx->pending_tc;
    p.active = true;
    p.metric = metric;
    p.nq = nq;
    p.k = k;
    p.bits = 0;
    p.d_queries = d_queries;
    p.d_allow = d_allow;
    p.d_rows = d_rows;
    p.d_dists = d_dists;
    p.d_counts = d_counts;
    if (defer_check) return LB_OK;
    return tc_finish(idx);
}

// ---- tensor-core plan, binary metrics ---------------------------------------------------------------------------
int run_tc_bits(lb_index* idx, int metric, const uint64_t* words, int n_words, const uint64_t* d_qwords, int nq, int k, uint32_t* d_rows,
                float* d_dists, uint32_t* d_counts, const uint64_t* d_allow, bool defer_check) {
    LB_TRY(ensure_bits_shadow(idx, words, n_words));
    Shadow& sh = idx->bits_shadow;
    const int row_b = 2 * sh.Dp;
    const int n_mtiles = (nq + tc::BM - 1) / tc::BM;
    const int cluster = n_mtiles >= 2 ? 2 : 1;
    const int nq_pad = (n_mtiles + cluster - 1) / cluster * cluster * tc::BM;
    LB_TRY(idx->w_qb.ensure((size_t)nq_pad * row_b));
    LB_TRY(idx->w_qaux.ensure((size_t)nq_pad * 4));
    const int warps = 8;
    tc::prepare_bits_queries_kernel<<<(nq_pad + warps - 1) / warps, warps * 32, 0, idx->stream>>>(d_qwords, nq, nq_pad, n_words, row_b,
                                                                                                idx->w_qb.as<unsigned char>(), idx->w_qaux.as<float>());
    LB_CUDA_TRY(cudaGetLastError());
    CoarseJob job;
    job.sh = &sh;
    job.mode = metric == LB_HAMMING ? tc::CM_I32_HAMMING : tc::CM_RATIO;
    job.qb = idx->w_qb.as<unsigned char>();
    job.qaux = idx->w_qaux.as<float>();
    job.bias = metric == LB_HAMMING ? sh.side.as<uint32_t>() : sh.side2.as<uint32_t>();
    job.n_ksteps = 2 * n_words;
    job.nq = nq;
    job.k = k;
    job.d_allow = d_allow;
    LB_TRY(coarse_pass(idx, job));

    tc::FinBitsArgs fb{};
    fill_fin_candidates(idx, job, fb.f);
    fb.f.metric = metric;
    fb.f.out_rows = d_rows;
    fb.f.out_dists = d_dists;
    fb.f.out_counts = d_counts;
    fb.words = words;
    fb.qwords = d_qwords;
    fb.n_words = n_words;
    const size_t fsmem = (size_t)(fb.f.M1 + fb.f.R) * 8 + (size_t)n_words * 8;
    LB_CUDA_TRY(ensure_dynamic_smem(tc::finalize_bits_kernel, (int)fsmem));
    tc::finalize_bits_kernel<<<nq, 256, fsmem, idx->stream>>>(fb);
    LB_CUDA_TRY(cudaGetLastError());
    idx->stats.kernels_launched += 2;
    idx->stats.plan_used = 3;
    idx->stats.algorithmic_bytes = (uint64_t)idx->n * row_b;
    idx->stats.algorithmic_flops = 2ull * (uint64_t)nq * idx->n * (uint64_t)n_words * 64;

    lb_index::PendingTc& p = idx->pending_tc;
    p.active = true;
    p.metric = metric;
    p.nq = nq;
    p.k = k;
    p.bits = 1;
    p.n_words = n_words;
    p.words = words;
    p.d_queries = d_qwords;
    p.d_allow = d_allow;
    p.d_rows = d_rows;
    p.d_dists = d_dists;
    p.d_counts = d_counts;
    if (defer_check) return LB_OK;
    return tc_finish(idx);
}

// ---- certification flags -> exact-scan fallback -----------------------------------------------------------------
int tc_finish(lb_index* idx, bool* changed, const uint32_t* head_ready) {
    if (changed) *changed = false;
    lb_index::PendingTc& p = idx->pending_tc;
    if (!p.active) return LB_OK;
    p.active = false;
    uint32_t* flags = idx->w_flags.as<uint32_t>();
    uint32_t head[4] = {0, 0, 0, 0};
    if (head_ready != nullptr) {
        memcpy(head, head_ready, 16);
    } else {
        LB_CUDA_TRY(cudaMemcpyAsync(head, flags, 16, cudaMemcpyDeviceToHost, idx->stream));
        LB_CUDA_TRY(cudaStreamSynchronize(idx->stream));
    }
    if (idx->timing) {
        float ms = 0;
        cudaEventElapsedTime(&ms, idx->ev[0], idx->ev[1]);
        idx->stats.ms_dominant = ms;
    }
    const int nq = p.nq, k = p.k, grid = p.grid;
    if (getenv("LYNSE_B200_TC_PROF") != nullptr && idx->w_prof.p) {
        std::vector<unsigned long long> pr((size_t)idx->sm_count * 8);
        LB_CUDA_TRY(cudaMemcpy(pr.data(), idx->w_prof.p, pr.size() * 8, cudaMemcpyDeviceToHost));
        double sum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        int n_lead = 0, n_cta = 0;
        if (p.pair) {
            double sc[5] = {0, 0, 0, 0, 0};
            double mx = 0, h_busy = 0, h_groups = 0;
            for (int b = 1; b < grid && b < idx->sm_count; b += 2) {
                for (int i = 0; i < 5; ++i) sc[i] += (double)pr[(size_t)b * 8 + i];
                mx = std::max(mx, (double)pr[(size_t)b * 8 + 3]);
                h_busy += (double)(pr[(size_t)b * 8 + 5] >> 24);           // helper-warp kernel: one helper warp per odd CTA
                h_groups += (double)(pr[(size_t)b * 8 + 5] & 0xFFFFFFull);
                for (int i = 0; i < 6; ++i) pr[(size_t)b * 8 + i] = 0;
            }
            if (sc[4] > 0)
                fprintf(stderr, "[lynse_b200] scan (one warp per odd CTA): %.0f cycles/tile, slow tiles %.3f/tile at %.0f cycles each, longest %.0f\n",
                        sc[0] / sc[4], sc[2] / sc[4], sc[2] > 0 ? sc[1] / sc[2] : 0.0, mx);
            if (sc[4] > 0 && h_groups > 0)
                fprintf(stderr, "[lynse_b200] helper (one warp per odd CTA): %.3f queued groups per tile, %.0f cycles per group\n", h_groups / sc[4],
                        h_busy / h_groups);
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
        fprintf(stderr, "[lynse_b200] coarse kernel: %.3f ms, %.0f SM MHz, grid %d, cluster %d, slots %d, P %d, uncertified %u\n", head[3] * 1e-6,
                (double)head[2] * 16.0 / (double)head[3] * 1e3, grid, p.cluster, p.n_slots, p.P, head[1]);
    if (head[0] != 0)
        return fail(LB_INTERNAL, "tensor-core coarse kernel: barrier wait timed out (code " + std::to_string(head[0]) + ")");
    idx->stats.n_fallback = head[1];
    if (head[3] > 0) idx->stats.coarse_sm_mhz = (float)((double)head[2] * 16.0 / (double)head[3] * 1e3);  // cycles >> 4 over nanoseconds of CTA 0
    if (head[1] == 0) return LB_OK;
    // Re-run the uncertified queries with the exact scan and overwrite their result slots.
    if (changed) *changed = true;
    std::vector<uint32_t> fl(nq);
    LB_CUDA_TRY(cudaMemcpy(fl.data(), flags + 4, (size_t)nq * 4, cudaMemcpyDeviceToHost));
    std::vector<uint32_t> qmap;
    for (int q = 0; q < nq; ++q)
        if (fl[q]) qmap.push_back((uint32_t)q);
    const int ns = (int)qmap.size();
    const size_t qrow = p.bits ? (size_t)p.n_words * 8 : (size_t)idx->dim * 4;
    LB_TRY(idx->w_sub_q.ensure((size_t)ns * qrow));
    LB_TRY(idx->w_qmap.ensure((size_t)ns * 4));
    LB_CUDA_TRY(cudaMemcpyAsync(idx->w_qmap.p, qmap.data(), (size_t)ns * 4, cudaMemcpyHostToDevice, idx->stream));
    for (int i = 0; i < ns; ++i)
        LB_CUDA_TRY(cudaMemcpyAsync(idx->w_sub_q.as<unsigned char>() + (size_t)i * qrow,
                                    reinterpret_cast<const unsigned char*>(p.d_queries) + (size_t)qmap[i] * qrow, qrow, cudaMemcpyDeviceToDevice,
                                    idx->stream));
    ScanRequest r;
    if (p.bits) {
        r.words = p.words;
        r.n_words = p.n_words;
        r.qwords = idx->w_sub_q.as<uint64_t>();
    } else {
        if (idx->dtype == LB_F16) r.corpus_h = idx->rows.as<__half>();
        else r.corpus = idx->rows.as<float>();
        r.dim = (int)idx->dim;
        r.queries = idx->w_sub_q.as<float>();
        r.small_seg = idx->small_seg.as<uint32_t>();
        r.n_small = idx->n_small;
    }
    r.n_rows = idx->n;
    r.nq = ns;
    r.k = k;
    r.metric = p.metric;
    r.allow_bits = p.d_allow;
    r.qmap = idx->w_qmap.as<uint32_t>();
    r.out_rows = p.d_rows;
    r.out_dists = p.d_dists;
    r.out_counts = p.d_counts;
    int kern = 0;
    const lb_search_stats keep = idx->stats;
    LB_TRY(run_scan(idx, r, &kern, nullptr));
    idx->stats = keep;
    idx->stats.kernels_launched += kern;
    LB_CUDA_TRY(cudaStreamSynchronize(idx->stream));
    return LB_OK;
}

}  // namespace lb

extern "C" {

// ---- diagnostics ------------------------------------------------------------------------------------------------------------------
int lb_debug_tc_scores(const float* queries, uint32_t nq, const float* rows, uint32_t n, uint32_t dim, int operand, float* out) {
    if (!queries || !rows || !out || nq == 0 || n == 0) return fail(LB_INVALID_ARGUMENT, "bad arguments");
    if (operand != tc::OPERAND_BF16 && operand != tc::OPERAND_U8) return fail(LB_INVALID_ARGUMENT, "operand must be 0 (bf16) or 1 (u8)");
    int device = 0;
    LB_CUDA_TRY(cudaGetDevice(&device));
    lb_index* idx = nullptr;
    LB_TRY(lb_index_create(&idx, dim, LB_F32, device));
    int st = lb_index_append_f32(idx, rows, n);
    float* dump = nullptr;
    if (st == LB_OK && operand_row_bytes((int)dim, operand) > TC_MAX_ROW_BYTES) st = fail(LB_UNSUPPORTED, "dimension too large for the tensor-core path");
    if (st == LB_OK) {
        std::lock_guard<std::mutex> lock(idx->mu);
        DeviceGuard g(idx->device);
        idx->shadow[tc::SHADOW_IP].operand = operand;
        int n_mtiles = ((int)nq + tc::BM - 1) / tc::BM;
        n_mtiles = (n_mtiles + 1) & ~1;  // room for the padded query tile of a 2-CTA cluster
        const size_t ld = (size_t)ceil_div(n, 128) * 128;  // the kernel's own leading dimension is tiles * rows per tile <= this
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
            const int bn = plan_bn(2 * idx->shadow[tc::SHADOW_IP].Dp, (int)nq > tc::BM);
            const size_t kld = (size_t)ceil_div(n, bn) * bn;
            e = cudaMemcpy2D(out, (size_t)n * 4, dump, kld * 4, (size_t)n * 4, nq, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) st = fail(LB_CUDA, cudaGetErrorString(e));
        }
    }
    if (dump) cudaFree(dump);
    std::string keep = lb_last_error();
    lb_index_destroy(idx);
    set_error(keep);
    return st;
}

}  // extern "C"
