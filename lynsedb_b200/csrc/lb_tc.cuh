// lb_tc.cuh — tensor-core coarse pass (tcgen05 / TMEM / TMA) + exact-order finalize.
//
// The dense metrics (IP, and through it cosine and L2) are a Q x C^T contraction
// (reference hot loop: src/storage/flat_mmap.rs:2179-2256 ip_scan_chunk_topk over
// simd::inner_product_batch8_f32, src/distance/simd.rs:1450-1525).  On B200 that
// contraction runs on the 5th-generation tensor cores over a bf16 shadow of the
// corpus; a per-query shortlist is kept in the accumulator epilogue, and the
// shortlist is re-scored in f32 in the reference's exact summation order, so the
// returned ids / order / scores are the reference's, not the bf16 ones.  A
// shortlist is only accepted when a rigorous bound proves no dropped row could
// enter the top-k (see finalize_kernel); otherwise the query is re-run by the
// exact scan of lb_scan.cuh.
//
// Kernel shape (one persistent CTA per SM, 192 threads):
//   warp 0      TMA producer: 64-row x 64-col bf16 boxes of the shadow, SWIZZLE_128B,
//               6 stages x 32 KiB in shared memory, mbarrier full/empty ring
//   warp 1      MMA issuer: tcgen05.mma.cta_group::1.kind::f16, M=128 (queries), N=64 (rows), K=16;
//               A (the 128 queries of this work item, bf16) lives in TMEM columns [0, Dp/2),
//               B comes from the shared-memory stages, D is double-buffered in TMEM columns [384,512)
//   warps 2..5  epilogue: lane == query.  tcgen05.ld the 64 scores of the tile, compare with the
//               thread's running threshold held in a register, insert the rare survivors in a
//               private 16-entry list in shared memory
// Work item = (query tile of 128, row partition); items of the same partition run on
// neighbouring CTAs at the same time so each shadow tile is fetched from HBM once and
// served to the other query tiles from L2.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include <type_traits>

#include "lb_metrics.cuh"
#include "lb_scan.cuh"

namespace lb {
namespace tc {

constexpr int BM = 128;           // queries per work item == TMEM lanes
constexpr int BN = 64;            // default corpus rows per accumulator tile (see TileCfg)
constexpr int KBLK = 64;          // bf16 elements per 128-byte swizzle row
constexpr int NSTAGES = 6;
constexpr int KP = 16;            // shortlist entries kept per (partition, query)
constexpr int MAX_DP = 768;       // padded dim limit: A occupies Dp/2 <= 384 TMEM columns
constexpr int STAGE_BYTES = 32768;  // one pipeline stage of corpus K-blocks
constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = 512;
constexpr uint32_t SMEM_LIST_OFF = NSTAGES * STAGE_BYTES;             // 196608
constexpr uint32_t SMEM_BAR_OFF = SMEM_LIST_OFF + 2 * KP * BM * 4;    // + 16384
constexpr uint32_t SMEM_BYTES = SMEM_BAR_OFF + 256 + 1024;            // + barriers + alignment slack

struct TcArgs {
    const __nv_bfloat16* qb;  // [n_mtiles*128][Dp] bf16 queries, zero padded
    int nq;
    int n_mtiles;             // query tiles of 128; qb is padded to a multiple of the cluster size tiles
    int Dp;                   // padded dim, multiple of 64, <= MAX_DP
    uint32_t n_rows;
    uint32_t tiles_total;     // ceil(n_rows / 64)
    uint32_t tiles_per_part;
    int P;                    // row partitions
    float* cand_score;        // [nq][P][KP]
    uint32_t* cand_row;       // [nq][P][KP]
    float* cand_thr;          // [nq][P]
    int share_floor;          // 1: partitions of a query share their shortlist floor through gthr (needs k <= KP - 4)
    uint32_t* gthr;           // [nq] zero-initialised: best published shortlist floor per query (orderable f32 bits)
    uint32_t* error_flag;     // set non-zero when a barrier wait timed out
    float* dump;              // optional [n_mtiles*128][tiles_total*64] raw scores (diagnostics)
    // Work mapping: cluster c serves query group (c % n_mgroups) of slot (c / n_mgroups); slot s walks the row
    // partitions s, s + n_slots, s + 2 n_slots, ...  All query groups of a slot stream the same shadow tiles at the
    // same time, so HBM is read once per slot and the other groups are served from L2.
    int n_slots;
    int parts_per_slot;
    uint32_t* progress;       // [n_slots][PROGRESS_STRIDE] tiles issued per (slot, query group); zeroed per launch; null = free-running
    int window;               // a query group never runs more than `window` tiles ahead of the slowest group of its slot
    int prefetch_tiles;       // L2 prefetch distance of the TMA producer, in tiles (0 = off)
    unsigned long long* prof; // optional [grid][8] cycle counters of the MMA issuer / epilogue (diagnostics)
    int debug_mode;           // diagnostics only (results are garbage): bit 0 = producer skips the TMA loads,
                              // bit 1 = epilogue releases accumulators unread, bit 2 = epilogue reads but does not scan
};
constexpr int PROGRESS_STRIDE = 32;

// ---- PTX wrappers ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
// non-blocking probe (try_wait may suspend the thread for a hardware time slice; test_wait never does)
__device__ __forceinline__ uint32_t mbar_test_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
// Bounded wait: a barrier that does not flip within ~2 s marks the launch as failed and lets every role drain,
// so a protocol bug can never hang the GPU.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, volatile uint32_t* abort_flag, uint32_t code) {
    // fast path: try_wait itself suspends the thread for a hardware time slice, so spin on it alone; the clock and
    // the abort flag are only consulted every 1024 failed probes (reading %globaltimer is slow)
    uint32_t spins = 0;
    uint64_t t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 1023u) == 0u) {
            if (*abort_flag) return false;
            const uint64_t now = globaltimer_ns();
            if (t0 == 0) t0 = now;
            if (now - t0 > 2000000000ull) {
                *abort_flag = code;
                return false;
            }
        }
    }
    return true;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T, bf16 x bf16 -> f32
__device__ __forceinline__ void umma_ts_bf16(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* tmap, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
// multicast variants: the box lands at the same CTA-relative offset of every CTA in cta_mask and completes
// tx bytes on the mbarrier at the same offset there; the commit arrives on the mbarrier of every CTA in cta_mask
__device__ __forceinline__ void tma_load_2d_mcast(uint32_t smem_dst, const CUtensorMap* tmap, int c0, int c1, uint32_t bar,
                                                  uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
        "[%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(smem_dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar), "h"(cta_mask)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* tmap, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(tmap), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void umma_commit_mcast(uint32_t bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major), bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, both K-major, N>>3 at [17,23), M>>4 at [24,29)

// ---- the coarse kernel -----------------------------------------------------------------------------------------
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// Tile shape: BN_ corpus rows per accumulator, NBUF accumulators in TMEM columns [512 - NBUF*BN_, 512);
// the A operand (queries) needs Dp/2 <= 512 - NBUF*BN_ columns.  Stages are always 32 KiB:
// KPS = 32768 / (BN_*128) K-blocks of [BN_ rows x 64 bf16] each.
template <int BN_, int NBUF>
struct TileCfg {
    static constexpr int kBN = BN_;
    static constexpr int kNBuf = NBUF;
    static constexpr int kDCol = TMEM_COLS - NBUF * BN_;
    static constexpr int kMaxDp = 2 * kDCol;
    static constexpr int kTileBytes = BN_ * KBLK * 2;
    static constexpr int kKPS = STAGE_BYTES / kTileBytes;
    static constexpr uint32_t kIdesc =
        (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN_ >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
};

// CLUSTER = 2: two CTAs of a cluster work on the same row partition for two neighbouring query tiles; each loads
// half of every corpus K-block and multicasts it into both CTAs' shared memory, so every shadow byte crosses
// the L2 -> SM fabric once per CTA pair instead of once per CTA.
template <int BN_, int NBUF, int CLUSTER>
__global__ void __launch_bounds__(NUM_THREADS, 1)
coarse_topk_kernel(const __grid_constant__ CUtensorMap tmap, TcArgs a) {
    using Cfg = TileCfg<BN_, NBUF>;
    constexpr int KPS = Cfg::kKPS;
    constexpr uint16_t kMask = (uint16_t)((1u << CLUSTER) - 1u);
    const uint32_t crank = CLUSTER > 1 ? cluster_ctarank() : 0u;
    // this cluster: query group `mgroup` (query tile mgroup*CLUSTER + crank in this CTA) of slot `slot`
    const int n_mgroups = (a.n_mtiles + CLUSTER - 1) / CLUSTER;
    const int cluster_id = (int)(blockIdx.x / CLUSTER);
    const int mgroup = cluster_id % n_mgroups, slot = cluster_id / n_mgroups;
    const int n_rounds = slot < a.n_slots ? a.parts_per_slot : 0;
    extern __shared__ __align__(16) unsigned char smem_tc[];
    const uint32_t smem_base = (smem_u32(smem_tc) + 1023u) & ~1023u;
    unsigned char* smem = smem_tc + (smem_base - smem_u32(smem_tc));
    float* l_score = reinterpret_cast<float*>(smem + SMEM_LIST_OFF);                     // [KP][BM]
    uint32_t* l_row = reinterpret_cast<uint32_t*>(smem + SMEM_LIST_OFF + KP * BM * 4);   // [KP][BM]
    const uint32_t bar_base = smem_base + SMEM_BAR_OFF;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (NSTAGES + s); };
    auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * NSTAGES + b); };
    auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * NSTAGES + 2 + b); };
    const uint32_t aready_bar = bar_base + 8u * (2 * NSTAGES + 4);
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + SMEM_BAR_OFF + 8 * (2 * NSTAGES + 5));
    volatile uint32_t* abort_flag = reinterpret_cast<volatile uint32_t*>(smem + SMEM_BAR_OFF + 8 * (2 * NSTAGES + 5) + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), CLUSTER);  // every CTA of the cluster must have drained the stage
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tfull_bar(b), 1);
            mbar_init(tempty_bar(b), 128);
        }
        mbar_init(aready_bar, 128);
        *abort_flag = 0;
        fence_barrier_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    }
    if (warp == 0) {
        __syncwarp();
        tmem_alloc(smem_u32(tmem_ptr_smem), TMEM_COLS);
        tmem_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    if (CLUSTER > 1) cluster_sync_all();  // peers' barriers are initialised before anything remote can arrive
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int nkb = a.Dp / KBLK;
    const int stages_per_tile = (nkb + KPS - 1) / KPS;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t stage_iter = 0;
            bool ok = true;
            // lockstep: only the rank-0 CTA of a cluster throttles; its peer follows through the shared stage ring
            bool lockstep = a.progress != nullptr && n_mgroups > 1 && crank == 0;
            uint32_t* prog = a.progress != nullptr ? a.progress + (size_t)slot * PROGRESS_STRIDE : nullptr;
            uint32_t seq = 0, known_min = 0;
            const uint32_t window = (uint32_t)a.window;
            const int pf = a.prefetch_tiles;
            constexpr int kRowsPer = BN_ / CLUSTER;
            auto prefetch_tile = [&](uint32_t t) {
                for (int kb = 0; kb < nkb; ++kb) tma_prefetch_2d(&tmap, kb * KBLK, (int)(t * BN_ + crank * kRowsPer));
            };
            for (int r = 0; r < n_rounds && ok; ++r) {
                const uint32_t part = (uint32_t)slot + (uint32_t)r * (uint32_t)a.n_slots;
                if (part >= (uint32_t)a.P) break;
                const uint32_t t0 = part * a.tiles_per_part;
                const uint32_t t1 = min(t0 + a.tiles_per_part, a.tiles_total);
                if (pf > 0)
                    for (uint32_t t = t0; t < min(t0 + (uint32_t)pf, t1); ++t) prefetch_tile(t);
                for (uint32_t t = t0; t < t1 && ok; ++t) {
                    if (lockstep && seq >= known_min + window) {
                        const uint64_t w0 = globaltimer_ns();
                        while (true) {
                            uint32_t mn = 0xFFFFFFFFu;
                            for (int m = 0; m < n_mgroups; ++m) mn = min(mn, ld_relaxed_gpu(prog + m));
                            known_min = mn;
                            if (seq < known_min + window) break;
                            if (globaltimer_ns() - w0 > 20000000ull) {  // a peer is not making progress: run free
                                lockstep = false;
                                break;
                            }
                            __nanosleep(100);
                        }
                    }
                    if (pf > 0 && t + (uint32_t)pf < t1) prefetch_tile(t + (uint32_t)pf);
                    for (int s = 0; s < stages_per_tile; ++s, ++stage_iter) {
                        const int stage = stage_iter % NSTAGES;
                        const uint32_t phase = (stage_iter / NSTAGES) & 1u;
                        if (!mbar_wait(empty_bar(stage), phase ^ 1u, abort_flag, 1)) { ok = false; break; }
                        const int kbc = min(KPS, nkb - s * KPS);
                        if (a.debug_mode & 1) {
                            mbar_arrive(full_bar(stage));
                            continue;
                        }
                        mbar_arrive_expect_tx(full_bar(stage), (uint32_t)kbc * Cfg::kTileBytes);
                        for (int kb = 0; kb < kbc; ++kb) {
                            if (CLUSTER == 1) {
                                tma_load_2d(smem_base + stage * STAGE_BYTES + kb * Cfg::kTileBytes, &tmap,
                                            (s * KPS + kb) * KBLK, (int)(t * BN_), full_bar(stage));
                            } else {
                                // this CTA fetches rows [crank*BN/CLUSTER, ...) of the K-block for the whole cluster
                                tma_load_2d_mcast(smem_base + stage * STAGE_BYTES + kb * Cfg::kTileBytes + crank * (kRowsPer * 128),
                                                  &tmap, (s * KPS + kb) * KBLK, (int)(t * BN_ + crank * kRowsPer),
                                                  full_bar(stage), kMask);
                            }
                        }
                    }
                    ++seq;
                    if (prog != nullptr && crank == 0 && n_mgroups > 1) st_relaxed_gpu(prog + mgroup, seq);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp walks the loops; one elected lane issues) =====================
        // tcgen05.mma issue is paced by execution (about one MMA of queue slack), so every cycle spent between two
        // MMAs on anything else is a tensor-pipe bubble.  The barrier of the NEXT stage (and the accumulator buffer
        // of the next tile) is therefore probed in the shadow of the current stage's MMAs, and the blocking wait is
        // only taken when that probe failed.  The loops stay warp-uniform so the descriptors live in uniform registers.
        const bool leader = elect_one();
        uint32_t stage_iter = 0, tile_iter = 0, item_iter = 0;
        bool ok = true;
        uint32_t full_ready = 0;    // full barrier of stage_iter already observed complete
        uint32_t tempty_ready = 0;  // accumulator buffer of tile_iter already observed free
        bool n_rounds_done = false;
        // Fast path (Dp == 768 with 64-row tiles): a tile is exactly 3 stages, so the 6-stage ring holds two tiles and
        // tile parity fixes the stage numbers, the accumulator buffer and every barrier / descriptor address at
        // compile time; only the phase bits are computed at run time.
        if (BN_ == 64 && NBUF == 2 && KPS == 4 && nkb == 12 && !(a.debug_mode & 8)) {
            const uint64_t desc_base = make_b_desc(smem_base);
            auto tile_body = [&](auto bc, uint32_t ph) -> bool {
                constexpr int b = decltype(bc)::value;
                if (!tempty_ready && !mbar_wait(tempty_bar(b), ph ^ 1u, abort_flag, 3)) return false;
                tempty_ready = 0;
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + Cfg::kDCol + b * BN_;
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                    constexpr int dummy = 0;
                    (void)dummy;
                    const int stage = b * 3 + s;
                    if (!full_ready && !mbar_wait(full_bar(stage), ph, abort_flag, 4)) return false;
                    if (leader) {
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4)
                            umma_ts_bf16(d_tmem, tmem_base + (uint32_t)((s * 16 + k4) * 8),
                                         desc_base + (uint64_t)((stage * STAGE_BYTES) >> 4) + (uint64_t)(k4 * 2), Cfg::kIdesc,
                                         (k4 == 0 && s == 0) ? 0u : 1u);
                    }
                    if (s < 2) {
                        full_ready = mbar_try_wait(full_bar(stage + 1), ph);
                    } else {
                        full_ready = mbar_try_wait(full_bar((1 - b) * 3), b == 0 ? ph : (ph ^ 1u));
                        tempty_ready = mbar_try_wait(tempty_bar(1 - b), b == 0 ? (ph ^ 1u) : ph);
                    }
                    if (leader) {
#pragma unroll
                        for (int kb = 1; kb < 4; ++kb)
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4)
                                umma_ts_bf16(d_tmem, tmem_base + (uint32_t)((s * 16 + kb * 4 + k4) * 8),
                                             desc_base + (uint64_t)((stage * STAGE_BYTES + kb * Cfg::kTileBytes) >> 4) + (uint64_t)(k4 * 2),
                                             Cfg::kIdesc, 1u);
                        if (CLUSTER == 1) umma_commit(empty_bar(stage));
                        else umma_commit_mcast(empty_bar(stage), kMask);
                    }
                    __syncwarp();
                }
                if (leader) umma_commit(tfull_bar(b));
                __syncwarp();
                return true;
            };
            for (int r = 0; r < n_rounds && ok; ++r, ++item_iter) {
                const uint32_t part = (uint32_t)slot + (uint32_t)r * (uint32_t)a.n_slots;
                if (part >= (uint32_t)a.P) break;
                const uint32_t t0 = part * a.tiles_per_part;
                const uint32_t t1 = min(t0 + a.tiles_per_part, a.tiles_total);
                if (!mbar_wait(aready_bar, item_iter & 1u, abort_flag, 2)) break;
                tcgen05_fence_after();
                for (uint32_t t = t0; t < t1 && ok; ++t, ++tile_iter) {
                    const uint32_t ph = (tile_iter >> 1) & 1u;
                    if ((tile_iter & 1u) == 0u) ok = tile_body(std::integral_constant<int, 0>{}, ph);
                    else ok = tile_body(std::integral_constant<int, 1>{}, ph);
                }
            }
            n_rounds_done = true;
        }
        for (int r = 0; r < n_rounds && ok && !n_rounds_done; ++r, ++item_iter) {
            const uint32_t part = (uint32_t)slot + (uint32_t)r * (uint32_t)a.n_slots;
            if (part >= (uint32_t)a.P) break;
            const uint32_t t0 = part * a.tiles_per_part;
            const uint32_t t1 = min(t0 + a.tiles_per_part, a.tiles_total);
            if (!mbar_wait(aready_bar, item_iter & 1u, abort_flag, 2)) break;
            tcgen05_fence_after();
            for (uint32_t t = t0; t < t1 && ok; ++t, ++tile_iter) {
                const uint32_t buf = tile_iter % NBUF;
                if (!tempty_ready && !mbar_wait(tempty_bar(buf), ((tile_iter / NBUF) & 1u) ^ 1u, abort_flag, 3)) { ok = false; break; }
                tempty_ready = 0;
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + Cfg::kDCol + buf * BN_;
                for (int s = 0; s < stages_per_tile; ++s, ++stage_iter) {
                    const int stage = stage_iter % NSTAGES;
                    const uint32_t phase = (stage_iter / NSTAGES) & 1u;
                    if (!full_ready && !mbar_wait(full_bar(stage), phase, abort_flag, 4)) { ok = false; break; }
                    const int kbc = min(KPS, nkb - s * KPS);
                    // descriptors of one stage differ only in the start-address field: +2 per K=16 step inside a
                    // 128-byte swizzle row, + kTileBytes/16 per K block
                    const uint64_t bdesc0 = make_b_desc(smem_base + stage * STAGE_BYTES);
                    const uint32_t a0 = tmem_base + (uint32_t)(s * KPS * (KBLK / 16) * 8);
                    if (leader) {
#pragma unroll
                        for (int k4 = 0; k4 < KBLK / 16; ++k4)
                            umma_ts_bf16(d_tmem, a0 + (uint32_t)(k4 * 8), bdesc0 + (uint64_t)(k4 * 2), Cfg::kIdesc,
                                         (k4 == 0) ? (s > 0 ? 1u : 0u) : 1u);
                    }
                    // probe what the next iteration will need while the pipe is busy
                    {
                        const uint32_t nsi = stage_iter + 1;
                        full_ready = mbar_try_wait(full_bar(nsi % NSTAGES), (nsi / NSTAGES) & 1u);
                        if (s == stages_per_tile - 1) {
                            const uint32_t nti = tile_iter + 1;
                            tempty_ready = mbar_try_wait(tempty_bar(nti % NBUF), ((nti / NBUF) & 1u) ^ 1u);
                        }
                    }
                    if (leader) {
#pragma unroll
                        for (int kb = 1; kb < KPS; ++kb) {
                            if (kb < kbc) {
#pragma unroll
                                for (int k4 = 0; k4 < KBLK / 16; ++k4)
                                    umma_ts_bf16(d_tmem, a0 + (uint32_t)((kb * (KBLK / 16) + k4) * 8),
                                                 bdesc0 + (uint64_t)(kb * (Cfg::kTileBytes >> 4) + k4 * 2), Cfg::kIdesc, 1u);
                            }
                        }
                        // frees the stage (in every CTA that multicast into it) when these MMAs have read it
                        if (CLUSTER == 1) umma_commit(empty_bar(stage));
                        else umma_commit_mcast(empty_bar(stage), kMask);
                    }
                    __syncwarp();
                }
                if (ok && leader) umma_commit(tfull_bar(buf));  // accumulator tile complete
                __syncwarp();
            }
        }
    } else {
        // ===================== epilogue: lane == query =====================
        const int quad = warp & 3;                    // TMEM lane quadrant this warp may access
        const int ql = quad * 32 + lane;              // query within the tile
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
        uint32_t tile_iter = 0;
        bool ok = true;
        for (int r = 0; r < n_rounds && ok; ++r) {
            const uint32_t part = (uint32_t)slot + (uint32_t)r * (uint32_t)a.n_slots, mt = (uint32_t)mgroup * CLUSTER + crank;
            if (part >= (uint32_t)a.P) break;
            const uint32_t t0 = part * a.tiles_per_part;
            const uint32_t t1 = min(t0 + a.tiles_per_part, a.tiles_total);
            const uint32_t gq = mt * BM + ql;
            const bool q_valid = gq < (uint32_t)a.nq;
            // A operand: this thread's query row, bf16 pairs, into TMEM columns [0, Dp/2)
            {
                const uint4* src = reinterpret_cast<const uint4*>(a.qb + (size_t)gq * a.Dp);
                for (int c = 0; c < a.Dp / 32; ++c) {
                    uint32_t w[16];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint4 x = __ldg(src + c * 4 + i);
                        w[4 * i + 0] = x.x; w[4 * i + 1] = x.y; w[4 * i + 2] = x.z; w[4 * i + 3] = x.w;
                    }
                    tmem_st_32x32b_x16(lane_addr + c * 16, w);
                }
                tmem_st_wait();
            }
            for (int j = 0; j < KP; ++j) {
                l_score[j * BM + ql] = -INFINITY;
                l_row[j * BM + ql] = ROW_NONE;
            }
            // thr_l: worst score kept in this thread's list.  thr_g: best such value any partition of this query has
            // published; a row at or below it can never be needed, so it also gates the list.
            float thr_l = q_valid ? -INFINITY : INFINITY;
            float thr_g = -INFINITY, thr_pub = -INFINITY;
            int min_pos = 0;
            uint32_t* gthr = a.gthr + (q_valid ? gq : 0);
            uint32_t g_bits = 0u;
            tcgen05_fence_before();
            mbar_arrive(aready_bar);
            const uint32_t row_end = a.n_rows;
            for (uint32_t t = t0; t < t1; ++t, ++tile_iter) {
                const uint32_t buf = tile_iter % NBUF;
                // refresh the shared floor every 8th tile; the load issued now is consumed 8 tiles later, so its
                // (loaded) L2 latency never sits on the per-tile critical path
                if (a.share_floor && (tile_iter & 7u) == 0u) {
                    if (g_bits != 0u) thr_g = fmaxf(thr_g, f32_from_orderable(g_bits));
                    g_bits = *reinterpret_cast<volatile uint32_t*>(gthr);
                }
                if (!mbar_wait(tfull_bar(buf), (tile_iter / NBUF) & 1u, abort_flag, 5)) { ok = false; break; }
                tcgen05_fence_after();
                if (a.debug_mode & 2) {
                    tcgen05_fence_before();
                    mbar_arrive(tempty_bar(buf));
                    continue;
                }
#pragma unroll
                for (int h = 0; h < BN_ / 64; ++h) {
                    uint32_t v[64];
                    tmem_ld_32x32b_x32(lane_addr + Cfg::kDCol + buf * BN_ + h * 64, v);
                    tmem_ld_32x32b_x32(lane_addr + Cfg::kDCol + buf * BN_ + h * 64 + 32, v + 32);
                    tmem_ld_wait();
                    if (h == BN_ / 64 - 1) {
                        tcgen05_fence_before();
                        mbar_arrive(tempty_bar(buf));  // accumulator is in registers: the MMA warp may reuse the buffer
                    }
                    const uint32_t row0 = t * BN_ + h * 64;
                    if (a.dump != nullptr) {
                        float* drow = a.dump + (size_t)gq * ((size_t)a.tiles_total * BN_) + row0;
#pragma unroll
                        for (int i = 0; i < 64; ++i) drow[i] = __uint_as_float(v[i]);
                    }
                    float thr = fmaxf(thr_l, thr_g);
                    if (a.debug_mode & 4) thr = INFINITY;
                    bool any = false;
#pragma unroll
                    for (int i = 0; i < 64; ++i) any |= (__uint_as_float(v[i]) > thr);
                    if (any) {
#pragma unroll
                        for (int i = 0; i < 64; ++i) {
                            const float sc = __uint_as_float(v[i]);
                            if (sc > thr && row0 + i < row_end) {
                                l_score[min_pos * BM + ql] = sc;
                                l_row[min_pos * BM + ql] = row0 + i;
                                float mn = INFINITY;
                                for (int j = 0; j < KP; ++j) {
                                    const float x = l_score[j * BM + ql];
                                    if (x < mn) { mn = x; min_pos = j; }
                                }
                                thr_l = mn;
                                thr = fmaxf(thr_l, thr_g);
                            }
                        }
                        if (a.share_floor && q_valid && thr_l > thr_pub && thr_l > thr_g) {  // list is full and its floor rose: publish
                            atomicMax(gthr, f32_orderable(thr_l));
                            thr_pub = thr_l;
                        }
                    }
                }
            }
            if (ok && q_valid) {
                const size_t o = ((size_t)gq * a.P + part) * KP;
                for (int j = 0; j < KP; ++j) {
                    a.cand_score[o + j] = l_score[j * BM + ql];
                    a.cand_row[o + j] = l_row[j * BM + ql];
                }
                // every row this thread dropped scored <= max(thr_l, thr_g) at the time, and both only grow
                a.cand_thr[(size_t)gq * a.P + part] = fmaxf(thr_l, thr_g);
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (CLUSTER > 1) cluster_sync_all();  // no CTA leaves while a peer may still multicast into it
    tcgen05_fence_after();
    if (threadIdx.x == 0 && *abort_flag) atomicMax(a.error_flag, *abort_flag);
    if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---- shadow / query preparation --------------------------------------------------------------------------------
enum ShadowKind { SHADOW_IP = 0, SHADOW_COSINE = 1, SHADOW_L2 = 2 };

// One warp per row: bf16 shadow row of Dp elements (zero padded) + max row norm (for the certification bound).
//   SHADOW_IP      c
//   SHADOW_COSINE  c / |c|            (zero rows stay zero: cosine distance 1.0, simd.rs:1631-1633)
//   SHADOW_L2      [c, n1, n2, n3]    with n1+n2+n3 ~ |c|^2 split into three bf16 pieces (columns dim..dim+2)
__global__ void build_shadow_kernel(const float* __restrict__ rows, uint64_t first_row, uint64_t n, int dim, int Dp,
                                    int kind, __nv_bfloat16* __restrict__ shadow, float* __restrict__ max_norm) {
    uint64_t row = first_row + (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= first_row + n) return;
    const float* r = rows + row * dim;
    float ss = 0.0f;
    for (int d = lane; d < dim; d += 32) {
        float x = __ldg(r + d);
        ss = fmaf(x, x, ss);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    float norm = sqrtf(ss);
    float scale = 1.0f;
    if (kind == SHADOW_COSINE) scale = norm > 0.0f ? 1.0f / norm : 0.0f;
    __nv_bfloat16* out = shadow + row * (uint64_t)Dp;
    for (int d = lane; d < Dp; d += 32) {
        float x = d < dim ? __ldg(r + d) * scale : 0.0f;
        if (kind == SHADOW_L2 && d >= dim && d < dim + 3) {
            float n1 = __bfloat162float(__float2bfloat16_rn(ss));
            float n2 = __bfloat162float(__float2bfloat16_rn(ss - n1));
            float n3 = (ss - n1) - n2;
            x = d == dim ? n1 : (d == dim + 1 ? n2 : n3);
        }
        out[d] = __float2bfloat16_rn(x);
    }
    if (lane == 0 && isfinite(norm)) atomicMax(reinterpret_cast<unsigned int*>(max_norm), __float_as_uint(norm));
}

// Queries: bf16 A operand rows (zero padded to n_mtiles*128 x Dp) + |q| per query.
//   SHADOW_IP      q                 SHADOW_COSINE  q / |q|          SHADOW_L2  [2q, -1, -1, -1]
__global__ void prepare_queries_kernel(const float* __restrict__ queries, int nq, int nq_pad, int dim, int Dp, int kind,
                                       __nv_bfloat16* __restrict__ qb, float* __restrict__ qnorm) {
    int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (q >= nq_pad) return;
    __nv_bfloat16* out = qb + (size_t)q * Dp;
    if (q >= nq) {
        for (int d = lane; d < Dp; d += 32) out[d] = __float2bfloat16_rn(0.0f);
        return;
    }
    const float* r = queries + (size_t)q * dim;
    float ss = 0.0f;
    for (int d = lane; d < dim; d += 32) {
        float x = __ldg(r + d);
        ss = fmaf(x, x, ss);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    float norm = sqrtf(ss);
    float scale = 1.0f;
    if (kind == SHADOW_COSINE) scale = norm > 0.0f ? 1.0f / norm : 0.0f;
    if (kind == SHADOW_L2) scale = 2.0f;
    for (int d = lane; d < Dp; d += 32) {
        float x = d < dim ? __ldg(r + d) * scale : 0.0f;
        if (kind == SHADOW_L2 && d >= dim && d < dim + 3) x = -1.0f;
        out[d] = __float2bfloat16_rn(x);
    }
    if (lane == 0) qnorm[q] = norm;
}

// ---- finalize: shortlist -> exact-order rescore -> certified top-k -------------------------------------------------
struct FinArgs {
    const float* cand_score;  // [nq][P][KP]
    const uint32_t* cand_row;
    const float* cand_thr;    // [nq][P]
    int P;
    int M1;                   // pow2 >= P*KP (<= 4096)
    int R;                    // rescore budget, pow2 <= 1024, >= k
    const float* corpus;
    int dim;
    const float* queries;     // original f32 queries [nq][dim]
    const float* qnorm;       // [nq]
    const float* max_norm;    // device scalar: max row norm
    int nq, k, metric;
    float eps_rel;            // relative error bound of the coarse score
    const uint32_t* small_seg;
    int n_small;
    uint32_t* out_rows;
    float* out_dists;
    uint32_t* out_counts;
    uint32_t* uncertified;    // [nq] flag
    uint32_t* n_uncertified;  // counter
};

template <bool ASC>
__global__ void __launch_bounds__(256) finalize_kernel(FinArgs a) {
    extern __shared__ __align__(16) unsigned char smem_fin[];
    uint64_t* s = reinterpret_cast<uint64_t*>(smem_fin);            // [M1] coarse keys
    uint64_t* e = s + a.M1;                                         // [R] exact keys
    float* sq = reinterpret_cast<float*>(e + a.R);                  // [dim_pad] query
    __shared__ uint32_t sh_T;      // orderable max of partition thresholds
    __shared__ uint32_t sh_ncand;
    const int q = blockIdx.x, tid = threadIdx.x;
    const int dim = a.dim;
    const bool vec = (dim & 3) == 0;
    if (tid == 0) { sh_T = 0; sh_ncand = 0; }
    for (int d = tid; d < dim; d += blockDim.x) sq[d] = a.queries[(size_t)q * dim + d];
    __syncthreads();
    const int total = a.P * KP;
    uint32_t local_valid = 0;
    for (int i = tid; i < a.M1; i += blockDim.x) {
        uint64_t key = KEY_NONE;
        if (i < total) {
            uint32_t row = a.cand_row[(size_t)q * total + i];
            if (row != ROW_NONE) {
                key = make_key<false>(a.cand_score[(size_t)q * total + i], row);  // best coarse score first
                ++local_valid;
            }
        }
        s[i] = key;
    }
    if (local_valid) atomicAdd(&sh_ncand, local_valid);
    for (int p = tid; p < a.P; p += blockDim.x) atomicMax(&sh_T, f32_orderable(a.cand_thr[(size_t)q * a.P + p]));
    bitonic_sort_u64(s, a.M1);
    const int ncand = (int)sh_ncand;
    const int rn = min(ncand, a.R);
    float T = f32_from_orderable(sh_T);  // every row dropped inside a partition has coarse score <= T
    if (ncand > a.R) T = fmaxf(T, key_score<false>(s[a.R]));  // ... and so has every candidate cut here
    for (int i = tid; i < a.R; i += blockDim.x) {
        uint64_t key = KEY_NONE;
        if (i < rn) {
            uint32_t row = key_row(s[i]);
            const float* c = a.corpus + (size_t)row * dim;
            float v;
            if (a.metric == LB_IP) {
                bool small = a.n_small > 0 && in_small_segment(a.small_seg, a.n_small, row);
                v = small ? ip_single_order<false>(sq, c, dim, vec) : ip_batch8_order<false>(sq, c, dim, vec);
            } else if (a.metric == LB_L2) {
                v = l2_squared<false>(sq, c, dim, vec);
            } else {
                v = cosine_distance<false>(sq, c, dim, vec);
            }
            key = make_key<ASC>(v, row);
        }
        e[i] = key;
    }
    bitonic_sort_u64(e, a.R);
    const int kk = min(a.k, rn);
    for (int i = tid; i < a.k; i += blockDim.x) {
        uint32_t row = ROW_NONE;
        float score = __int_as_float(0x7fc00000);
        if (i < kk) {
            row = key_row(e[i]);
            score = key_score<ASC>(e[i]);
        }
        a.out_rows[(size_t)q * a.k + i] = row;
        a.out_dists[(size_t)q * a.k + i] = score;
    }
    if (tid == 0) {
        a.out_counts[q] = kk;
        bool certified;
        if (T == -INFINITY) {
            certified = true;  // nothing was dropped anywhere: the shortlist is the whole corpus
        } else if (kk < a.k) {
            certified = false;
        } else {
            const float worst = key_score<ASC>(e[a.k - 1]);
            const float qn = a.qnorm[q], cn = *a.max_norm;
            if (a.metric == LB_IP) {
                // dropped row: exact <= coarse + eps <= T + eps
                const float eps = a.eps_rel * qn * cn;
                certified = worst > T + eps;
            } else if (a.metric == LB_COSINE) {
                // coarse = cos of the normalised bf16 vectors; dropped row: dist >= 1 - (T + eps)
                const float eps = a.eps_rel + 4e-6f;
                certified = worst < 1.0f - (T + eps);
            } else {
                // coarse = 2 q.c - |c|^2 ; dropped row: dist >= |q|^2 - (T + eps)
                const float eps = 2.0f * a.eps_rel * qn * cn + 1e-5f * (qn * qn + cn * cn);
                certified = worst < qn * qn - (T + eps);
            }
        }
        if (!certified) {
            a.uncertified[q] = 1;
            atomicAdd(a.n_uncertified, 1u);
        } else {
            a.uncertified[q] = 0;
        }
    }
}

// ---- diagnostics: tcgen05.mma issue-rate probe ------------------------------------------------------------------
// One warp issues `iters` MMAs (M=128, K=16, bf16) round-robin over `n_acc` independent accumulators of N columns;
// operands are whatever is in shared memory / TMEM (timing only).  Reports SM cycles from first issue to last commit.
__device__ __forceinline__ void umma_ss_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

template <int N, int NACC, bool TS>
__global__ void __launch_bounds__(64, 1) mma_rate_kernel(int iters16, int commit_every16, unsigned long long* cycles_out) {
    extern __shared__ __align__(16) unsigned char smem_probe[];
    const uint32_t smem_base = (smem_u32(smem_probe) + 1023u) & ~1023u;
    unsigned char* smem = smem_probe + (smem_base - smem_u32(smem_probe));
    const uint32_t bar = smem_base + 49152;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + 49152 + 32);
    for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    const uint32_t bar2 = bar + 8;  // receives the intermediate commits; nobody waits on it
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(bar2, 1);
        fence_barrier_init();
    }
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        __syncwarp();
        tmem_alloc(smem_u32(tmem_ptr_smem), TMEM_COLS);
        tmem_relinquish();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    if (warp == 0) {
        const bool leader = elect_one();
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        const uint64_t a_desc = make_b_desc(smem_base);           // 128 rows x 64 bf16, SW128
        const uint64_t b_desc = make_b_desc(smem_base + 16384);   // up to 256 rows x 64 bf16
        constexpr uint32_t d_col0 = TMEM_COLS - NACC * N;
        long long t0 = clock64(), t1 = t0;
        if (leader) {
            for (int it = 0; it < iters16; ++it) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const uint32_t d = tmem_base + d_col0 + (uint32_t)((j % NACC) * N);
                    const uint32_t acc = (j < NACC) ? (it > 0 ? 1u : 0u) : 1u;
                    if (TS)
                        umma_ts_bf16(d, tmem_base + (uint32_t)(j * 8), b_desc + (uint64_t)((j & 3) * 2), idesc, acc);
                    else
                        umma_ss_bf16(d, a_desc + (uint64_t)((j & 3) * 2), b_desc + (uint64_t)((j & 3) * 2), idesc, acc);
                }
                if (commit_every16 > 0 && (it + 1) % commit_every16 == 0) umma_commit(bar2);
            }
            umma_commit(bar);
            t1 = clock64();
        }
        __syncwarp();
        while (!mbar_try_wait(bar, 0)) {
        }
        long long t2 = clock64();
        if (leader) {
            cycles_out[2 * blockIdx.x] = (unsigned long long)(t2 - t0);
            cycles_out[2 * blockIdx.x + 1] = (unsigned long long)(t1 - t0);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace tc
}  // namespace lb
