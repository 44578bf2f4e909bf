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
// Shadow layout in HBM (a derived structure, so it is stored the way the tensor core wants to read it): rows are
// grouped in tiles of 64, a tile is cut in K blocks of 64 bf16, and every (tile, K block) is two 4 KiB half blocks
// of 32 rows x 128 B whose 16-byte chunks are already permuted with the 128-byte shared-memory swizzle
// (chunk c of row r sits at chunk c ^ (r & 7)).  A pipeline stage is therefore ONE contiguous TMA box
// (SWIZZLE_NONE) instead of 4 boxes of 32-64 strided rows, DRAM pages are read front to back, and the bytes
// land in shared memory exactly as a SWIZZLE_128B K-major UMMA descriptor expects them.
//   byte offset of element (row, d):  (((row/64) * NKB + d/64) * 2 + (row%64)/32) * 4096
//                                     + (row%32) * 128 + ((((d%64)/8) ^ (row & 7)) << 4) + (d%8) * 2
//
// Kernels: lb_tc1.cuh (one CTA per 128 queries; used when a batch has a single query tile) and lb_tc2.cuh
// (CTA pairs, tcgen05 cta_group::2, for two or more query tiles).  Both keep the A operand (the 128 queries
// of the CTA, bf16) in TMEM columns [0, Dp/2), stream 64-row corpus tiles through a shared-memory ring, double
// buffer the f32 accumulators in TMEM columns [384, 512), and run a lane == query epilogue that keeps a private
// KP-entry shortlist per (query, row partition).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include <type_traits>

#include "lb_metrics.cuh"
#include "lb_scan.cuh"

namespace lb {
namespace tc {

constexpr int BM = 128;           // queries per work item == TMEM lanes
constexpr int BN = 64;            // corpus rows per accumulator tile
constexpr int KBLK = 64;          // bf16 elements per 128-byte swizzle row
constexpr int KP = 16;            // shortlist entries kept per (partition, query)
constexpr int MAX_DP = 768;       // padded dim limit: A occupies Dp/2 <= 384 TMEM columns
constexpr int HALF_BLOCK_BYTES = 4096;  // 32 rows x one K block, swizzled
constexpr int KPS = 4;                  // K blocks per pipeline stage
constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = 512;
constexpr int DCOL = TMEM_COLS - 2 * BN;  // accumulators: two buffers of BN columns at [384, 512)
constexpr uint32_t SMEM_RING_BYTES = 196608;                          // staging ring of either kernel
constexpr uint32_t SMEM_BAR_OFF = SMEM_RING_BYTES;
constexpr uint32_t SMEM_BYTES = SMEM_BAR_OFF + 256 + 1024;            // + barriers + alignment slack

struct TcArgs {
    const __nv_bfloat16* qb;  // [n_mtiles*128][Dp] bf16 queries, zero padded
    int nq;
    int n_mtiles;             // query tiles of 128; qb is padded to a multiple of the cluster size tiles
    int Dp;                   // padded dim, multiple of 64, <= MAX_DP
    int rem_kb;               // (Dp / 64) % KPS: K blocks of the last, partial stage of a tile (0 = none; uses tmap_rem)
    uint32_t n_rows;
    uint32_t tiles_total;     // ceil(n_rows / 64)
    uint32_t tiles_per_part;
    int P;                    // row partitions
    int lists_per_part;       // shortlists per (query, partition): 1, or 2 when the pair kernel runs two epilogue sets
    const uint64_t* allow_bits;  // optional row filter (bit r = row r allowed, LSB-first u64 words): disallowed rows never
                                 // enter a shortlist, so floors, certification and result are those of the allowed rows
    float* cand_score;        // [nq][P * lists_per_part][KP]
    uint32_t* cand_row;       // [nq][P * lists_per_part][KP]
    float* cand_thr;          // [nq][P * lists_per_part]
    int share_floor;          // 1: partitions of a query share a shortlist floor through gthr; 2: gthr holds a floor seeded by a
                              // pre-pass over a sample of the corpus and is only read (large k)
    int floor_group;          // m: a published floor is the minimum of the floors of m consecutive partitions, so that at
                              // least m * KP rows score above it (m = 1 when k <= KP - 4, else ~10 k / KP)
    float* gfloor;            // [nq][P] floors of the single partitions (initialised to -inf; used when m > 1)
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
    int sample_tiles;         // > 0: every CTA first scans tiles [0, sample_tiles) only to warm up its shortlist floors
                              // (round -1, nothing recorded), so the real partitions never start with an open gate
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
// 4-D box of the tiled shadow: coordinates (0, 0, half, K block index)
__device__ __forceinline__ void tma_load_4d(uint32_t smem_dst, const CUtensorMap* tmap, int c2, int c3, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(smem_dst), "l"(tmap), "r"(0), "r"(0), "r"(c2), "r"(c3), "r"(bar)
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

// ---- pieces shared by the two coarse kernels ------------------------------------------------------------------
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// A operand: the calling thread's query row (bf16 pairs) into TMEM columns [0, Dp/2) of its lane.
__device__ __forceinline__ void load_query_to_tmem(const __nv_bfloat16* qrow, int Dp, uint32_t lane_addr) {
    const uint4* src = reinterpret_cast<const uint4*>(qrow);
    for (int c = 0; c < Dp / 32; ++c) {
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

__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// Per-thread shortlist of one (query, row partition): the KP best coarse scores, held in REGISTERS (every index is
// a compile-time constant) and gated by a register threshold — shared-memory lists stall for thousands of cycles
// behind the tensor core's operand reads and the TMA writes.  lmin = worst score kept; thr_g = best floor any
// partition of the query has published through gthr (a row at or below it is outside the global top KP).
struct Shortlist {
    float sc[KP];
    uint32_t rw[KP];
    float lmin, thr_g, thr_pub;
    uint32_t g_bits;
    uint32_t* gthr;
    float* gfloor_q;     // floors of this query's partitions (floor groups)
    int group_m, n_parts, part;
    bool q_valid, share, may_publish;

    __device__ __forceinline__ void set_groups(float* gfloor_row, int m, int P) {
        gfloor_q = gfloor_row;
        group_m = m;
        n_parts = P;
    }
    // The list is full and its floor rose: make it visible to the other partitions of the query.  With floor groups
    // the value that may gate everybody is the smallest floor of the m partitions of my group (every member has KP
    // rows above its own floor, hence m * KP rows above the minimum); incomplete groups publish nothing.
    __device__ __forceinline__ void publish() {
        if (group_m <= 1) {
            atomicMax(gthr, f32_orderable(lmin));
        } else {
            __stcg(gfloor_q + part, lmin);
            const int g0 = (part / group_m) * group_m, g1 = g0 + group_m;
            if (g1 <= n_parts) {
                float mn = lmin;
                int p = g0;
                for (; p + 8 <= g1; p += 8) {  // independent loads: one L2 round trip per eight floors
                    float f[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) f[i] = __ldcg(gfloor_q + p + i);
#pragma unroll
                    for (int i = 0; i < 8; ++i) mn = fminf(mn, f[i]);
                }
                for (; p < g1; ++p) mn = fminf(mn, __ldcg(gfloor_q + p));
                if (mn > -INFINITY) atomicMax(gthr, f32_orderable(mn));
            }
        }
        thr_pub = lmin;
    }

    __device__ __forceinline__ void reset(bool valid, int share_floor, uint32_t* gthr_q) {
#pragma unroll
        for (int j = 0; j < KP; ++j) {
            sc[j] = -INFINITY;
            rw[j] = ROW_NONE;
        }
        q_valid = valid;
        share = share_floor != 0;
        may_publish = share_floor == 1;
        gthr = gthr_q;
        if (share_floor == 2 && valid) {  // seeded floor: available from the first tile
            const uint32_t bits = *reinterpret_cast<volatile uint32_t*>(gthr_q);
            if (bits != 0u) thr_g = fmaxf(thr_g, f32_from_orderable(bits));
        }
        lmin = -INFINITY;
        thr_pub = -INFINITY;
    }
    __device__ __forceinline__ void init_floor() {
        thr_g = -INFINITY;
        g_bits = 0u;
    }
    __device__ __forceinline__ float gate() const { return q_valid ? fmaxf(lmin, thr_g) : INFINITY; }
    // end of the warm-up round: the KP-th best score of the sample is a valid floor for every partition (KP rows of
    // the corpus score at least that much), so it becomes the shared floor and the list starts over
    __device__ __forceinline__ void absorb_sample() {
        if (may_publish && q_valid && group_m <= 1 && lmin > thr_g) {
            thr_g = lmin;
            atomicMax(gthr, f32_orderable(lmin));
        }
    }
    // refresh the shared floor every 8th tile; the load issued now is consumed 8 tiles later, so its (loaded) L2
    // latency never sits on the per-tile critical path
    __device__ __forceinline__ void poll_floor(uint32_t tile_iter) {
        if (share && (tile_iter & 7u) == 0u) {
            if (g_bits != 0u) thr_g = fmaxf(thr_g, f32_from_orderable(g_bits));
            g_bits = *reinterpret_cast<volatile uint32_t*>(gthr);
            if (may_publish && group_m > 1 && q_valid && lmin > thr_pub) publish();
        }
    }
    // replace the current minimum by (score, row) and recompute the minimum: ~70 ALU instructions, no memory
    __device__ __forceinline__ void insert(float score, uint32_t row) {
        bool done = false;
#pragma unroll
        for (int j = 0; j < KP; ++j) {
            const bool hit = !done && sc[j] == lmin;
            sc[j] = hit ? score : sc[j];
            rw[j] = hit ? row : rw[j];
            done = done || hit;
        }
        float m0 = fmin3(sc[0], sc[1], sc[2]), m1 = fmin3(sc[3], sc[4], sc[5]);
        m0 = fmin3(m0, sc[6], sc[7]);
        m1 = fmin3(m1, sc[8], sc[9]);
        m0 = fmin3(m0, sc[10], sc[11]);
        m1 = fmin3(m1, sc[12], sc[13]);
        lmin = fmin3(fminf(m0, m1), sc[14], sc[15]);
    }
    // 64 accumulator columns of this thread's query = rows row0 .. row0+63.  Called by whole warps.
    __device__ __forceinline__ void scan64(const uint32_t* v, uint32_t row0, uint32_t row_end, bool disabled,
                                           const uint64_t* __restrict__ allow = nullptr) {
        float thr = disabled ? INFINITY : gate();
        // tile maximum with 3-input max: 32 instructions for 64 scores
        float m0 = fmax3(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]));
        float m1 = fmax3(__uint_as_float(v[3]), __uint_as_float(v[4]), __uint_as_float(v[5]));
#pragma unroll
        for (int i = 6; i + 3 < 64; i += 4) {
            m0 = fmax3(m0, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
            m1 = fmax3(m1, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
        }
        m0 = fmax3(m0, __uint_as_float(v[62]), __uint_as_float(v[63]));
        // Slow path, entered by the whole warp when any lane has a candidate.  It must stay SMALL: a fully unrolled
        // "for each of the 64 scores: compare, insert" is ~90 KB of code whose sparse execution misses the instruction
        // cache at every step (~7700 cycles per tile measured).  So: (1) a 64-bit hit mask per lane from straight
        // compares, (2) a rolled loop that pops each lane's lowest hit and fetches the score with a select tree.
        if (__any_sync(0xffffffffu, fmaxf(m0, m1) > thr)) {
            uint32_t mlo = 0u, mhi = 0u;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                mlo |= (__uint_as_float(v[i]) > thr ? 1u : 0u) << i;
                mhi |= (__uint_as_float(v[32 + i]) > thr ? 1u : 0u) << i;
            }
            if (row0 + 64u > row_end) {  // last tile: rows past the end of the corpus never enter a list
                const uint32_t n_ok = row_end > row0 ? row_end - row0 : 0u;
                mlo &= n_ok >= 32u ? 0xffffffffu : ((1u << n_ok) - 1u);
                mhi &= n_ok >= 64u ? 0xffffffffu : (n_ok > 32u ? ((1u << (n_ok - 32u)) - 1u) : 0u);
            }
            bool filtered = false;  // this lane lost a hit to the row filter: its tile maximum may be a disallowed row
            if (allow != nullptr) {
                const uint64_t w = row0 < row_end ? __ldg(allow + (row0 >> 6)) : 0ull;  // row0 is a multiple of 64
                const uint32_t alo = (uint32_t)w, ahi = (uint32_t)(w >> 32);
                filtered = ((mlo & ~alo) | (mhi & ~ahi)) != 0u;
                mlo &= alo;
                mhi &= ahi;
            }
            // a lane with exactly one hit (the usual case) already holds its score: it is the tile maximum
            const bool one_hit = !filtered && __popc(mlo) + __popc(mhi) == 1;
            if (one_hit) {
                const int idx = mlo != 0u ? __ffs((int)mlo) - 1 : 32 + __ffs((int)mhi) - 1;
                insert(fmaxf(m0, m1), row0 + (uint32_t)idx);
                mlo = 0u;
                mhi = 0u;
            }
#pragma unroll 1
            while (__any_sync(0xffffffffu, (mlo | mhi) != 0u)) {
                if ((mlo | mhi) != 0u) {
                    int idx;
                    if (mlo != 0u) {
                        idx = __ffs((int)mlo) - 1;
                        mlo &= mlo - 1u;
                    } else {
                        idx = 32 + __ffs((int)mhi) - 1;
                        mhi &= mhi - 1u;
                    }
                    // v[idx] with a run-time idx: 6-level select tree over the register array (63 selects)
                    uint32_t t5[32], t4[16], t3[8], t2[4], t1[2];
#pragma unroll
                    for (int j = 0; j < 32; ++j) t5[j] = (idx & 32) ? v[32 + j] : v[j];
#pragma unroll
                    for (int j = 0; j < 16; ++j) t4[j] = (idx & 16) ? t5[16 + j] : t5[j];
#pragma unroll
                    for (int j = 0; j < 8; ++j) t3[j] = (idx & 8) ? t4[8 + j] : t4[j];
#pragma unroll
                    for (int j = 0; j < 4; ++j) t2[j] = (idx & 4) ? t3[4 + j] : t3[j];
#pragma unroll
                    for (int j = 0; j < 2; ++j) t1[j] = (idx & 2) ? t2[2 + j] : t2[j];
                    const float x = __uint_as_float((idx & 1) ? t1[1] : t1[0]);
                    if (x > gate()) insert(x, row0 + (uint32_t)idx);  // the gate may have risen since the mask was taken
                }
            }
            // floor groups publish from poll_floor (every 8th tile): the group minimum costs m loads
            if (may_publish && q_valid && group_m <= 1 && lmin > thr_pub && lmin > thr_g) publish();
        }
    }
    // `sub`: which of the partition's lists_per_part shortlists this is (the pair kernel with two epilogue sets)
    __device__ __forceinline__ void flush(const TcArgs& a, uint32_t gq, uint32_t part, uint32_t sub = 0) {
        if (!q_valid) return;
        const size_t list = (size_t)gq * ((size_t)a.P * a.lists_per_part) + (size_t)part * a.lists_per_part + sub;
        const size_t o = list * KP;
#pragma unroll
        for (int j = 0; j < KP; j += 4) {
            *reinterpret_cast<float4*>(a.cand_score + o + j) = make_float4(sc[j], sc[j + 1], sc[j + 2], sc[j + 3]);
            *reinterpret_cast<uint4*>(a.cand_row + o + j) = make_uint4(rw[j], rw[j + 1], rw[j + 2], rw[j + 3]);
        }
        // every row this thread dropped scored <= max(lmin, thr_g) at the time, and both only grow
        a.cand_thr[list] = fmaxf(lmin, thr_g);
    }
};

// ---- shadow / query preparation --------------------------------------------------------------------------------
enum ShadowKind { SHADOW_IP = 0, SHADOW_COSINE = 1, SHADOW_L2 = 2 };

// One warp per row: the row's Dp bf16 values (zero padded) written into the tiled, pre-swizzled layout described
// at the top of this file, + max row norm (for the certification bound).
//   SHADOW_IP      c
//   SHADOW_COSINE  c / |c|            (zero rows stay zero: cosine distance 1.0, simd.rs:1631-1633)
//   SHADOW_L2      [c, n1, n2, n3]    with n1+n2+n3 ~ |c|^2 split into three bf16 pieces (columns dim..dim+2)
__host__ __device__ inline size_t shadow_chunk_offset(uint64_t row, int chunk /* 16-byte chunk = 8 elements */, int nkb) {
    const uint64_t tile = row >> 6;
    const uint32_t half = (uint32_t)(row >> 5) & 1u, r = (uint32_t)row & 31u;
    const uint32_t kb = (uint32_t)chunk >> 3, c = (uint32_t)chunk & 7u;
    return (size_t)(((tile * (uint64_t)nkb + kb) * 2 + half) * HALF_BLOCK_BYTES) + (size_t)r * 128 + (size_t)((c ^ (r & 7u)) << 4);
}
static __global__ void build_shadow_kernel(const float* __restrict__ rows, uint64_t first_row, uint64_t n, int dim, int Dp,
                                    int kind, unsigned char* __restrict__ shadow, float* __restrict__ max_norm) {
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
    const int nkb = Dp / KBLK;
    for (int chunk = lane; chunk < Dp / 8; chunk += 32) {
        __align__(16) __nv_bfloat16 v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int d = chunk * 8 + e;
            float x = d < dim ? __ldg(r + d) * scale : 0.0f;
            if (kind == SHADOW_L2 && d >= dim && d < dim + 3) {
                float n1 = __bfloat162float(__float2bfloat16_rn(ss));
                float n2 = __bfloat162float(__float2bfloat16_rn(ss - n1));
                float n3 = (ss - n1) - n2;
                x = d == dim ? n1 : (d == dim + 1 ? n2 : n3);
            }
            v[e] = __float2bfloat16_rn(x);
        }
        *reinterpret_cast<uint4*>(shadow + shadow_chunk_offset(row, chunk, nkb)) = *reinterpret_cast<const uint4*>(v);
    }
    if (lane == 0 && isfinite(norm)) atomicMax(reinterpret_cast<unsigned int*>(max_norm), __float_as_uint(norm));
}

// Queries: bf16 A operand rows (zero padded to n_mtiles*128 x Dp) + |q| per query.
//   SHADOW_IP      q                 SHADOW_COSINE  q / |q|          SHADOW_L2  [2q, -1, -1, -1]
static __global__ void prepare_queries_kernel(const float* __restrict__ queries, int nq, int nq_pad, int dim, int Dp, int kind,
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

// ---- seeded floors for large k ---------------------------------------------------------------------------------
// After a pre-pass over a sample of the corpus: the r-th best coarse score of the sample, per query, becomes the
// floor the main pass starts with (gthr).  It is a heuristic gate — about r * (rows / sample rows) rows of the corpus
// score above it — and the certification of finalize_kernel is what keeps the result exact: a floor that turns out
// too high leaves the query uncertified and it is re-run by the exact scan.
static __global__ void __launch_bounds__(256) seed_floor_kernel(const float* __restrict__ cand_score, const uint32_t* __restrict__ cand_row, int P,
                                                         int M, int r, uint32_t* __restrict__ gthr) {
    extern __shared__ __align__(16) unsigned char smem_seed[];
    uint64_t* s = reinterpret_cast<uint64_t*>(smem_seed);
    const int q = blockIdx.x, total = P * KP;
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
        uint64_t key = KEY_NONE;
        if (i < total && cand_row[(size_t)q * total + i] != ROW_NONE) key = make_key<false>(cand_score[(size_t)q * total + i], (uint32_t)i);
        s[i] = key;
    }
    bitonic_sort_u64(s, M);
    if (threadIdx.x == 0) gthr[q] = (r >= 1 && r <= M && s[r - 1] != KEY_NONE) ? f32_orderable(key_score<false>(s[r - 1])) : 0u;
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

// Exact-order rescore of one row by EIGHT threads: thread i owns AVX lane i (elements i, i+8, i+16, ...), so every
// accumulator still sees its products in the reference's order; the horizontal sums are the reference's trees done
// with shuffles.  The row comes from `rowbuf` (shared memory, filled with coalesced loads by the same warp), 256
// columns at a time.  lanes: lane & 7 = AVX lane, lane >> 3 = which of the warp's four rows.
constexpr int FIN_COLS = 256;   // columns staged per round
__device__ __forceinline__ float hsum8_shfl(float a) {   // simd.rs:1427-1436 over the 8 threads of a row
    a = a + __shfl_xor_sync(0xffffffffu, a, 4);           // s[i] = acc[i] + acc[i+4]
    a = a + __shfl_xor_sync(0xffffffffu, a, 1);           // t0 = s0 + s1, t2 = s2 + s3
    return a + __shfl_xor_sync(0xffffffffu, a, 2);        // t0 + t2
}
__device__ __forceinline__ float rescore_row_lanes(int metric, bool two_acc_ip, const float* __restrict__ sq, const float* __restrict__ c,
                                                   int dim, float* rowbuf /* this row's FIN_COLS floats */, bool row_ok, int lane) {
    const int i = lane & 7, sub = lane & 7;
    const int chunks = dim >> 3;
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
    for (int col0 = 0; col0 < chunks * 8; col0 += FIN_COLS) {
        const int ncol = min(FIN_COLS, chunks * 8 - col0);
        __syncwarp();
        // the row's 8 threads fetch its columns [col0, col0 + ncol): 8 x 16 B = 128 contiguous bytes per step
        if (row_ok)
            for (int x = sub * 4; x < ncol; x += 32) *reinterpret_cast<float4*>(rowbuf + x) = __ldg(reinterpret_cast<const float4*>(c + col0 + x));
        __syncwarp();
        if (row_ok) {
            const int j0 = col0 >> 3, nj = ncol >> 3;
#pragma unroll 4
            for (int jj = 0; jj < nj; ++jj) {
                const float qv = sq[col0 + 8 * jj + i], cv = rowbuf[8 * jj + i];
                const bool odd = ((j0 + jj) & 1) != 0;
                if (metric == LB_IP) {
                    if (two_acc_ip && odd) a1 = fmaf(qv, cv, a1);
                    else a0 = fmaf(qv, cv, a0);
                } else if (metric == LB_L2) {
                    const float d = qv - cv;
                    if (odd) a1 = fmaf(d, d, a1);
                    else a0 = fmaf(d, d, a0);
                } else {
                    a0 = fmaf(qv, cv, a0);
                    a1 = fmaf(qv, qv, a1);
                    a2 = fmaf(cv, cv, a2);
                }
            }
        }
    }
    const int t0 = chunks * 8;
    if (metric == LB_IP) {
        if (two_acc_ip) a0 = a0 + a1;
        float out = hsum8_shfl(a0);
        for (int x = t0; x < dim; ++x) out = out + sq[x] * __ldg(c + (row_ok ? x : 0));
        return out;
    } else if (metric == LB_L2) {
        a0 = a0 + a1;
        float sum = hsum8_shfl(a0);
        for (int x = t0; x < dim; ++x) {
            const float diff = sq[x] - __ldg(c + (row_ok ? x : 0));
            sum = sum + diff * diff;
        }
        return sum;
    }
    float dot = hsum8_shfl(a0), na = hsum8_shfl(a1), nb = hsum8_shfl(a2);
    for (int x = t0; x < dim; ++x) {
        const float qa = sq[x], cb = __ldg(c + (row_ok ? x : 0));
        dot = dot + qa * cb;
        na = na + qa * qa;
        nb = nb + cb * cb;
    }
    const float denom = sqrtf(na * nb);
    if (denom < 1e-30f) return 1.0f;
    return 1.0f - dot / denom;
}

template <bool ASC>
__global__ void __launch_bounds__(1024) finalize_kernel(FinArgs a) {
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
    if (vec) {
        // eight threads per row, four rows per warp, blockDim/8 rows per pass of the block (rows are independent:
        // warp-local syncs only)
        float* rowbuf = sq + ((dim + 3) & ~3) + (tid >> 3) * FIN_COLS;
        const int lane = tid & 31;
        const int rows_per_pass = (int)(blockDim.x >> 3);
        for (int base = 0; base < a.R; base += rows_per_pass) {
            const int i = base + (tid >> 3);
            const bool row_ok = i < rn;
            const uint32_t row = row_ok ? key_row(s[i]) : 0u;
            const float* c = a.corpus + (size_t)row * dim;
            const bool two = a.metric == LB_IP && row_ok && a.n_small > 0 && in_small_segment(a.small_seg, a.n_small, row);
            const float v = rescore_row_lanes(a.metric, two, sq, c, dim, rowbuf, row_ok, lane);
            if ((tid & 7) == 0 && i < a.R) e[i] = row_ok ? make_key<ASC>(v, row) : KEY_NONE;
        }
    } else {
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
