// lb_tc.cuh — tensor-core coarse pass (tcgen05 / TMEM / TMA) + exact-order finalize.
//
// The dense metrics (IP, and through it cosine and L2) are a Q x C^T contraction
// (reference hot loop: src/storage/flat_mmap.rs:2179-2256 ip_scan_chunk_topk over
// simd::inner_product_batch8_f32, src/distance/simd.rs:1450-1525); so are the binary
// metrics once the bits are bytes (|a & b| = a . b over {0,1}; packed_binary_search,
// flat_mmap.rs:1345-1409).  On B200 that contraction runs on the 5th-generation tensor
// cores over a narrow shadow of the corpus — 8-bit operands (tcgen05 kind::i8, exact
// s32 accumulators) for IP / cosine / the binary metrics, bf16 (kind::f16) for L2 and
// for corpora the 8-bit quantisation does not suit —; a per-query shortlist (or, for
// large k, every row above a seeded floor) is kept in the accumulator epilogue and
// re-scored in f32 in the reference's exact summation order, so the returned ids /
// order / scores are the reference's, not the shadow's.  A shortlist is only accepted
// when a bound built from MEASURED norms proves no dropped row could enter the top-k
// (finalize_kernel::certify); otherwise the query is re-run by the exact scan of
// lb_scan.cuh.
//
// Shadow layout in HBM (a derived structure, so it is stored the way the tensor core wants to read it): rows are
// grouped in tiles of 64, a tile is cut in K blocks of 128 operand bytes (64 bf16 or 128 8-bit elements), and every (tile, K block) is two 4 KiB half blocks
// of 32 rows x 128 B whose 16-byte chunks are already permuted with the 128-byte shared-memory swizzle
// (chunk c of row r sits at chunk c ^ (r & 7)).  A pipeline stage is therefore ONE contiguous TMA box
// (SWIZZLE_NONE) instead of 4 boxes of 32-64 strided rows, DRAM pages are read front to back, and the bytes
// land in shared memory exactly as a SWIZZLE_128B K-major UMMA descriptor expects them.
//   byte offset of bf16 element (row, d):  (((row/64) * NKB + d/64) * 2 + (row%64)/32) * 4096
//                                          + (row%32) * 128 + ((((d%64)/8) ^ (row & 7)) << 4) + (d%8) * 2
//   (8-bit operands: the same with 128 elements per K block and 16 per chunk — shadow_chunk_offset)
//
// Kernels: lb_tc1.cuh (one CTA per 128 queries; used when a batch has a single query tile) and lb_tc2.cuh
// (CTA pairs, tcgen05 cta_group::2, for two or more query tiles).  Both keep the A operand (the 128 queries
// of the CTA) in TMEM columns [0, row bytes / 4), stream corpus tiles through a shared-memory ring, keep two or three
// 32-bit accumulator tiles in the TMEM columns behind it, and run a lane == query epilogue that keeps a private
// KP-entry shortlist per (query, row partition) or appends the rows above a seeded floor to a hit region.
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
constexpr uint32_t SMEM_SCRATCH_OFF = SMEM_BAR_OFF + 256;               // per-epilogue-warp scratch (side values of a tile)
constexpr uint32_t SMEM_SCRATCH_BYTES = 8 * 3 * 128 * 4;                // 8 epilogue warps x EPI_SCRATCH_SLOTS x EPI_SCRATCH_WORDS words
constexpr uint32_t SMEM_BYTES = SMEM_SCRATCH_OFF + SMEM_SCRATCH_BYTES + 1024;  // + barriers + scratch + alignment slack
// Pair kernel with helper warps (lb_tc2.cuh, HELP_): the scratch area holds, per (scanner, helper) warp pair, a queue of
// HQ_SLOTS hit groups (16 accumulators x 32 lanes each) and a control block: words [0, HQ_SLOTS) first row of the
// slot's group (HQ_END = the scanner is done), mbarriers full[s] at byte 32 + 8 s and empty[s] at byte 64 + 8 s, and two
// words per lane that are deliberately NOT synchronised: the helper's gate (single writer, only grows; the scanner tests
// against whatever it reads — a stale gate only queues a group the helper then drops) and the floor of the scanner's
// pre-pass.  (Through global memory instead, the scanner's load sits on every tile's critical path: +10 %.)
constexpr uint32_t HQ_SLOTS = 4;
constexpr uint32_t HQ_ENTRY_WORDS = 16 * 32;
constexpr uint32_t HQ_CTRL_WORDS = 96;   // ... [32, 64) the helper's gate per lane, [64, 96) the floor of the scanner's pre-pass
constexpr uint32_t HQ_END = 0xFFFFFFFFu;
constexpr uint32_t HQ_BYTES = 4 * (HQ_SLOTS * HQ_ENTRY_WORDS + HQ_CTRL_WORDS) * 4;
constexpr uint32_t SMEM_BYTES_HELP = SMEM_SCRATCH_OFF + HQ_BYTES + 1024;
static_assert(SMEM_BYTES_HELP <= 227 * 1024, "helper queues do not fit next to the staging ring");

// How the 32-bit accumulator of a (query, row) pair becomes the coarse KEY the epilogue ranks by (larger = better).
//   mode            operands        accumulator   key                                   side values
//   CM_F32          bf16 x bf16     f32           acc                                   —
//   CM_F32_BIAS     bf16 x bf16     f32           acc - side[row]        (L2: 2q.c-|c|^2) side = f32 |c|^2 per row
//   CM_I32          u8/s8 x u8      s32           acc                    (integer)      —
//   CM_I32_HAMMING  {0,1} x {0,1}   s32           2 acc - side[row]      (integer)      side = popcount(row)
//   CM_RATIO        {0,1} x {0,1}   s32           acc / (pa + pb)        (f32)          side = (f32) popcount(row), qaux = popcount(query)
// CM_RATIO serves Jaccard / Tanimoto (similarity r / (1 - r)) and Dice (2 r): both grow with r = |a & b| / (|a| + |b|), so
// r ranks the rows for all of them; the epilogue never divides on the per-score path (see KeyFn).
enum CoarseMode { CM_F32 = 0, CM_F32_BIAS = 1, CM_I32 = 2, CM_I32_HAMMING = 3, CM_RATIO = 4 };
template <int MODE>
struct ModeTraits {
    static constexpr bool kI8 = MODE >= CM_I32;                            // tcgen05 kind::i8 (K = 32 per MMA) instead of kind::f16 (K = 16)
    static constexpr bool kIntKey = MODE == CM_I32 || MODE == CM_I32_HAMMING;
    static constexpr bool kBias = MODE == CM_F32_BIAS || MODE >= CM_I32_HAMMING;
    using Key = typename std::conditional<kIntKey, int, float>::type;
};

struct TcArgs {
    const unsigned char* qb;  // [n_mtiles*128][2*Dp bytes] A operand rows (bf16 or 8-bit), zero padded
    int nq;
    int n_mtiles;             // query tiles of 128; qb is padded to a multiple of the cluster size tiles
    int Dp;                   // operand row length in 2-byte units (bf16 elements; 8-bit operands: elements / 2), multiple of 64
    int rem_kb;               // (Dp / 64) % KPS: K blocks of the last, partial stage of a tile (0 = none; uses tmap_rem)
    int n_ksteps;             // MMA K steps (32 operand bytes each) that hold data; the zero padding behind them is not multiplied
    uint32_t n_rows;
    uint32_t tiles_total;     // ceil(n_rows / rows per tile)
    uint32_t tiles_per_part;
    int P;                    // row partitions
    int lists_per_part;       // shortlists per (query, partition): 1, or 2 when the pair kernel runs two epilogue sets
    const uint64_t* allow_bits;  // optional row filter (bit r = row r allowed, LSB-first u64 words): disallowed rows never
                                 // enter a shortlist, so floors, certification and result are those of the allowed rows
    const uint32_t* bias;     // per-row side value (CoarseMode), padded to whole tiles (+inf / huge past the last row)
    const float* qaux;        // per-query side value (CoarseMode), padded like qb
    const uint32_t* idesc_extra;  // device word OR-ed into the instruction descriptor (8-bit operands: signedness bits)
    uint32_t* cand_key;       // [nq][P * lists_per_part][KP]  key bits (f32 or s32, see ModeTraits)
    uint32_t* cand_row;       // [nq][P * lists_per_part][KP]
    uint32_t* cand_thr;       // [nq][P * lists_per_part]      key bits: every row dropped by the list scored <= this
    int share_floor;          // 1: partitions of a query share a shortlist floor through gthr; 2: gthr holds a floor seeded by a
                              // pre-pass over a sample of the corpus and is only read (large k)
    uint32_t* gthr;           // [nq] zero-initialised: best published shortlist floor per query (orderable key bits)
    int pbest_first;          // diagnostics: the exchange publishes the BEST key of a list instead of its second best
    int pbest_depth;          // which of a list's best keys the exchange publishes: 2 .. 4 (the floor then has depth x P rows above it)
    uint32_t* pbest2;         // optional [nq][PBEST_STRIDE] zero-initialised: second-best key (orderable bits) each partition of a
                              // query holds so far.  When all partitions run at once, the smallest of them is a floor with at
                              // least two rows above it in EVERY partition — 2 P rows, far tighter early in the pass than the
                              // 16th best of one partition's rows so far, which is all a single list knows
    // hit mode (with share_floor == 2): rows above the seeded floor are appended to the hit region of their
    // (query, shortlist slot) instead of going through the register shortlists — a private region per epilogue
    // thread and partition, so appending is one store and a register increment, no atomic
    uint32_t* hit_count;      // [nq][P * lists_per_part] entries written per region (may exceed hit_cap: overflow); null = shortlist mode
    uint2* hit_buf;           // [nq][P * lists_per_part][hit_cap] (key bits, row)
    uint32_t hit_cap;         // entries per region
    // helper-warp kernel with the second-best exchange: every partition first runs its first pre_tiles tiles through the
    // tensor cores only to learn its two best keys (group maxima: nothing is queued), waits for the floor the exchange
    // then yields, and starts over with the whole partition — so the lists never see a partition's first tiles without
    // a floor, when every row is a hit (0 = off)
    int pre_tiles;
    uint32_t* error_flag;     // set non-zero when a barrier wait timed out
    float* dump;              // optional [n_mtiles*128][tiles_total*rows per tile] keys as f32 (diagnostics)
    // Work mapping: cluster c serves query group (c % n_mgroups) of slot (c / n_mgroups); slot s walks the row
    // partitions s, s + n_slots, s + 2 n_slots, ...  All query groups of a slot stream the same shadow tiles at the
    // same time, so HBM is read once per slot and the other groups are served from L2.
    int n_slots;
    int parts_per_slot;
    uint32_t* progress;       // [n_slots][PROGRESS_STRIDE] tiles issued per (slot, query group); zeroed per launch; null = free-running
    int poll_mask;            // shared floors are re-read when (tile & poll_mask) == 0 (7: every 8th tile; 0: every tile)
    int window;               // a query group never runs more than `window` tiles ahead of the slowest group of its slot
    unsigned long long* prof; // optional [grid][8] cycle counters of the MMA issuer / epilogue (diagnostics)
    int debug_mode;           // diagnostics only (results are garbage): bit 0 = producer skips the TMA loads,
                              // bit 1 = epilogue releases accumulators unread, bit 2 = epilogue reads but does not scan
};
constexpr int PROGRESS_STRIDE = 32;
constexpr int PBEST_STRIDE = 20;   // partitions per query the second-best exchange serves (five 16-byte loads per poll)
constexpr uint32_t EPI_SCRATCH_SLOTS = 3;    // side values travel global -> shared memory with cp.async, two tiles ahead of their use
constexpr uint32_t EPI_SCRATCH_WORDS = 128;  // per slot: the side values of the (up to) 128 rows of a tile

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
// the same with 8-bit integer operands (K = 32 per instruction), s32 accumulators
__device__ __forceinline__ void umma_ts_i8(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Instruction descriptors (cute::UMMA::InstrDescriptor): accumulator format at [4,6) (1 = f32, 2 = s32), A / B format at
// [7,10) / [10,13) (kind::f16: 1 = bf16; kind::i8: 0 = unsigned, 1 = signed), both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t IDESC_A_SIGNED = 1u << 7, IDESC_B_SIGNED = 1u << 10;
template <bool I8>
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (I8 ? (2u << 4) : ((1u << 4) | (1u << 7) | (1u << 10))) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
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
__device__ __forceinline__ void tmem_ld_32x32b_x64(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]),
          "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]),
          "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]),
          "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]),
          "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// The same wait, naming the 32 destination registers of the load it completes as read-write operands: when other work
// runs between a tcgen05.ld and its wait (the pipelined epilogue), no use of those registers can be scheduled above it.
__device__ __forceinline__ void tmem_ld_wait_32(uint32_t* v) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]),
                   "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]),
                   "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]),
                   "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 8 bytes global -> shared without passing through registers (LDGSTS); completion through the per-thread async groups
__device__ __forceinline__ void cp_async_8(uint32_t smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

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

// A operand: the calling thread's query row (2*Dp bytes: bf16 pairs or 8-bit quads per word) into TMEM columns
// [0, Dp/2) of its lane.
__device__ __forceinline__ void load_query_to_tmem(const unsigned char* qrow, int Dp, uint32_t lane_addr) {
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

// ---- coarse keys ---------------------------------------------------------------------------------------------
template <class K>
struct KeyOps;
template <>
struct KeyOps<float> {
    static __device__ __forceinline__ float lowest() { return -INFINITY; }
    static __device__ __forceinline__ float highest() { return INFINITY; }
    static __device__ __forceinline__ float max3(float a, float b, float c) { return fmax3(a, b, c); }
    static __device__ __forceinline__ float min3(float a, float b, float c) { return fmin3(a, b, c); }
    static __device__ __forceinline__ float from_bits(uint32_t u) { return __uint_as_float(u); }
    static __device__ __forceinline__ uint32_t bits(float x) { return __float_as_uint(x); }
    static __device__ __forceinline__ uint32_t orderable(float x) { return f32_orderable(x); }
    static __device__ __forceinline__ float from_orderable(uint32_t o) { return f32_from_orderable(o); }
    static __device__ __forceinline__ float as_f32(float x) { return x; }
};
template <>
struct KeyOps<int> {
    static __device__ __forceinline__ int lowest() { return (int)0x80000000; }
    static __device__ __forceinline__ int highest() { return 0x7fffffff; }
    static __device__ __forceinline__ int max3(int a, int b, int c) { return max(max(a, b), c); }  // VIMNMX3
    static __device__ __forceinline__ int min3(int a, int b, int c) { return min(min(a, b), c); }
    static __device__ __forceinline__ int from_bits(uint32_t u) { return (int)u; }
    static __device__ __forceinline__ uint32_t bits(int x) { return (uint32_t)x; }
    static __device__ __forceinline__ uint32_t orderable(int x) { return (uint32_t)x ^ 0x80000000u; }  // lowest() -> 0 = "unset"
    static __device__ __forceinline__ int from_orderable(uint32_t o) { return (int)(o ^ 0x80000000u); }
    static __device__ __forceinline__ float as_f32(int x) { return (float)x; }
};
// host / finalize side: key bits -> a 32-bit value whose unsigned order is the key order
__host__ __device__ inline uint32_t key_bits_orderable(uint32_t bits, bool int_key) { return int_key ? (bits ^ 0x80000000u) : f32_orderable_bits(bits); }

// Accumulator word + side values -> what the epilogue compares.  test(): a value that exceeds test_thr() whenever the
// key exceeds the lane's gate (it may also exceed it for a few rows just below the gate: every hit is re-checked with
// its real key); key(): the coarse key itself.  For all modes but CM_RATIO the two are the same number.  CM_RATIO's key
// is a quotient; its test is the linear form  inter - rho (pa + pb) > -0.01  (rho = the gate, all counts < 2^11 so the
// f32 evaluation is off by < 1e-4): one fused multiply-add per score instead of a division.
template <int MODE>
struct KeyFn {
    using K = typename ModeTraits<MODE>::Key;
    float qaux, rho, rho_pa;
    __device__ __forceinline__ void set_gate(K thr) {
        if (MODE == CM_RATIO) {
            const float th = KeyOps<K>::as_f32(thr);
            rho = fmaxf(th, -1.0f);       // keys are >= 0: -1 passes everything, and keeps 0 * rho finite
            rho_pa = rho * qaux;
        }
    }
    __device__ __forceinline__ K test_thr(K thr) const { return MODE == CM_RATIO ? KeyOps<K>::from_bits(__float_as_uint(-0.01f)) : thr; }
    __device__ __forceinline__ K test(uint32_t acc, uint32_t side) const {
        if (MODE == CM_F32 || MODE == CM_I32) return KeyOps<K>::from_bits(acc);
        if (MODE == CM_F32_BIAS) return KeyOps<K>::from_bits(__float_as_uint(__fsub_rn(__uint_as_float(acc), __uint_as_float(side))));
        if (MODE == CM_I32_HAMMING) return KeyOps<K>::from_bits((uint32_t)((int)acc + (int)acc - (int)side));
        // intersection count (< 2^23) -> f32 without a conversion instruction
        const float inter = __fsub_rn(__uint_as_float(acc | 0x4B000000u), 8388608.0f);
        return KeyOps<K>::from_bits(__float_as_uint(__fsub_rn(__fmaf_rn(-rho, __uint_as_float(side), inter), rho_pa)));
    }
    __device__ __forceinline__ K key(uint32_t acc, uint32_t side) const {
        if (MODE != CM_RATIO) return test(acc, side);
        const float inter = (float)(int)acc, den = __fadd_rn(qaux, __uint_as_float(side));  // popcounts: exact in f32
        // empty query and empty row: the reference's distance is 0 (the best), i.e. ratio "1"
        return KeyOps<K>::from_bits(__float_as_uint(den > 0.0f ? __fdividef(inter, den) : 1.0f));
    }
};

// Per-thread shortlist of one (query, row partition): the KP best coarse keys, held in REGISTERS (every index is
// a compile-time constant) and gated by a register threshold — shared-memory lists stall for thousands of cycles
// behind the tensor core's operand reads and the TMA writes.  lmin = worst key kept; thr_g = best floor any
// partition of the query has published through gthr (a row at or below it is outside the global top KP).
// HITS (seeded floor, large k): there is no list; a row above the floor goes to this thread's hit region.
template <int MODE, bool HITS>
struct Shortlist {
    using K = typename ModeTraits<MODE>::Key;
    using O = KeyOps<K>;
    static constexpr int NL = HITS ? 1 : KP;
    K sc[NL];
    uint32_t rw[NL];
    K lmin, thr_g, thr_pub;
    uint32_t g_bits;
    uint32_t* gthr;
    // second-best exchange (list mode, every partition of the query in flight): best and second-best key of this list,
    // where this list publishes its second best, the query's row of the exchange, the words loaded by the last poll
    K best1, best2, best3, best4;
    int pb_depth;
    uint32_t* pb_slot;
    const uint32_t* pb_row;
    uint4 pb_v[PBEST_STRIDE / 4];
    uint32_t pb_n;             // partitions of the query (0 = exchange off)
    bool pb_loaded, pb_first;
    uint2* hit_buf;
    uint32_t hit_n, hit_cap;
    KeyFn<MODE> fn;
    bool q_valid, share, may_publish;

    // The list is full and its floor rose: make it visible to the other partitions of the query.
    __device__ __forceinline__ void publish() {
        atomicMax(gthr, O::orderable(lmin));
        thr_pub = lmin;
    }
    __device__ __forceinline__ void reset(bool valid, const TcArgs& a, uint32_t gq, uint32_t part, uint32_t sub) {
#pragma unroll
        for (int j = 0; j < NL; ++j) {
            sc[j] = O::lowest();
            rw[j] = ROW_NONE;
        }
        q_valid = valid;
        share = a.share_floor != 0;
        may_publish = !HITS && a.share_floor == 1;
        gthr = a.gthr + (valid ? gq : 0);
        hit_buf = nullptr;
        hit_n = 0;
        hit_cap = 0;
        if (a.share_floor == 2 && valid) {  // seeded floor: available from the first tile
            const uint32_t bits = *reinterpret_cast<volatile uint32_t*>(gthr);
            if (bits != 0u) thr_g = max(thr_g, O::from_orderable(bits));
        }
        if (HITS && valid) {
            hit_cap = a.hit_cap;
            hit_buf = a.hit_buf + ((size_t)gq * ((size_t)a.P * a.lists_per_part) + (size_t)part * a.lists_per_part + sub) * a.hit_cap;
        }
        lmin = O::lowest();
        thr_pub = O::lowest();
        best1 = O::lowest();
        best2 = O::lowest();
        best3 = O::lowest();
        best4 = O::lowest();
        pb_depth = a.pbest_depth;
        pb_n = 0u;
        pb_loaded = false;
        pb_first = false;
        pb_slot = nullptr;
        pb_row = nullptr;
        if (!HITS && valid && a.pbest2 != nullptr && a.share_floor == 1 && a.parts_per_slot == 1 && a.lists_per_part == 1 && a.P <= PBEST_STRIDE) {
            pb_n = (uint32_t)a.P;
            pb_first = a.pbest_first != 0;
            pb_row = a.pbest2 + (size_t)gq * PBEST_STRIDE;
            pb_slot = a.pbest2 + (size_t)gq * PBEST_STRIDE + part;
        }
    }
    __device__ __forceinline__ void init_floor(float qaux) {
        thr_g = O::lowest();
        g_bits = 0u;
        fn.qaux = qaux;
    }
    __device__ __forceinline__ K gate() const { return q_valid ? (HITS ? thr_g : max(lmin, thr_g)) : O::highest(); }
    // refresh the shared floor every 8th tile; the load issued now is consumed 8 tiles later, so its (loaded) L2
    // latency never sits on the per-tile critical path
    __device__ __forceinline__ void poll_floor(uint32_t tile_iter, uint32_t mask) {
        if (!HITS && share && (tile_iter & mask) == 0u) {
            if (g_bits != 0u) thr_g = max(thr_g, O::from_orderable(g_bits));
            g_bits = *reinterpret_cast<volatile uint32_t*>(gthr);
            if (pb_n != 0u) {
                // the smallest second-best over the query's partitions (0 = some partition has not published yet: no floor)
                if (pb_loaded) {
                    uint32_t m = 0xFFFFFFFFu;
#pragma unroll
                    for (int i = 0; i < PBEST_STRIDE / 4; ++i) {
                        const uint32_t w[4] = {pb_v[i].x, pb_v[i].y, pb_v[i].z, pb_v[i].w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) m = min(m, (uint32_t)(4 * i + j) < pb_n ? w[j] : 0xFFFFFFFFu);
                    }
                    if (m != 0u) thr_g = max(thr_g, O::from_orderable(m));
                }
#pragma unroll
                for (int i = 0; i < PBEST_STRIDE / 4; ++i) {
                    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(pb_v[i].x), "=r"(pb_v[i].y), "=r"(pb_v[i].z), "=r"(pb_v[i].w)
                                 : "l"(pb_row + 4 * i));
                }
                pb_loaded = true;
            }
        }
    }
    // replace the current minimum by (key, row) and recompute the minimum: ~70 ALU instructions, no memory
    __device__ __forceinline__ void insert(K key, uint32_t row) {
        if (HITS) return;
        if (pb_n != 0u) {
            // the list's four best keys, sorted; the exchange publishes the pb_depth-th of them (single writer, only grows)
            const K old1 = best1, oldp = pb_depth <= 2 ? best2 : (pb_depth == 3 ? best3 : best4);
            K x = key, t;
            t = max(best1, x); x = min(best1, x); best1 = t;
            t = max(best2, x); x = min(best2, x); best2 = t;
            t = max(best3, x); x = min(best3, x); best3 = t;
            best4 = max(best4, x);
            const K newp = pb_depth <= 2 ? best2 : (pb_depth == 3 ? best3 : best4);
            if (pb_first) {
                if (best1 > old1) *reinterpret_cast<volatile uint32_t*>(pb_slot) = O::orderable(best1);
            } else if (newp > oldp) {
                *reinterpret_cast<volatile uint32_t*>(pb_slot) = O::orderable(newp);
            }
        }
        bool done = false;
#pragma unroll
        for (int j = 0; j < NL; ++j) {
            const bool hit = !done && sc[j] == lmin;
            sc[j] = hit ? key : sc[j];
            rw[j] = hit ? row : rw[j];
            done = done || hit;
        }
        if (!HITS) {
            K m0 = O::min3(sc[0], sc[1], sc[2]), m1 = O::min3(sc[3 % NL], sc[4 % NL], sc[5 % NL]);
            m0 = O::min3(m0, sc[6 % NL], sc[7 % NL]);
            m1 = O::min3(m1, sc[8 % NL], sc[9 % NL]);
            m0 = O::min3(m0, sc[10 % NL], sc[11 % NL]);
            m1 = O::min3(m1, sc[12 % NL], sc[13 % NL]);
            lmin = O::min3(min(m0, m1), sc[14 % NL], sc[15 % NL]);
        }
    }
    // hit region append: one store and a register increment (counts past the capacity too: finalize sees the overflow)
    __device__ __forceinline__ void append(K key, uint32_t row) {
        if (hit_n < hit_cap) hit_buf[hit_n] = make_uint2(O::bits(key), row);
        ++hit_n;
    }
    // 16 accumulators of this thread's query = rows row0 .. row0+15, at least one lane of the warp above its gate.
    // `side`: the rows' side values in shared memory.
    __device__ __forceinline__ void slow16(const uint32_t* v, const uint32_t* side, K gmax, uint32_t row0, uint32_t row_end, K thr, K tthr,
                                           uint32_t allow16, bool have_allow) {
        constexpr bool kBias = ModeTraits<MODE>::kBias;
        if (HITS) {
            // straight-line, predicated: appending is cheap, and no loop, vote or select tree keeps this path short
            uint32_t ok = 0xffffu;
            if (row0 + 16u > row_end) ok = (1u << (row_end > row0 ? row_end - row0 : 0u)) - 1u;
            if (have_allow) ok &= allow16;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const uint32_t sv = kBias ? side[i] : 0u;
                if (fn.test(v[i], sv) > tthr && ((ok >> i) & 1u)) {
                    const K x = fn.key(v[i], sv);
                    if (MODE != CM_RATIO || x > thr) append(x, row0 + (uint32_t)i);
                }
            }
            return;
        }
        uint32_t mask = 0u;
#pragma unroll
        for (int i = 0; i < 16; ++i) mask |= (fn.test(v[i], kBias ? side[i] : 0u) > tthr ? 1u : 0u) << i;
        const uint32_t raw = mask;
        if (row0 + 16u > row_end) {  // last tile: rows past the end of the corpus never enter a list
            const uint32_t n_ok = row_end > row0 ? row_end - row0 : 0u;
            mask &= (1u << n_ok) - 1u;
        }
        if (have_allow) mask &= allow16;
        // a lane with exactly one hit (the usual case) already holds its key: it is the group maximum — unless the
        // masks above removed a hit, in which case the maximum may belong to a removed row (CM_RATIO: the maximum is
        // a test value, not a key)
        if (MODE != CM_RATIO && mask == raw && __popc(mask) == 1) {
            insert(gmax, row0 + (uint32_t)(__ffs((int)mask) - 1));
            mask = 0u;
        }
#pragma unroll 1
        while (__any_sync(0xffffffffu, mask != 0u)) {
            if (mask != 0u) {
                const int idx = __ffs((int)mask) - 1;
                mask &= mask - 1u;
                // v[idx] with a run-time idx: 4-level select tree over the register array (15 selects)
                uint32_t t3[8], t2[4], t1[2];
#pragma unroll
                for (int j = 0; j < 8; ++j) t3[j] = (idx & 8) ? v[8 + j] : v[j];
#pragma unroll
                for (int j = 0; j < 4; ++j) t2[j] = (idx & 4) ? t3[4 + j] : t3[j];
#pragma unroll
                for (int j = 0; j < 2; ++j) t1[j] = (idx & 2) ? t2[2 + j] : t2[j];
                const K x = fn.key((idx & 1) ? t1[1] : t1[0], kBias ? side[idx] : 0u);
                if (x > gate()) insert(x, row0 + (uint32_t)idx);  // the gate may have risen since the mask was taken
            }
        }
    }
    // 16 * NG accumulators of this thread's query = rows row0 .. row0 + 16 NG - 1 (row0 a multiple of 32).  Called by whole warps.
    // The slow path must stay SMALL: a fully unrolled "for each of the 64 scores: compare, insert" is ~90 KB of code
    // whose sparse execution misses the instruction cache at every step (~7700 cycles per tile measured).  So: the
    // maxima of the 16-score groups gate (1) a 16-bit hit mask per lane from straight compares and (2) a rolled
    // loop that pops each lane's lowest hit and fetches the accumulator with a select tree.
    // The scan in two parts, so that a caller can put work (releasing the accumulator buffer) between the cheap test and
    // the rare slow path: scan_fast computes the group maxima against the current gate and tells whether any lane of the
    // warp has a hit; scan_slow visits the groups that do.  Nothing else may touch the gate in between.
    template <int NG>
    struct ScanState {
        K g[NG];
        K thr, tthr;
    };
    template <int NG>
    __device__ __forceinline__ bool scan_fast(const uint32_t* v, const uint32_t* side, bool disabled, ScanState<NG>& st) {
        constexpr bool kBias = ModeTraits<MODE>::kBias;
        st.thr = disabled ? O::highest() : gate();
        fn.set_gate(st.thr);
        st.tthr = fn.test_thr(st.thr);
#pragma unroll
        for (int j = 0; j < NG; ++j) {
            uint32_t s[16];
            if (kBias) {
#pragma unroll
                for (int i = 0; i < 16; i += 4) {  // same address in every lane: broadcast reads
                    const uint4 s4 = *reinterpret_cast<const uint4*>(side + 16 * j + i);
                    s[i] = s4.x; s[i + 1] = s4.y; s[i + 2] = s4.z; s[i + 3] = s4.w;
                }
            }
            const uint32_t* a = v + 16 * j;
            // group maximum with 3-input max: 8 instructions for 16 test values
            K m = O::max3(fn.test(a[0], kBias ? s[0] : 0u), fn.test(a[1], kBias ? s[1] : 0u), fn.test(a[2], kBias ? s[2] : 0u));
#pragma unroll
            for (int i = 3; i + 1 < 16; i += 2) m = O::max3(m, fn.test(a[i], kBias ? s[i] : 0u), fn.test(a[i + 1], kBias ? s[i + 1] : 0u));
            st.g[j] = max(m, fn.test(a[15], kBias ? s[15] : 0u));
        }
        K gall = st.g[0];
#pragma unroll
        for (int j = 1; j < NG; ++j) gall = max(gall, st.g[j]);
        return __any_sync(0xffffffffu, gall > st.tthr) != 0;
    }
    template <int NG>
    __device__ __forceinline__ void scan_slow(const uint32_t* v, const uint32_t* side, uint32_t row0, uint32_t row_end,
                                              const uint64_t* __restrict__ allow, const ScanState<NG>& st) {
        uint64_t w = ~0ull;
        const bool have_allow = allow != nullptr;
        if (have_allow) w = row0 < row_end ? __ldg(allow + (row0 >> 6)) >> (row0 & 63u) : 0ull;  // bit 0 = row0
#pragma unroll
        for (int j = 0; j < NG; ++j)
            if (__any_sync(0xffffffffu, st.g[j] > st.tthr))
                slow16(v + 16 * j, side + 16 * j, st.g[j], row0 + 16u * j, row_end, st.thr, st.tthr, (uint32_t)(w >> (16 * j)) & 0xffffu, have_allow);
        if (may_publish && q_valid && lmin > thr_pub && lmin > thr_g) publish();
    }
    // 16 * NG accumulators of this thread's query = rows row0 .. row0 + 16 NG - 1 (row0 a multiple of 32).  Called by whole warps.
    // The slow path must stay SMALL: a fully unrolled "for each of the 64 scores: compare, insert" is ~90 KB of code
    // whose sparse execution misses the instruction cache at every step (~7700 cycles per tile measured).  So: the
    // maxima of the 16-score groups gate (1) a 16-bit hit mask per lane from straight compares and (2) a rolled
    // loop that pops each lane's lowest hit and fetches the accumulator with a select tree.
    template <int NG>
    __device__ __forceinline__ void scan_groups(const uint32_t* v, const uint32_t* side, uint32_t row0, uint32_t row_end, bool disabled,
                                                const uint64_t* __restrict__ allow = nullptr) {
        ScanState<NG> st;
        if (scan_fast<NG>(v, side, disabled, st)) scan_slow<NG>(v, side, row0, row_end, allow, st);
    }
    __device__ __forceinline__ void scan64(const uint32_t* v, const uint32_t* side, uint32_t row0, uint32_t row_end, bool disabled,
                                           const uint64_t* __restrict__ allow = nullptr) {
        scan_groups<4>(v, side, row0, row_end, disabled, allow);
    }
    // `sub`: which of the partition's lists_per_part shortlists this is (the pair kernel with two epilogue sets)
    __device__ __forceinline__ void flush(const TcArgs& a, uint32_t gq, uint32_t part, uint32_t sub = 0) {
        if (!q_valid) return;
        const size_t list = (size_t)gq * ((size_t)a.P * a.lists_per_part) + (size_t)part * a.lists_per_part + sub;
        if (HITS) {
            a.hit_count[list] = hit_n;
        } else {
            const size_t o = list * KP;
#pragma unroll
            for (int j = 0; j < NL; j += 4) {
                *reinterpret_cast<uint4*>(a.cand_key + o + j) = make_uint4(O::bits(sc[j]), O::bits(sc[(j + 1) % NL]), O::bits(sc[(j + 2) % NL]), O::bits(sc[(j + 3) % NL]));
                *reinterpret_cast<uint4*>(a.cand_row + o + j) = make_uint4(rw[j], rw[(j + 1) % NL], rw[(j + 2) % NL], rw[(j + 3) % NL]);
            }
        }
        // every row this thread dropped scored <= max(lmin, thr_g) at the time, and both only grow
        a.cand_thr[list] = O::bits(max(lmin, thr_g));
    }
};

// ---- shadow / query preparation --------------------------------------------------------------------------------
enum ShadowKind { SHADOW_IP = 0, SHADOW_COSINE = 1, SHADOW_L2 = 2 };
enum OperandKind { OPERAND_BF16 = 0, OPERAND_U8 = 1 };

// Per-shadow statistics, accumulated on the device while the shadow is built.  c' = the transformed row the shadow
// approximates (c; c / |c| for cosine), c~ = the value the operand really holds (bf16(c'), or zero + scale * u8).
// The certification of finalize_kernel uses cmax and emax:  |q~.c~ - q'.c'| <= |dq| cmax + |q~| emax  (Cauchy-Schwarz),
// with every norm MEASURED, which is both rigorous and tighter than the worst case of the element-wise rounding.
struct ShadowStats {
    uint32_t cmax_bits;   // max over rows of |c'|           (f32 bits of a non-negative value: unsigned order = value order)
    uint32_t emax_bits;   // max over rows of |c' - c~|
    uint32_t vmin_ord;    // orderable bits of the smallest element of c' (range of the 8-bit quantisation)
    uint32_t vmax_ord;    // ... and of the largest
    uint32_t nonfinite;   // rows holding NaN / inf: the plan is not used for such a corpus
    uint32_t e16max_bits; // max over rows of |c' - bf16(c')| (written by the range pass: what a bf16 operand would have cost)
    uint32_t pad[2];
};
// Per-query statistics written by the query preparation kernels.
struct QStat {
    float qt_norm;   // |q~|  : norm of the operand's value of the transformed query q' (q; q/|q|; for L2 q — the factor 2 is applied later)
    float dq_norm;   // |q' - q~|
    float q_norm;    // |q| of the original query
    float s_q;       // 8-bit operands: q~_i = s_q * qhat_i
    float sum_qhat;  // 8-bit operands: sum of qhat_i (exact: |sum| < 2^24)
    float pad[3];
};

__host__ __device__ inline size_t shadow_chunk_offset(uint64_t row, int chunk /* 16-byte chunk of the operand row */, int nkb) {
    const uint64_t tile = row >> 6;
    const uint32_t half = (uint32_t)(row >> 5) & 1u, r = (uint32_t)row & 31u;
    const uint32_t kb = (uint32_t)chunk >> 3, c = (uint32_t)chunk & 7u;
    return (size_t)(((tile * (uint64_t)nkb + kb) * 2 + half) * HALF_BLOCK_BYTES) + (size_t)r * 128 + (size_t)((c ^ (r & 7u)) << 4);
}
__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
// One warp per row, first pass: |c|, the element range of c' and non-finite rows (everything the 8-bit quantisation needs
// to know before it can write a byte).
template <class RT>
__global__ void shadow_range_kernel(const RT* __restrict__ rows, uint64_t first_row, uint64_t n, int dim, int kind,
                                    ShadowStats* __restrict__ st) {
    const uint64_t row = first_row + (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= first_row + n) return;
    const RT* r = rows + row * dim;
    float ss = 0.0f, mn = INFINITY, mx = -INFINITY;
    for (int d = lane; d < dim; d += 32) {
        const float x = ldrow(r + d);
        ss = fmaf(x, x, ss);
        mn = fminf(mn, x);
        mx = fmaxf(mx, x);
    }
    ss = warp_sum(ss);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    float scale = 1.0f;
    if (kind == SHADOW_COSINE) {
        const float norm = sqrtf(ss);
        scale = norm > 0.0f ? 1.0f / norm : 0.0f;
    }
    float e16 = 0.0f;
    for (int d = lane; d < dim; d += 32) {
        const float x = ldrow(r + d) * scale;
        const float de = x - __bfloat162float(__float2bfloat16_rn(x));
        e16 = fmaf(de, de, e16);
    }
    e16 = warp_sum(e16);
    if (lane != 0) return;
    if (!isfinite(ss)) {
        atomicAdd(&st->nonfinite, 1u);
        return;
    }
    atomicMin(&st->vmin_ord, f32_orderable(mn * scale));
    atomicMax(&st->vmax_ord, f32_orderable(mx * scale));
    atomicMax(&st->e16max_bits, __float_as_uint(sqrtf(e16)));
}

// One warp per row: the row's operand values (zero padded to the operand row length) written into the tiled,
// pre-swizzled layout described at the top of this file, + the statistics the certification needs.
//   SHADOW_IP      c' = c
//   SHADOW_COSINE  c' = c / |c|       (zero rows stay zero: cosine distance 1.0, simd.rs:1631-1633)
//   SHADOW_L2      c' = c; |c|^2 either as three bf16 columns [dim, dim+3) (n1 + n2 + n3 ~ |c|^2: the contraction itself
//                  forms 2 q.c - |c|^2 against a query of [2q, -1, -1, -1]) or, when side != null, as the f32 side value
//                  the epilogue subtracts (CM_F32_BIAS: rows whose padded length has no room for the columns)
// OPERAND_BF16: c~ = bf16(c').  OPERAND_U8: c~ = zero + scale * u8, u8 = clamp(rint((c' - zero) / scale), 0, 255).
template <int OPK, class RT>
__global__ void build_shadow_kernel(const RT* __restrict__ rows, uint64_t first_row, uint64_t n, int dim, int row_bytes, int kind,
                                    unsigned char* __restrict__ shadow, float* __restrict__ side, ShadowStats* __restrict__ st,
                                    float q_scale, float q_zero) {
    const uint64_t row = first_row + (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= first_row + n) return;
    const RT* r = rows + row * dim;
    float ss = 0.0f;
    for (int d = lane; d < dim; d += 32) {
        const float x = ldrow(r + d);
        ss = fmaf(x, x, ss);
    }
    ss = warp_sum(ss);
    const float norm = sqrtf(ss);
    float scale = 1.0f;
    if (kind == SHADOW_COSINE) scale = norm > 0.0f ? 1.0f / norm : 0.0f;
    const int nkb = row_bytes / 128;
    constexpr int EPC = OPK == OPERAND_BF16 ? 8 : 16;  // elements per 16-byte chunk
    const float inv_q = OPK == OPERAND_U8 ? 1.0f / q_scale : 0.0f;
    float err = 0.0f, tss = 0.0f;  // |c' - c~|^2 and |c'|^2
    for (int chunk = lane; chunk < row_bytes / 16; chunk += 32) {
        __align__(16) unsigned char out[16];
#pragma unroll
        for (int e = 0; e < EPC; ++e) {
            const int d = chunk * EPC + e;
            const float x = d < dim ? ldrow(r + d) * scale : 0.0f;
            float held;
            if (OPK == OPERAND_BF16) {
                __nv_bfloat16 b = __float2bfloat16_rn(x);
                held = __bfloat162float(b);
                if (kind == SHADOW_L2 && side == nullptr && d >= dim && d < dim + 3) {
                    const float n1 = __bfloat162float(__float2bfloat16_rn(ss));
                    const float n2 = __bfloat162float(__float2bfloat16_rn(ss - n1));
                    const float n3 = (ss - n1) - n2;
                    b = __float2bfloat16_rn(d == dim ? n1 : (d == dim + 1 ? n2 : n3));
                }
                reinterpret_cast<__nv_bfloat16*>(out)[e] = b;
            } else {
                int u = d < dim ? __float2int_rn((x - q_zero) * inv_q) : 0;
                u = min(max(u, 0), 255);
                out[e] = (unsigned char)u;
                held = d < dim ? fmaf(q_scale, (float)u, q_zero) : 0.0f;
            }
            const float de = x - held;
            err = fmaf(de, de, err);
            tss = fmaf(x, x, tss);
        }
        *reinterpret_cast<uint4*>(shadow + shadow_chunk_offset(row, chunk, nkb)) = *reinterpret_cast<const uint4*>(out);
    }
    err = warp_sum(err);
    tss = warp_sum(tss);
    if (lane == 0) {
        if (kind == SHADOW_L2 && side != nullptr) side[row] = ss;
        if (isfinite(ss)) {
            atomicMax(&st->cmax_bits, __float_as_uint(sqrtf(tss)));
            atomicMax(&st->emax_bits, __float_as_uint(sqrtf(err)));
        } else {
            atomicAdd(&st->nonfinite, 1u);
        }
    }
}

// Packed one-bit rows -> one byte per bit ({0,1}, the 8-bit operand of the binary metrics) in the tiled layout,
// + popcount(row) as the row's side value.  One warp per row; bit i of a row = word i/64, bit i%64 (simd.rs:750-757).
static __global__ void build_bits_shadow_kernel(const uint64_t* __restrict__ words, uint64_t first_row, uint64_t n, int n_words,
                                                int row_bytes, unsigned char* __restrict__ shadow, uint32_t* __restrict__ side_u32,
                                                float* __restrict__ side_f32) {
    const uint64_t row = first_row + (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= first_row + n) return;
    const uint64_t* w = words + row * n_words;
    const int nkb = row_bytes / 128;
    uint32_t pop = 0;
    for (int chunk = lane; chunk < row_bytes / 16; chunk += 32) {  // 16 bits -> 16 bytes
        const int word = chunk >> 2, sh = (chunk & 3) * 16;
        const uint32_t bits = word < n_words ? (uint32_t)(__ldg(w + word) >> sh) & 0xffffu : 0u;
        pop += __popc(bits);
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t nib = (bits >> (4 * j)) & 0xfu;
            o[j] = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
        }
        *reinterpret_cast<uint4*>(shadow + shadow_chunk_offset(row, chunk, nkb)) = make_uint4(o[0], o[1], o[2], o[3]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pop += __shfl_xor_sync(0xffffffffu, pop, o);
    if (lane == 0) {
        side_u32[row] = pop;
        side_f32[row] = (float)pop;
    }
}
static __global__ void fill_u32_kernel(uint32_t* out, uint64_t n, uint32_t value) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) out[i] = value;
}
// packed query words -> {0,1} bytes (A operand rows, zero padded) + popcount as the query's side value
static __global__ void prepare_bits_queries_kernel(const uint64_t* __restrict__ qwords, int nq, int nq_pad, int n_words, int row_bytes,
                                                   unsigned char* __restrict__ qb, float* __restrict__ qaux) {
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= nq_pad) return;
    uint32_t pop = 0;
    for (int chunk = lane; chunk < row_bytes / 16; chunk += 32) {
        const int word = chunk >> 2, sh = (chunk & 3) * 16;
        const uint32_t bits = (q < nq && word < n_words) ? (uint32_t)(__ldg(qwords + (size_t)q * n_words + word) >> sh) & 0xffffu : 0u;
        pop += __popc(bits);
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t nib = (bits >> (4 * j)) & 0xfu;
            o[j] = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
        }
        *reinterpret_cast<uint4*>(qb + (size_t)q * row_bytes + (size_t)chunk * 16) = make_uint4(o[0], o[1], o[2], o[3]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pop += __shfl_xor_sync(0xffffffffu, pop, o);
    if (lane == 0) qaux[q] = (float)pop;
}

// Queries, bf16 operand: A operand rows (zero padded to nq_pad x row_bytes) + statistics.
//   SHADOW_IP  q' = q       SHADOW_COSINE  q' = q / |q|       SHADOW_L2  q' = q, operand = 2 q~ (exact doubling)
static __global__ void prepare_queries_kernel(const float* __restrict__ queries, int nq, int nq_pad, int dim, int row_bytes, int kind,
                                              int norm_cols /* L2: operand columns [dim, dim+3) = -1 (the shadow holds |c|^2 there) */,
                                              unsigned char* __restrict__ qb, QStat* __restrict__ qstat) {
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= nq_pad) return;
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(qb + (size_t)q * row_bytes);
    const int Dp = row_bytes / 2;
    if (q >= nq) {
        for (int d = lane; d < Dp; d += 32) out[d] = __float2bfloat16_rn(0.0f);
        return;
    }
    const float* r = queries + (size_t)q * dim;
    float ss = 0.0f;
    for (int d = lane; d < dim; d += 32) {
        const float x = ldrow(r + d);
        ss = fmaf(x, x, ss);
    }
    ss = warp_sum(ss);
    const float norm = sqrtf(ss);
    float scale = 1.0f;
    if (kind == SHADOW_COSINE) scale = norm > 0.0f ? 1.0f / norm : 0.0f;
    float err = 0.0f, hss = 0.0f;
    for (int d = lane; d < Dp; d += 32) {
        const float x = d < dim ? ldrow(r + d) * scale : 0.0f;
        const __nv_bfloat16 b = __float2bfloat16_rn(x);
        const float held = __bfloat162float(b);
        out[d] = kind == SHADOW_L2 ? __float2bfloat16_rn((norm_cols && d >= dim && d < dim + 3) ? -1.0f : 2.0f * held) : b;
        const float de = x - held;
        err = fmaf(de, de, err);
        hss = fmaf(held, held, hss);
    }
    err = warp_sum(err);
    hss = warp_sum(hss);
    if (lane == 0) {
        QStat s{};
        s.qt_norm = sqrtf(hss);
        s.dq_norm = sqrtf(err);
        s.q_norm = norm;
        s.s_q = 1.0f;
        qstat[q] = s;
    }
}

// Queries, 8-bit operand, pass 1: per-query element range of q' and whether any query of the batch has a negative element
// (then the whole batch is quantised as signed bytes: the A format is one bit of the instruction descriptor).
static __global__ void query_range_kernel(const float* __restrict__ queries, int nq, int dim, int kind, float* __restrict__ qrange /*[nq][2]*/,
                                          uint32_t* __restrict__ idesc_extra) {
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= nq) return;
    const float* r = queries + (size_t)q * dim;
    float ss = 0.0f, mn = INFINITY, mx = -INFINITY;
    for (int d = lane; d < dim; d += 32) {
        const float x = ldrow(r + d);
        ss = fmaf(x, x, ss);
        mn = fminf(mn, x);
        mx = fmaxf(mx, x);
    }
    ss = warp_sum(ss);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane != 0) return;
    if (kind == SHADOW_COSINE) {
        const float norm = sqrtf(ss), scale = norm > 0.0f ? 1.0f / norm : 0.0f;
        mn *= scale;
        mx *= scale;
    }
    qrange[2 * q] = mn;
    qrange[2 * q + 1] = mx;
    if (mn < 0.0f) atomicOr(idesc_extra, IDESC_A_SIGNED);
}
// pass 2: q~_i = s_q * qhat_i with qhat a u8 (s_q = max / 255) or, for a batch with negative elements, an s8 (s_q = max|.| / 127)
static __global__ void quantise_queries_kernel(const float* __restrict__ queries, int nq, int nq_pad, int dim, int row_bytes, int kind,
                                               const float* __restrict__ qrange, const uint32_t* __restrict__ idesc_extra,
                                               unsigned char* __restrict__ qb, QStat* __restrict__ qstat) {
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= nq_pad) return;
    unsigned char* out = qb + (size_t)q * row_bytes;
    if (q >= nq) {
        for (int d = lane * 4; d < row_bytes; d += 128) *reinterpret_cast<uint32_t*>(out + d) = 0u;
        return;
    }
    const bool is_signed = (__ldg(idesc_extra) & IDESC_A_SIGNED) != 0u;
    const float* r = queries + (size_t)q * dim;
    float ss = 0.0f;
    for (int d = lane; d < dim; d += 32) {
        const float x = ldrow(r + d);
        ss = fmaf(x, x, ss);
    }
    ss = warp_sum(ss);
    const float norm = sqrtf(ss);
    float scale = 1.0f;
    if (kind == SHADOW_COSINE) scale = norm > 0.0f ? 1.0f / norm : 0.0f;
    const float amax = fmaxf(fabsf(qrange[2 * q]), fabsf(qrange[2 * q + 1]));
    const float levels = is_signed ? 127.0f : 255.0f;
    const float s_q = (amax > 0.0f && isfinite(amax)) ? amax / levels : 1.0f;
    const float inv = 1.0f / s_q;
    float err = 0.0f, hss = 0.0f, hsum = 0.0f;
    for (int d0 = lane * 4; d0 < row_bytes; d0 += 128) {
        uint32_t word = 0u;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int d = d0 + e;
            const float x = d < dim ? ldrow(r + d) * scale : 0.0f;
            int u = __float2int_rn(x * inv);
            u = is_signed ? min(max(u, -127), 127) : min(max(u, 0), 255);
            const float held = s_q * (float)u;
            const float de = x - held;
            err = fmaf(de, de, err);
            hss = fmaf(held, held, hss);
            hsum += (float)u;
            word |= ((uint32_t)u & 0xffu) << (8 * e);
        }
        *reinterpret_cast<uint32_t*>(out + d0) = word;
    }
    err = warp_sum(err);
    hss = warp_sum(hss);
    hsum = warp_sum(hsum);
    if (lane == 0) {
        QStat s{};
        s.qt_norm = sqrtf(hss);
        s.dq_norm = sqrtf(err);
        s.q_norm = norm;
        s.s_q = s_q;
        s.sum_qhat = hsum;
        qstat[q] = s;
    }
}

// ---- seeded floors for large k ---------------------------------------------------------------------------------
// After a pre-pass over a sample of the corpus: the r-th best coarse key of the sample, per query, becomes the
// floor the main pass starts with (gthr).  It is a heuristic gate — about r * (rows / sample rows) rows of the corpus
// score above it — and the certification of finalize_kernel is what keeps the result exact: a floor that turns out
// too high leaves the query uncertified and it is re-run by the exact scan.
static __global__ void __launch_bounds__(256) seed_floor_kernel(const uint32_t* __restrict__ cand_key, const uint32_t* __restrict__ cand_row, int P,
                                                                int M, int r, int int_key, uint32_t* __restrict__ gthr) {
    extern __shared__ __align__(16) unsigned char smem_seed[];
    uint64_t* s = reinterpret_cast<uint64_t*>(smem_seed);
    const int q = blockIdx.x, total = P * KP;
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
        uint64_t key = KEY_NONE;
        if (i < total && cand_row[(size_t)q * total + i] != ROW_NONE)
            key = ((uint64_t)(~key_bits_orderable(cand_key[(size_t)q * total + i], int_key != 0)) << 32) | (uint32_t)i;  // best key first
        s[i] = key;
    }
    bitonic_sort_u64(s, M);
    if (threadIdx.x == 0) gthr[q] = (r >= 1 && r <= M && s[r - 1] != KEY_NONE) ? ~(uint32_t)(s[r - 1] >> 32) : 0u;
}

// ---- finalize: shortlist -> exact-order rescore -> certified top-k -------------------------------------------------
struct FinArgs {
    // list mode: the shortlists of the partitions
    const uint32_t* cand_key;  // [nq][P][KP] key bits
    const uint32_t* cand_row;
    const uint32_t* cand_thr;  // [nq][P] key bits
    int P;
    // hit mode (hit_count != null): the rows above the seeded floor gthr[q], in P hit regions of hit_cap entries
    const uint32_t* hit_count;  // [nq][P]
    const uint2* hit_buf;       // [nq][P][hit_cap]
    uint32_t hit_cap;
    const uint32_t* gthr;
    int int_key;              // keys are s32 (8-bit operands) instead of f32
    int M1;                   // pow2 >= number of candidates read (<= 4096)
    int R;                    // rescore budget, pow2 <= 1024, >= k
    int R1;                   // first round of the rescore (pow2, k <= R1 <= R; 0 = one round)
    const void* corpus;       // f32 or binary16 rows (finalize_kernel's RT)
    int dim;
    const float* queries;     // original f32 queries [nq][dim]
    const QStat* qstat;       // [nq]
    const ShadowStats* sstat;
    int operand;              // OperandKind
    float c_scale, c_zero;    // OPERAND_U8: c~ = c_zero + c_scale * u8
    int nq, k, metric;
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
constexpr int FIN_COLS = 128;   // columns staged per round (128: 17 KiB of row buffers per 256-thread block, seven blocks per SM)
constexpr int FIN_STRIDE = FIN_COLS + 8;  // row pitch of the staging buffer: the four rows of a warp land in different banks
__device__ __forceinline__ float hsum8_shfl(float a) {   // simd.rs:1427-1436 over the 8 threads of a row
    a = a + __shfl_xor_sync(0xffffffffu, a, 4);           // s[i] = acc[i] + acc[i+4]
    a = a + __shfl_xor_sync(0xffffffffu, a, 1);           // t0 = s0 + s1, t2 = s2 + s3
    return a + __shfl_xor_sync(0xffffffffu, a, 2);        // t0 + t2
}
// the row's 8 threads fetch its columns [col0, col0 + ncol) into rowbuf: 128 contiguous bytes per step
__device__ __forceinline__ void stage_row(float* rowbuf, const float* __restrict__ c, int col0, int ncol, int sub) {
    for (int x = sub * 4; x < ncol; x += 32) *reinterpret_cast<float4*>(rowbuf + x) = __ldg(reinterpret_cast<const float4*>(c + col0 + x));
}
__device__ __forceinline__ void stage_row(float* rowbuf, const __half* __restrict__ c, int col0, int ncol, int sub) {
    for (int x = sub * 8; x < ncol; x += 64) {
        const Vec8 v = load8<true>(c + col0 + x, true);
        *reinterpret_cast<float4*>(rowbuf + x) = make_float4(v.v[0], v.v[1], v.v[2], v.v[3]);
        *reinterpret_cast<float4*>(rowbuf + x + 4) = make_float4(v.v[4], v.v[5], v.v[6], v.v[7]);
    }
}
template <class CP>
__device__ __forceinline__ float rescore_row_lanes(int metric, bool two_acc_ip, const float* __restrict__ sq, CP c,
                                                   int dim, float* rowbuf /* this row's FIN_COLS floats */, bool row_ok, int lane) {
    const int i = lane & 7, sub = lane & 7;
    const int chunks = dim >> 3;
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
    for (int col0 = 0; col0 < chunks * 8; col0 += FIN_COLS) {
        const int ncol = min(FIN_COLS, chunks * 8 - col0);
        __syncwarp();
        if (row_ok) stage_row(rowbuf, c, col0, ncol, sub);
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
        for (int x = t0; x < dim; ++x) out = out + sq[x] * ldrow(c + (row_ok ? x : 0));
        return out;
    } else if (metric == LB_L2) {
        a0 = a0 + a1;
        float sum = hsum8_shfl(a0);
        for (int x = t0; x < dim; ++x) {
            const float diff = sq[x] - ldrow(c + (row_ok ? x : 0));
            sum = sum + diff * diff;
        }
        return sum;
    }
    float dot = hsum8_shfl(a0), na = hsum8_shfl(a1), nb = hsum8_shfl(a2);
    for (int x = t0; x < dim; ++x) {
        const float qa = sq[x], cb = ldrow(c + (row_ok ? x : 0));
        dot = dot + qa * cb;
        na = na + qa * qa;
        nb = nb + cb * cb;
    }
    const float denom = sqrtf(na * nb);
    if (denom < 1e-30f) return 1.0f;
    return 1.0f - dot / denom;
}

// Candidates of query q -> sort keys in s[0, M) (best coarse key first), sorted; M = a power of two <= M1 that holds
// them.  Returns M; through shared words: the number of candidates and the orderable bits of T — every row that is NOT
// a candidate has a coarse key <= T.
__device__ __forceinline__ int gather_candidates(const FinArgs& a, int q, uint64_t* s, uint32_t* sh_T, uint32_t* sh_ncand, uint32_t* sh_overflow) {
    const int tid = threadIdx.x;
    const bool ik = a.int_key != 0;
    if (a.hit_count != nullptr) {
        // hit regions of the query's P shortlist slots: compact the written entries into s (order does not matter, they
        // are sorted next); more than M1 of them, or a region that ran out of room, leaves the query uncertified
        for (int i = tid; i < a.M1; i += blockDim.x) s[i] = KEY_NONE;
        __syncthreads();
        // one flat sweep over (region, entry): neighbouring threads read neighbouring entries, and every thread has many
        // independent loads in flight (a loop over the regions would pay one memory round trip per region)
        const uint32_t cap = a.hit_cap, total = (uint32_t)a.P * cap;
        const uint32_t* cnts = a.hit_count + (size_t)q * a.P;
        const uint2* regs = a.hit_buf + (size_t)q * a.P * cap;
        for (uint32_t idx = tid; idx < total; idx += blockDim.x) {
            const uint32_t r = idx / cap, i = idx - r * cap;
            const uint32_t cnt = __ldg(cnts + r);
            if (i == 0 && cnt > cap) *sh_overflow = 1u;
            if (i < min(cnt, cap)) {
                const uint2 h = regs[idx];
                const uint32_t slot = atomicAdd(sh_ncand, 1u);
                if (slot < (uint32_t)a.M1) s[slot] = ((uint64_t)(~key_bits_orderable(h.x, ik)) << 32) | h.y;
            }
        }
        __syncthreads();
        if (tid == 0) {
            *sh_T = a.gthr[q];  // orderable bits of the seeded floor (0 = none: nothing was dropped)
            if (*sh_ncand > (uint32_t)a.M1) {
                *sh_overflow = 1u;
                *sh_ncand = (uint32_t)a.M1;
            }
        }
        __syncthreads();
    } else {
        const int total = a.P * KP;
        uint32_t local_valid = 0;
        for (int i = tid; i < a.M1; i += blockDim.x) {
            uint64_t key = KEY_NONE;
            if (i < total) {
                const uint32_t row = a.cand_row[(size_t)q * total + i];
                if (row != ROW_NONE) {
                    key = ((uint64_t)(~key_bits_orderable(a.cand_key[(size_t)q * total + i], ik)) << 32) | row;
                    ++local_valid;
                }
            }
            s[i] = key;
        }
        for (int p = tid; p < a.P; p += blockDim.x) atomicMax(sh_T, key_bits_orderable(a.cand_thr[(size_t)q * a.P + p], ik));
        if (local_valid) atomicAdd(sh_ncand, local_valid);
        __syncthreads();
    }
    int M = a.M1;
    if (a.hit_count != nullptr) {  // sort only as many slots as hold candidates
        M = 32;
        while (M < (int)*sh_ncand) M <<= 1;
        M = min(M, a.M1);
    }
    bitonic_sort_u64(s, M);
    return M;
}
// orderable key bits -> the key as a real number
__device__ __forceinline__ double key_value(uint32_t ord, bool int_key) {
    return int_key ? (double)(int)(ord ^ 0x80000000u) : (double)f32_from_orderable(ord);
}

template <bool ASC, class RT = float>
__global__ void __launch_bounds__(1024) finalize_kernel(FinArgs a) {
    extern __shared__ __align__(16) unsigned char smem_fin[];
    uint64_t* s = reinterpret_cast<uint64_t*>(smem_fin);            // [M1] coarse keys
    uint64_t* e = s + a.M1;                                         // [R] exact keys
    float* sq = reinterpret_cast<float*>(e + a.R);                  // [dim_pad] query
    __shared__ uint32_t sh_T;      // orderable max of the floors
    __shared__ uint32_t sh_ncand, sh_overflow;
    const int q = blockIdx.x, tid = threadIdx.x;
    const int dim = a.dim;
    // lane-parallel rescoring needs 16-byte row pieces: dim % 4 == 0 for f32 rows, dim % 8 == 0 for binary16 rows
    const bool vec = (dim & (std::is_same<RT, float>::value ? 3 : 7)) == 0;
    const bool ik = a.int_key != 0;
    const RT* corpus = reinterpret_cast<const RT*>(a.corpus);
    if (tid == 0) { sh_T = 0; sh_ncand = 0; sh_overflow = 0; }
    for (int d = tid; d < dim; d += blockDim.x) sq[d] = a.queries[(size_t)q * dim + d];
    __syncthreads();
    gather_candidates(a, q, s, &sh_T, &sh_ncand, &sh_overflow);
    const int ncand = (int)sh_ncand;
    __shared__ int sh_done;
    // Rescore the candidates [lo, hi) of the coarse order into e[lo, hi) (KEY_NONE past the last candidate).
    auto rescore = [&](int lo, int hi) {
        const int rn = min(ncand, hi);
        if (vec) {
            // eight threads per row, four rows per warp, blockDim/8 rows per pass of the block (rows are independent:
            // warp-local syncs only)
            float* rowbuf = sq + ((dim + 3) & ~3) + (tid >> 3) * FIN_STRIDE;
            const int lane = tid & 31;
            const int rows_per_pass = (int)(blockDim.x >> 3);
            for (int base = lo; base < hi; base += rows_per_pass) {
                const int i = base + (tid >> 3);
                const bool row_ok = i < rn;
                const uint32_t row = row_ok ? key_row(s[i]) : 0u;
                const RT* c = corpus + (size_t)row * dim;
                const bool two = a.metric == LB_IP && row_ok && a.n_small > 0 && in_small_segment(a.small_seg, a.n_small, row);
                const float v = rescore_row_lanes(a.metric, two, sq, c, dim, rowbuf, row_ok, lane);
                if ((tid & 7) == 0 && i < hi) e[i] = row_ok ? make_key<ASC>(v, row) : KEY_NONE;
            }
        } else {
            for (int i = lo + tid; i < hi; i += blockDim.x) {
                uint64_t key = KEY_NONE;
                if (i < rn) {
                    uint32_t row = key_row(s[i]);
                    const RT* c = corpus + (size_t)row * dim;
                    const bool qvec = (dim & 3) == 0;
                    float v;
                    if (a.metric == LB_IP) {
                        bool small = a.n_small > 0 && in_small_segment(a.small_seg, a.n_small, row);
                        v = small ? ip_single_order<false>(sq, c, dim, qvec) : ip_batch8_order<false>(sq, c, dim, qvec);
                    } else if (a.metric == LB_L2) {
                        v = l2_squared<false>(sq, c, dim, qvec);
                    } else {
                        v = cosine_distance<false>(sq, c, dim, qvec);
                    }
                    key = make_key<ASC>(v, row);
                }
                e[i] = key;
            }
        }
    };
    // Thread 0: with the first `cut` candidates rescored and sorted in e, is the top k proven?  Every row that is not
    // among those candidates has a coarse key <= T: the largest partition floor (nothing is dropped above it inside a
    // partition) or the key of the first candidate left out.
    auto certify = [&](int cut) -> bool {
        uint32_t T_ord = sh_T;
        if (ncand > cut) T_ord = max(T_ord, ~(uint32_t)(s[cut] >> 32));
        const int kk = min(a.k, min(ncand, cut));
        if (sh_overflow) return false;                       // the hit buffer overflowed: candidates were lost
        if (T_ord <= (ik ? 0u : 0x007FFFFFu)) return true;   // no floor above the lowest key and no candidate left out: the candidates are the whole corpus
        if (kk < a.k) return false;
        // Bound on the exact "similarity" (q.c; cos; 2 q.c - |c|^2) of any dropped row, in real arithmetic:
        //   sim <= value(T) + |dq| cmax + |q~| emax + accumulation slop + rounding slop of the exact f32 score.
        // Every norm is measured (QStat, ShadowStats); the slack factors cover the f32 evaluation of those norms.
        const QStat qs = a.qstat[q];
        const double cmax = (double)__uint_as_float(a.sstat->cmax_bits) * 1.0001, emax = (double)__uint_as_float(a.sstat->emax_bits) * 1.0001;
        const double u24 = 5.9604644775390625e-8;
        const double qtn = (double)qs.qt_norm * 1.0001, qn = (double)qs.q_norm;
        const double dqn = (double)qs.dq_norm * 1.0001 + 4.0 * u24 * qtn;   // + rounding of the held values themselves
        double T = key_value(T_ord, ik);
        double eps = dqn * cmax + qtn * (emax + 4.0 * u24 * cmax);
        if (a.operand == OPERAND_U8) {
            // key = sum qhat chat exactly;  q~.c~ = s_q s_c key + s_q z_c sum(qhat)
            const double sq_ = (double)qs.s_q;
            T = sq_ * (double)a.c_scale * T + sq_ * (double)a.c_zero * (double)qs.sum_qhat;
            eps += 1e-9 * (fabs(T) + 1.0);
        } else {
            // tensor-core accumulation of the bf16 products in f32 (order unspecified): Dp * 2^-21 relative to |q~||c~|
            eps += (double)((dim + 63) & ~63) * 4.76837158203125e-7 * qtn * (cmax + emax);
        }
        const double worst = (double)key_score<ASC>(e[a.k - 1]);
        if (a.metric == LB_IP) {
            eps += (double)(dim / 8 + 16) * u24 * qn * cmax;                 // rounding of the exact f32 dot product
            return worst > T + eps;
        } else if (a.metric == LB_COSINE) {
            eps += (double)(dim / 4 + 64) * u24;                             // normalisations + the exact f32 cosine
            return worst < 1.0 - (T + eps);
        }
        // key = 2 q~.c~ - fl(|c|^2);  exact dist = |q - c|^2 >= |q|^2 - (T + eps)
        eps = 2.0 * eps + (double)(dim / 8 + 32) * u24 * 2.0 * (qn + cmax) * (qn + cmax);
        return worst < qn * qn * (1.0 - (double)(dim / 8 + 16) * u24) - (T + eps);
    };
    // Two rounds: most queries are already proven by the first half of the rescore budget (the rows it fetches are
    // what the kernel waits for); the rest take the second half and are judged on the whole budget as before.
    const int R1 = (a.R >= 64 && a.R1 > 0) ? a.R1 : a.R;
    int cut = R1;
    rescore(0, R1);
    bitonic_sort_u64(e, R1);
    if (tid == 0) sh_done = (R1 == a.R || ncand <= R1 || certify(R1)) ? 1 : 0;
    __syncthreads();
    bool certified = true;
    if (!sh_done) {
        cut = a.R;
        rescore(R1, a.R);
        bitonic_sort_u64(e, a.R);
        if (tid == 0) certified = certify(a.R);
    } else if (tid == 0 && (R1 == a.R || ncand <= R1)) {
        certified = certify(R1);   // (the short-circuit above skipped it)
    }
    const int kk = min(a.k, min(ncand, cut));
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
        if (!certified) {
            a.uncertified[q] = 1;
            atomicAdd(a.n_uncertified, 1u);
        } else {
            a.uncertified[q] = 0;
        }
    }
}

// ---- finalize for the binary metrics: candidates -> exact integer counts on the packed rows -> certified top-k ----
struct FinBitsArgs {
    FinArgs f;                 // candidate sources, k, metric, outputs (corpus / queries / stats unused)
    const uint64_t* words;     // packed rows [n][n_words]
    const uint64_t* qwords;    // packed queries [nq][n_words]
    int n_words;
};
static __global__ void __launch_bounds__(256) finalize_bits_kernel(FinBitsArgs b) {
    const FinArgs& a = b.f;
    extern __shared__ __align__(16) unsigned char smem_finb[];
    uint64_t* s = reinterpret_cast<uint64_t*>(smem_finb);           // [M1] coarse keys
    uint64_t* e = s + a.M1;                                         // [R] exact keys
    uint64_t* sqw = e + a.R;                                        // [n_words] query
    __shared__ uint32_t sh_T, sh_ncand, sh_overflow;
    const int q = blockIdx.x, tid = threadIdx.x, nw = b.n_words;
    const bool ik = a.int_key != 0;
    if (tid == 0) { sh_T = 0; sh_ncand = 0; sh_overflow = 0; }
    for (int w = tid; w < nw; w += blockDim.x) sqw[w] = b.qwords[(size_t)q * nw + w];
    __syncthreads();
    gather_candidates(a, q, s, &sh_T, &sh_ncand, &sh_overflow);
    const int ncand = (int)sh_ncand;
    const int rn = min(ncand, a.R);
    uint32_t T_ord = sh_T;
    if (ncand > a.R) T_ord = max(T_ord, ~(uint32_t)(s[a.R] >> 32));
    uint32_t qpop = 0;
    for (int w = 0; w < nw; ++w) qpop += __popcll(sqw[w]);
    for (int i = tid; i < a.R; i += blockDim.x) {
        uint64_t key = KEY_NONE;
        if (i < rn) {
            const uint32_t row = key_row(s[i]);
            const uint64_t* rw = b.words + (size_t)row * nw;
            uint32_t x = 0, inter = 0, rpop = 0;
            for (int w = 0; w < nw; ++w) {
                const uint64_t rv = __ldg(rw + w), qv = sqw[w];
                x += __popcll(rv ^ qv);
                inter += __popcll(rv & qv);
                rpop += __popcll(rv);
            }
            float d;
            if (a.metric == LB_HAMMING) d = packed_finish(LB_HAMMING, x, 0);
            else if (a.metric == LB_DICE) d = packed_finish(LB_DICE, inter, qpop + rpop);
            else d = packed_finish(LB_JACCARD, inter, qpop + rpop - inter);
            key = make_key<true>(d, row);
        }
        e[i] = key;
    }
    __syncthreads();
    bitonic_sort_u64(e, a.R);
    const int kk = min(a.k, rn);
    for (int i = tid; i < a.k; i += blockDim.x) {
        uint32_t row = ROW_NONE;
        float score = __int_as_float(0x7fc00000);
        if (i < kk) {
            row = key_row(e[i]);
            score = key_score<true>(e[i]);
        }
        a.out_rows[(size_t)q * a.k + i] = row;
        a.out_dists[(size_t)q * a.k + i] = score;
    }
    if (tid == 0) {
        a.out_counts[q] = kk;
        bool certified;
        if (sh_overflow) certified = false;
        else if (T_ord <= (ik ? 0u : 0x007FFFFFu)) certified = true;
        else if (kk < a.k) certified = false;
        else {
            const double worst = (double)key_score<true>(e[a.k - 1]);
            const double T = key_value(T_ord, ik);
            if (a.metric == LB_HAMMING) {
                certified = worst < (double)qpop - T;  // key = 2 inter - pb = pa - hamming, exact: ties are not certified
            } else {
                // key ~ r = inter / (pa + pb), off by the fast division's few ulps: a dropped row has r <= T (1 + 1e-6),
                // i.e. a distance of at least 1 - 2 r (Dice) or 1 - r / (1 - r) (Jaccard / Tanimoto) of that bound
                const double rb = fmin(T * (1.0 + 1e-6) + 1e-9, 0.5);
                const double dmin = a.metric == LB_DICE ? 1.0 - 2.0 * rb : (rb < 0.5 ? 1.0 - rb / (1.0 - rb) : 0.0);
                certified = worst < dmin - 2e-7;
            }
        }
        a.uncertified[q] = certified ? 0u : 1u;
        if (!certified) atomicAdd(a.n_uncertified, 1u);
    }
}

}  // namespace tc
}  // namespace lb
