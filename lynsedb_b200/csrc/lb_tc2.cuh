// lb_tc2.cuh — tensor-core coarse pass on CTA PAIRS (tcgen05 cta_group::2), for batches of two or more query tiles.
//
// The two CTAs of a cluster form one tcgen05 CTA pair:
//   * the MMA is M = 256 (128 queries in the TMEM of each CTA) x N = 64 corpus rows x K = 16, issued by ONE
//     thread of the even CTA for both SMs;
//   * each CTA stages only HALF of every corpus tile (its 32 rows: one 16 KiB box of the tiled shadow per stage)
//     and the tensor cores of both SMs read both halves, so the same 192 KiB ring holds 12 stages = four 768-dim
//     tiles per CTA;
//   * barriers: full[s] (even CTA only; both CTAs' TMA complete_tx on it), empty[s] / tfull[b] (both CTAs;
//     tcgen05.commit multicast), tempty[b] / aready (even CTA only; one RELAXED remote arrive per epilogue warp of
//     both CTAs — a release at cluster scope costs ~1400 cycles under a loaded memory system).
// Work: cluster c serves query group (c % n_mgroups) of slot (c / n_mgroups); the query groups of a slot stream the
// same row partitions in lockstep (TcArgs::progress / window) so HBM is read once per slot.
// Warp roles: 0 TMA producer, 1 MMA issuer (even CTA), 2..5 epilogue (2..9 with two epilogue sets); in the HELP_ shape
// 2..5 are scanners and 6..9 the helper warps that own the shortlists (see the comment at the kernel).
#pragma once
#include "lb_tc.cuh"

namespace lb {
namespace tc {

// Tile shapes.  BN_ = 64: each CTA stages one 32-row half block per K block (16 KiB stages, 12 of them), accumulators
// at TMEM columns [384, 512), any Dp <= 768.  BN_ = 128 (Dp <= 512, where the A operand leaves 256 columns free): each
// CTA stages a whole 64-row shadow tile per K block (even CTA: tile 2t, odd CTA: tile 2t+1; 32 KiB stages, 6 of them),
// accumulators at [256, 512); a tile is twice the tensor work for the same MMA -> epilogue -> MMA handshake, which is
// what bounds the pass when a tile is only a few hundred cycles of MMA (small dimensions).
// NACC_ = accumulator tiles in flight (TMEM columns [kDCol, 512)): 2 everywhere, 3 for BN_ = 128 with Dp <= 256, where
// the A operand leaves room.  The epilogue's cost per tile varies a lot at large k (a shortlist insertion is ~800
// cycles, and the MMA may only reuse a buffer when all eight epilogue warps of the pair have drained it): a third
// buffer lets the tensor pipe run two tiles ahead, so the pass pays each warp's AVERAGE epilogue time, not the
// slowest warp of every tile.
template <int BN_, int NACC_ = 2>
struct PairCfg {
    static constexpr int kRowsPerCta = BN_ / 2;
    static constexpr int kKbBytes = kRowsPerCta * 128;                  // one K block of this CTA's rows
    static constexpr int kStageBytes = KPS * kKbBytes;                  // 16 / 32 KiB
    static constexpr int kNStages = SMEM_RING_BYTES / kStageBytes;      // 12 / 6
    static constexpr int kNAcc = NACC_;
    static constexpr int kDCol = TMEM_COLS - NACC_ * BN_;               // 384 / 256 / 128
    static constexpr int kMaxDp = 2 * kDCol;                            // 768 / 512 / 256
};

__device__ __forceinline__ uint32_t mapa_rank(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    // relaxed: the arrive only orders tcgen05 traffic (tcgen05.fence::before_thread_sync precedes it); a release at
    // cluster scope would wait for every outstanding memory operation of the thread under a loaded memory system
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// `bar` is this CTA's own barrier address with the peer bit cleared: the bytes are counted on the barrier at that
// offset in the even CTA of the pair.
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_4d_pair(uint32_t smem_dst, const CUtensorMap* tmap, int c2, int c3, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(smem_dst), "l"(tmap), "r"(0), "r"(0), "r"(c2), "r"(c3), "r"(bar)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T over both CTAs of the pair: bf16 x bf16 -> f32 (K = 16) or 8-bit x 8-bit -> s32 (K = 32)
template <bool I8>
__device__ __forceinline__ void umma_pair_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    if (I8)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
            ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
            ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
}
__device__ __forceinline__ void umma_pair_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// EPI_ = sets of four epilogue warps per CTA.  With two sets (BN_ = 128 only) warps w and w + 4 share a TMEM
// sub-partition, i.e. the same 32 queries, and each takes one 64-column half of every accumulator tile into its own
// shortlists (TcArgs::lists_per_part = 2): the epilogue is bound by the issue rate of a single warp per sub-partition
// at large k, and two warps double it.
// MODE_ = CoarseMode: operand kind of the MMAs and how an accumulator becomes the key the epilogue ranks by.
// HITS_: the epilogue appends rows above a seeded floor to hit regions instead of keeping shortlists (TcArgs::hit_buf).
// HELP_ (list mode, one epilogue set, keys without side values, one partition per slot): four HELPER warps (6..9) own the
// shortlists.  The scanner warps (2..5) only read the accumulators, release them, compute the group maxima against the
// gate their helper last wrote, and pass the 16-row groups that hold a hit through a small shared-memory queue.  A
// list insertion is ~1500 cycles of dependent instructions on a warp that is alone on its scheduler; with the lists in
// the epilogue warp itself some warp of the pair's eight is in that path at nearly every tile of a short shard, and the
// MMAs (two accumulators in flight) wait for it: 2090 cycles per tile against 1536 of tensor work on a 1.25M-row shard
// of C2, exactly 1552 when the scan is switched off.  Scanner w and helper w + 4 share a scheduler, so the helper's
// latency-bound instruction stream fills issue slots the scanner leaves empty.
template <int BN_, int NACC_ = 2, int EPI_ = 1, int MODE_ = CM_F32, bool HITS_ = false, bool HELP_ = false>
__global__ void __launch_bounds__(64 + 128 * (EPI_ + (HELP_ ? 1 : 0)), 1)   // (warps are allocated four at a time: 320 threads are budgeted as 384, 168 registers)
coarse_pair_kernel(const __grid_constant__ CUtensorMap tmap_full, const __grid_constant__ CUtensorMap tmap_rem, TcArgs a) {
    static_assert(!HELP_ || (EPI_ == 1 && !HITS_ && BN_ == 128 && (MODE_ == CM_I32 || MODE_ == CM_F32)), "helper warps: list mode, one set, keys without side values");
    using Cfg = PairCfg<BN_, NACC_>;
    constexpr uint32_t NACC = NACC_;
    static_assert(EPI_ == 1 || (EPI_ == 2 && BN_ == 128), "two epilogue sets split a 128-column tile in halves");
    constexpr int P_NSTAGES = Cfg::kNStages;
    constexpr int P_STAGE_BYTES = Cfg::kStageBytes;
    using MT = ModeTraits<MODE_>;
    using Key = typename MT::Key;
    using KO = KeyOps<Key>;
    constexpr bool I8 = MT::kI8;
    constexpr int DCOL = Cfg::kDCol;   // shadows tc::DCOL
    constexpr int BN = BN_;            // shadows tc::BN
    // Pre-pass of the second-best exchange (helper-warp kernel): a partition's tile sequence is its first `pre` tiles
    // (group maxima only), then all of its tiles.  Producer, MMA issuer and scanners walk the same flattened sequence.
    auto pre_of = [&](uint32_t t0, uint32_t t1) -> uint32_t {
        return (HELP_ && a.pre_tiles > 0 && a.pbest2 != nullptr) ? min((uint32_t)a.pre_tiles, t1 - t0) : 0u;
    };
    const uint32_t crank = cluster_ctarank();  // 0 = even CTA (issues the MMAs), 1 = odd CTA
    const int n_mgroups = (a.n_mtiles + 1) / 2;
    const int cluster_id = (int)(blockIdx.x >> 1);
    const int mgroup = cluster_id % n_mgroups, slot = cluster_id / n_mgroups;
    const int n_rounds = slot < a.n_slots ? a.parts_per_slot : 0;
    constexpr int first_round = 0;
    extern __shared__ __align__(16) unsigned char smem_tc2[];
    const uint32_t smem_base = (smem_u32(smem_tc2) + 1023u) & ~1023u;
    unsigned char* smem = smem_tc2 + (smem_base - smem_u32(smem_tc2));
    const uint32_t bar_base = smem_base + SMEM_BAR_OFF;
    const uint32_t full0 = bar_base, empty0 = bar_base + 8u * P_NSTAGES, tfull0 = bar_base + 8u * (2 * P_NSTAGES),
                   tempty0 = tfull0 + 8u * NACC, aready_bar = tfull0 + 16u * NACC;
    static_assert(8 * (2 * P_NSTAGES + 2 * NACC_ + 2) <= 256, "barrier block overflows its 256 bytes");
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + SMEM_BAR_OFF + 8 * (2 * P_NSTAGES + 2 * NACC_ + 1));
    volatile uint32_t* abort_flag = reinterpret_cast<volatile uint32_t*>(smem + SMEM_BAR_OFF + 8 * (2 * P_NSTAGES + 2 * NACC_ + 1) + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long prof_c0 = clock64();
    const uint64_t prof_t0 = globaltimer_ns();
    if (threadIdx.x == 0) {
        for (int s = 0; s < P_NSTAGES; ++s) {
            mbar_init(full0 + 8u * s, 1);
            mbar_init(empty0 + 8u * s, 1);
        }
        for (uint32_t b = 0; b < NACC; ++b) {
            mbar_init(tfull0 + 8u * b, 1);
            mbar_init(tempty0 + 8u * b, 8 * EPI_);  // 4 * EPI_ epilogue warps x 2 CTAs
        }
        mbar_init(aready_bar, 8);
        *abort_flag = 0;
        fence_barrier_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_full) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_rem) : "memory");
    }
    if (warp == 0) {
        __syncwarp();
        tmem_alloc_pair(smem_u32(tmem_ptr_smem), TMEM_COLS);
        tmem_relinquish_pair();
    }
    if (HELP_ && warp >= 6) {   // queue barriers; every lane's gate and pre-pass floor at "nothing seen yet"
        const uint32_t cb = smem_base + SMEM_SCRATCH_OFF + 4 * HQ_SLOTS * HQ_ENTRY_WORDS * 4 + (uint32_t)(warp & 3) * HQ_CTRL_WORDS * 4;
        if (lane == 0) {
            for (uint32_t sI = 0; sI < HQ_SLOTS; ++sI) {
                mbar_init(cb + 32u + 8u * sI, 1);
                mbar_init(cb + 64u + 8u * sI, 1);
            }
            fence_barrier_init();
        }
        volatile uint32_t* gw = reinterpret_cast<volatile uint32_t*>(smem + SMEM_SCRATCH_OFF + 4 * HQ_SLOTS * HQ_ENTRY_WORDS * 4) + (warp & 3) * HQ_CTRL_WORDS + 32;
        gw[lane] = KO::bits(KO::lowest());
        gw[32 + lane] = KO::bits(KO::lowest());
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();  // both CTAs' barriers are initialised before anything remote can arrive
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int nkb = a.Dp / KBLK;
    const int n_full = nkb / KPS, rem = a.rem_kb;
    const int spt = n_full + (rem ? 1 : 0);  // stages per tile

    if (warp == 0) {
        // ===================== TMA producer (both CTAs; each loads its 32 rows of every tile) =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            bool ok = true;
            // lockstep: only the even CTA throttles; its peer follows through the shared stage ring
            bool lockstep = a.progress != nullptr && n_mgroups > 1 && crank == 0;
            uint32_t* prog = a.progress != nullptr ? a.progress + (size_t)slot * PROGRESS_STRIDE : nullptr;
            uint32_t seq = 0, known_min = 0;
            const uint32_t window = (uint32_t)a.window;
            const uint32_t pair_full0 = full0 & PEER_BIT_MASK;
            for (int r = first_round; r < n_rounds && ok; ++r) {
                const uint32_t part = (uint32_t)slot + (uint32_t)r * (uint32_t)a.n_slots;
                if (part >= (uint32_t)a.P) break;
                const uint32_t t0 = part * a.tiles_per_part;
                const uint32_t t1 = min(t0 + a.tiles_per_part, a.tiles_total);
                const uint32_t pre = pre_of(t0, t1), n_seq = pre + (t1 - t0);
                for (uint32_t i = 0; i < n_seq && ok; ++i) {
                    const uint32_t t = t0 + (i < pre ? i : i - pre);
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
                    for (int s = 0; s < spt; ++s) {
                        if (!mbar_wait(empty0 + 8u * stage, phase ^ 1u, abort_flag, 1)) { ok = false; break; }
                        const int kbc = s < n_full ? KPS : rem;
                        if ((a.debug_mode & 1) || ((a.debug_mode & 16) && seq >= 4)) {
                            // bit 4: real data for the first tiles only, then the stale stages are re-used without loads
                            if (crank == 0) mbar_arrive(full0 + 8u * stage);
                        } else {
                            // the pair's full barrier lives in the even CTA and counts the bytes landing in both CTAs
                            if (crank == 0) mbar_arrive_expect_tx(full0 + 8u * stage, 2u * (uint32_t)kbc * Cfg::kKbBytes);
                            if (BN_ == 64)   // my 32-row half of shadow tile t
                                tma_load_4d_pair(smem_base + stage * P_STAGE_BYTES, s < n_full ? &tmap_full : &tmap_rem, (int)crank,
                                                 (int)(t * (uint32_t)nkb) + s * KPS, pair_full0 + 8u * stage);
                            else             // the whole shadow tile 2t + crank (both halves)
                                tma_load_4d_pair(smem_base + stage * P_STAGE_BYTES, s < n_full ? &tmap_full : &tmap_rem, 0,
                                                 (int)((2u * t + crank) * (uint32_t)nkb) + s * KPS, pair_full0 + 8u * stage);
                        }
                        if (++stage == P_NSTAGES) { stage = 0; phase ^= 1u; }
                    }
                    ++seq;
                    if (prog != nullptr && crank == 0 && n_mgroups > 1) st_relaxed_gpu(prog + mgroup, seq);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: even CTA only, warp-uniform loops, one elected lane issues ============
        // Issue is paced by execution with little queue slack, so the next stage's barrier (and the next tile's
        // accumulator buffer) is probed with test_wait in the shadow of the current stage's MMAs.
        if (crank == 0) {
            const bool leader = elect_one();
            uint32_t stage = 0, phase = 0, tile_iter = 0, item_iter = 0;
            bool ok = true;
            uint32_t full_ready = 0, tempty_ready = 0;
            const uint64_t desc_base = make_b_desc(smem_base);
            const uint32_t P_IDESC = make_idesc<I8>(256, BN_) | (a.idesc_extra != nullptr ? __ldg(a.idesc_extra) : 0u);
            long long w_tempty = 0, w_full = 0, n_w_tempty = 0, n_w_full = 0;
            const long long mma_c0 = clock64();
            for (int r = first_round; r < n_rounds && ok; ++r, ++item_iter) {
                const uint32_t part = (uint32_t)slot + (uint32_t)r * (uint32_t)a.n_slots;
                if (part >= (uint32_t)a.P) break;
                const uint32_t t0 = part * a.tiles_per_part;
                const uint32_t t1 = min(t0 + a.tiles_per_part, a.tiles_total);
                if (!mbar_wait(aready_bar, item_iter & 1u, abort_flag, 2)) break;
                tcgen05_fence_after();
                const uint32_t n_seq = pre_of(t0, t1) + (t1 - t0);   // the MMAs do not depend on which tile it is
                for (uint32_t i = 0; i < n_seq && ok; ++i, ++tile_iter) {
                    const uint32_t buf = tile_iter % NACC;
                    if (!tempty_ready) {
                        const long long c0 = clock64();
                        if (!mbar_wait(tempty0 + 8u * buf, ((tile_iter / NACC) & 1u) ^ 1u, abort_flag, 3)) { ok = false; break; }
                        w_tempty += clock64() - c0;
                        ++n_w_tempty;
                    }
                    tempty_ready = 0;
                    tcgen05_fence_after();
                    const uint32_t d_tmem = tmem_base + DCOL + buf * BN;
#pragma unroll 1
                    for (int s = 0; s < spt; ++s) {
                        if (!full_ready) {
                            const long long c0 = clock64();
                            if (!mbar_wait(full0 + 8u * stage, phase, abort_flag, 4)) { ok = false; break; }
                            w_full += clock64() - c0;
                            ++n_w_full;
                        }
                        const int kbc = s < n_full ? KPS : rem;
                        const uint64_t bdesc0 = desc_base + (uint64_t)((stage * P_STAGE_BYTES) >> 4);
                        const uint32_t a0 = tmem_base + (uint32_t)(s * KPS * 4 * 8);
                        const int ks0 = s * KPS * 4;  // first K step of this stage; steps >= n_ksteps are zero padding
                        if (leader) {
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4)
                                if (ks0 + k4 < a.n_ksteps)
                                    umma_pair_ts<I8>(d_tmem, a0 + (uint32_t)(k4 * 8), bdesc0 + (uint64_t)(k4 * 2), P_IDESC,
                                                      (k4 == 0) ? (s > 0 ? 1u : 0u) : 1u);
                        }
                        {
                            uint32_t ns = stage + 1, nph = phase;
                            if (ns == P_NSTAGES) { ns = 0; nph ^= 1u; }
                            full_ready = mbar_test_wait(full0 + 8u * ns, nph);
                            if (s == spt - 1) {
                                const uint32_t nti = tile_iter + 1;
                                tempty_ready = mbar_test_wait(tempty0 + 8u * (nti % NACC), ((nti / NACC) & 1u) ^ 1u);
                            }
                        }
                        if (leader) {
#pragma unroll
                            for (int kb = 1; kb < KPS; ++kb) {
                                if (kb < kbc) {
#pragma unroll
                                    for (int k4 = 0; k4 < 4; ++k4)
                                        if (ks0 + kb * 4 + k4 < a.n_ksteps)
                                            umma_pair_ts<I8>(d_tmem, a0 + (uint32_t)((kb * 4 + k4) * 8),
                                                              bdesc0 + (uint64_t)(kb * (Cfg::kKbBytes >> 4) + k4 * 2), P_IDESC, 1u);
                                }
                            }
                            umma_pair_commit(empty0 + 8u * stage);  // frees the stage in both CTAs once these MMAs have read it
                        }
                        __syncwarp();
                        if (++stage == P_NSTAGES) { stage = 0; phase ^= 1u; }
                    }
                    if (ok && leader) umma_pair_commit(tfull0 + 8u * buf);  // accumulator tile complete (both CTAs)
                    __syncwarp();
                }
            }
            if (a.prof != nullptr && leader) {
                unsigned long long* pr = a.prof + (size_t)blockIdx.x * 8;
                pr[0] = (unsigned long long)(clock64() - mma_c0);
                pr[1] = (unsigned long long)w_tempty;
                pr[2] = (unsigned long long)w_full;
                pr[3] = (unsigned long long)n_w_tempty;
                pr[4] = (unsigned long long)n_w_full;
                pr[5] = tile_iter;
            }
        }
    } else if (HELP_) {
        // ===================== scanner warps 2..5 and helper warps 6..9 (see the kernel comment) =====================
        if constexpr (HELP_) {
            const int quad = warp & 3;
            const int ql = quad * 32 + lane;
            const uint32_t gq = ((uint32_t)mgroup * 2u + crank) * BM + (uint32_t)ql;
            const bool q_valid = gq < (uint32_t)a.nq;
            uint32_t* ent = reinterpret_cast<uint32_t*>(smem + SMEM_SCRATCH_OFF) + quad * (HQ_SLOTS * HQ_ENTRY_WORDS);
            uint32_t* ctrl = reinterpret_cast<uint32_t*>(smem + SMEM_SCRATCH_OFF + 4 * HQ_SLOTS * HQ_ENTRY_WORDS * 4) + quad * HQ_CTRL_WORDS;
            const uint32_t cb = smem_base + SMEM_SCRATCH_OFF + 4 * HQ_SLOTS * HQ_ENTRY_WORDS * 4 + (uint32_t)quad * HQ_CTRL_WORDS * 4;
            const uint32_t qfull0 = cb + 32u, qempty0 = cb + 64u;   // full[s] / empty[s] of this pair's queue
            volatile uint32_t* hg_gate = ctrl + 32 + lane;    // written by the helper, read by the scanner (unsynchronised on purpose)
            volatile uint32_t* hg_floor = ctrl + 64 + lane;   // written by the scanner, read by the helper
            const uint32_t part = (uint32_t)slot;   // one partition per slot
            const bool have_part = n_rounds > 0 && part < (uint32_t)a.P;
            const uint32_t t0 = part * a.tiles_per_part;
            const uint32_t t1 = have_part ? min(t0 + a.tiles_per_part, a.tiles_total) : t0;
            if (warp < 6) {
                // ---------- scanner: accumulators -> group maxima -> queue ----------
                const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
                const uint32_t even_tempty0 = mapa_rank(tempty0, 0), even_aready = mapa_rank(aready_bar, 0);
                long long e_wait = 0, e_ld = 0, e_scan = 0, e_slow = 0, n_slow = 0, e_max = 0;
                uint32_t tile_iter = 0, head = 0;
                bool ok = true;
                if (have_part) {
                    load_query_to_tmem(a.qb + (size_t)gq * a.Dp * 2, a.Dp, lane_addr);
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(even_aready);
                }
                // pre-pass state: this partition's two best keys so far, the floor the exchange yields at its end
                const uint32_t pre = have_part ? pre_of(t0, t1) : 0u, n_seq = pre + (t1 - t0);
                const bool exch = pre != 0u && q_valid && a.share_floor == 1 && a.P <= PBEST_STRIDE;
                const uint32_t* pb_row = a.pbest2 != nullptr ? a.pbest2 + (size_t)(q_valid ? gq : 0u) * PBEST_STRIDE : nullptr;
                Key best1 = KO::lowest(), best2 = KO::lowest(), pub2 = KO::lowest(), floor0 = KO::lowest();
                for (uint32_t i = 0; i < n_seq && ok; ++i, ++tile_iter) {
                    const uint32_t t = t0 + (i < pre ? i : i - pre);
                    const bool pre_mode = i < pre;
                    const bool whole = (t + 1u) * BN <= a.n_rows;   // no padding rows among the group maxima
                    if (pre != 0u && i == pre) {
                        // every partition of the query has published its second best (they run side by side: bounded wait;
                        // a late one only means this pass starts without a floor, as it does without a pre-pass)
                        for (uint32_t it = 0; it < 1024u; ++it) {
                            bool done = true;
                            if (exch) {
                                uint4 w4[PBEST_STRIDE / 4];   // the query's row of the exchange: five independent 16-byte loads
#pragma unroll
                                for (int c = 0; c < PBEST_STRIDE / 4; ++c)
                                    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                                                 : "=r"(w4[c].x), "=r"(w4[c].y), "=r"(w4[c].z), "=r"(w4[c].w)
                                                 : "l"(pb_row + 4 * c));
                                uint32_t mn = 0xFFFFFFFFu;
#pragma unroll
                                for (int c = 0; c < PBEST_STRIDE / 4; ++c) {
                                    const uint32_t w[4] = {w4[c].x, w4[c].y, w4[c].z, w4[c].w};
#pragma unroll
                                    for (int e = 0; e < 4; ++e) mn = min(mn, 4 * c + e < a.P ? w[e] : 0xFFFFFFFFu);
                                }
                                done = mn != 0u;
                                if (done) floor0 = KO::from_orderable(mn);
                            }
                            if (__all_sync(0xffffffffu, done)) break;
                            __nanosleep(256);
                        }
                        // the helper folds it into its floor: every row this warp drops from now on scored <= max(gate, floor0),
                        // and the list's final floor (cand_thr) must cover that
                        *hg_floor = KO::bits(floor0);
                    }
                    const uint32_t buf = tile_iter % NACC;
                    const Key thr = pre_mode ? KO::highest() : (q_valid && !(a.debug_mode & 4) ? max(KO::from_bits(*hg_gate), floor0) : KO::highest());
                    const long long ec0 = clock64();
                    if (!mbar_wait(tfull0 + 8u * buf, (tile_iter / NACC) & 1u, abort_flag, 5)) { ok = false; break; }
                    const long long ec1 = clock64();
                    tcgen05_fence_after();
                    uint32_t v[64];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        tmem_ld_32x32b_x64(lane_addr + DCOL + buf * BN + h * 64, v);
                        tmem_ld_wait();
                        if (h == 1) {
                            tcgen05_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_cluster(even_tempty0 + 8u * buf);  // accumulator is in registers
                            e_wait += ec1 - ec0;
                            e_ld += clock64() - ec1;
                        }
                        const long long sc0 = clock64();
                        // group maxima with 3-input max, ONE vote for the half (as the epilogue warps' fast path does): the
                        // per-group votes only run in a half that holds a hit
                        Key g[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t* x = v + 16 * j;
                            Key m = KO::max3(KO::from_bits(x[0]), KO::from_bits(x[1]), KO::from_bits(x[2]));
#pragma unroll
                            for (int i = 3; i + 1 < 16; i += 2) m = KO::max3(m, KO::from_bits(x[i]), KO::from_bits(x[i + 1]));
                            g[j] = max(m, KO::from_bits(x[15]));
                        }
                        if (pre_mode) {   // one row of this partition per group of corpus rows: its two best keys so far
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                if (whole || t * BN + (uint32_t)(h * 64 + j * 16 + 16) <= a.n_rows) {
                                    const Key lo = min(g[j], best1);
                                    best1 = max(g[j], best1);
                                    best2 = max(best2, lo);
                                }
                        }
                        const Key gall = max(max(g[0], g[1]), max(g[2], g[3]));
                        if (__any_sync(0xffffffffu, gall > thr)) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                if (!ok || !__any_sync(0xffffffffu, g[j] > thr)) continue;
                                const uint32_t* x = v + 16 * j;
                                // back-pressure: the slot this entry goes to must have been consumed
                                const uint32_t qs = head % HQ_SLOTS;
                                if (!mbar_wait(qempty0 + 8u * qs, ((head / HQ_SLOTS) & 1u) ^ 1u, abort_flag, 6)) { ok = false; continue; }
                                uint32_t* e = ent + qs * HQ_ENTRY_WORDS;
#pragma unroll
                                for (int i = 0; i < 16; ++i) e[i * 32 + lane] = x[i];
                                if (lane == 0) ctrl[qs] = t * BN + (uint32_t)(h * 64 + j * 16);
                                __syncwarp();
                                if (lane == 0) mbar_arrive(qfull0 + 8u * qs);
                                ++head;
                            }
                        }
                        const long long sd = clock64() - sc0;
                        e_scan += sd;
                        if (sd > 400) { ++n_slow; e_slow += sd; }
                        if (sd > e_max) e_max = sd;
                        if (!ok) break;
                    }
                    if (pre_mode && exch && best2 > pub2) {
                        st_relaxed_gpu(a.pbest2 + (size_t)gq * PBEST_STRIDE + part, KO::orderable(best2));  // single writer until the helper's lists take over
                        pub2 = best2;
                    }
                    if (pre_mode && exch && i + 1u == pre && !(pub2 > KO::lowest())) {
                        // a partition of a single ragged tile with fewer than two whole groups: publish "no floor" rather than
                        // let the other partitions of the query wait out their bound
                        Key none;
                        if constexpr (MT::kIntKey) none = KO::lowest() + 1;
                        else none = -3.0e38f;
                        st_relaxed_gpu(a.pbest2 + (size_t)gq * PBEST_STRIDE + part, KO::orderable(none));
                    }
                }
                // end of the stream (also after a failed wait: the helper must not be left waiting)
                {
                    const uint32_t qs = head % HQ_SLOTS;
                    if (*abort_flag == 0u && mbar_wait(qempty0 + 8u * qs, ((head / HQ_SLOTS) & 1u) ^ 1u, abort_flag, 6)) {
                        if (lane == 0) ctrl[qs] = HQ_END;
                        __syncwarp();
                        if (lane == 0) mbar_arrive(qfull0 + 8u * qs);
                    }
                }
                if (a.prof != nullptr && warp == 2 && lane == 0) {
                    unsigned long long* pr = a.prof + (size_t)blockIdx.x * 8;
                    pr[6] = (unsigned long long)e_wait;
                    pr[7] = (unsigned long long)e_ld;
                    if (crank == 1) {
                        pr[0] = (unsigned long long)e_scan;
                        pr[1] = (unsigned long long)e_slow;
                        pr[2] = (unsigned long long)n_slow;
                        pr[3] = (unsigned long long)e_max;
                        pr[4] = tile_iter;
                    }
                }
            } else {
                // ---------- helper: queue -> shortlists, floors, the gate the scanner tests against ----------
                Shortlist<MODE_, false> sl;
                const float qaux = a.qaux != nullptr ? __ldg(a.qaux + gq) : 0.0f;
                sl.init_floor(qaux);
                if (have_part) {
                    sl.reset(q_valid, a, gq, part, 0u);
                    *hg_gate = KO::bits(sl.gate());
                }
                const bool have_allow = a.allow_bits != nullptr;
                uint32_t tail = 0;
                long long h_busy = 0;   // cycles between taking a group off the queue and being ready for the next (diagnostics)
                while (have_part) {
                    const uint32_t qs = tail % HQ_SLOTS;
                    if (!mbar_wait(qfull0 + 8u * qs, (tail / HQ_SLOTS) & 1u, abort_flag, 7)) break;   // (an aborted launch ends here)
                    const uint32_t* e = ent + qs * HQ_ENTRY_WORDS;
                    const uint32_t row0 = ctrl[qs];
                    if (row0 == HQ_END) break;
                    uint32_t w[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) w[i] = e[i * 32 + lane];
                    __syncwarp();
                    if (lane == 0) mbar_arrive(qempty0 + 8u * qs);   // the group is in registers: the slot is free
                    ++tail;
                    const long long hc0 = clock64();
                    if (!(a.debug_mode & 128)) sl.poll_floor(tail, (uint32_t)a.poll_mask >> 1);   // every 4th entry (tiles: every 8th)
                    sl.thr_g = max(sl.thr_g, KO::from_bits(*hg_floor));   // the scanner's pre-pass floor
                    const Key thr = sl.gate();
                    sl.fn.set_gate(thr);
                    const Key tthr = sl.fn.test_thr(thr);
                    Key gmax = KO::max3(KO::from_bits(w[0]), KO::from_bits(w[1]), KO::from_bits(w[2]));
#pragma unroll
                    for (int i = 3; i + 1 < 16; i += 2) gmax = KO::max3(gmax, KO::from_bits(w[i]), KO::from_bits(w[i + 1]));
                    gmax = max(gmax, KO::from_bits(w[15]));
                    if (__any_sync(0xffffffffu, gmax > tthr)) {
                        uint32_t allow16 = 0xffffu;
                        if (have_allow) allow16 = row0 < a.n_rows ? (uint32_t)(__ldg(a.allow_bits + (row0 >> 6)) >> (row0 & 63u)) & 0xffffu : 0u;
                        sl.slow16(w, nullptr, gmax, row0, a.n_rows, thr, tthr, allow16, have_allow);
                        if (sl.may_publish && sl.q_valid && sl.lmin > sl.thr_pub && sl.lmin > sl.thr_g) sl.publish();
                    }
                    *hg_gate = KO::bits(sl.gate());
                    h_busy += clock64() - hc0;
                }
                if (a.prof != nullptr && warp == 6 && lane == 0 && crank == 1)   // odd CTA, slot 5: busy cycles << 24 | groups taken
                    a.prof[(size_t)blockIdx.x * 8 + 5] = ((unsigned long long)h_busy << 24) | (unsigned long long)(tail & 0xFFFFFFu);
                if (have_part) {
                    sl.thr_g = max(sl.thr_g, KO::from_bits(*hg_floor));   // (no entry may have arrived since it was set)
                    sl.flush(a, gq, part, 0u);
                }
            }
        }
    } else {
        // ===================== epilogue: lane == query (each CTA reads its own 128 accumulator lanes) ==========
        const int quad = warp & 3;                 // TMEM sub-partition of this warp = the 32 queries it serves
        const int eset = EPI_ == 2 ? (warp - 2) >> 2 : 0;  // which 64-column half of a tile this warp scans (two sets)
        const int ql = quad * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
        const uint32_t even_tempty0 = mapa_rank(tempty0, 0), even_aready = mapa_rank(aready_bar, 0);
        Shortlist<MODE_, HITS_> sl;
        uint32_t tile_iter = 0;
        bool ok = true;
        long long e_wait = 0, e_ld = 0, e_scan = 0, e_slow = 0, n_slow = 0, e_max = 0;
        const uint32_t gq = ((uint32_t)mgroup * 2u + crank) * BM + (uint32_t)ql;
        const bool q_valid = gq < (uint32_t)a.nq;
        const float qaux = a.qaux != nullptr ? __ldg(a.qaux + gq) : 0.0f;
        // Side values of the rows being scanned: cp.async moves them global -> this warp's shared-memory ring two tiles
        // ahead of their use (lane l carries rows 2l, 2l+1 of each 64-row half this warp scans), so no register and no
        // instruction of the tile loop ever waits for the load; the scan reads them back as broadcasts.
        constexpr int HPW = EPI_ == 2 ? 1 : BN_ / 64;  // halves per warp
        uint32_t* scratch = reinterpret_cast<uint32_t*>(smem + SMEM_SCRATCH_OFF) + (warp - 2) * (EPI_SCRATCH_SLOTS * EPI_SCRATCH_WORDS / EPI_);
        constexpr uint32_t SLOT_WORDS = EPI_SCRATCH_WORDS / EPI_;  // 128 (four warps: two halves) or 64 (eight warps: one half)
        auto side_fetch = [&](uint32_t t, uint32_t slot) {
            if (MT::kBias) {
#pragma unroll
                for (int hh = 0; hh < HPW; ++hh) {
                    const int h = EPI_ == 2 ? eset : hh;
                    cp_async_8(smem_u32(scratch + slot * SLOT_WORDS + hh * 64 + 2 * lane), a.bias + (size_t)t * BN + h * 64 + 2 * lane);
                }
            }
            cp_async_commit();
        };
        sl.init_floor(qaux);
        for (int r = first_round; r < n_rounds && ok; ++r) {
            const uint32_t part = (uint32_t)slot + (uint32_t)r * (uint32_t)a.n_slots;
            if (part >= (uint32_t)a.P) break;
            const uint32_t t0 = part * a.tiles_per_part;
            const uint32_t t1 = min(t0 + a.tiles_per_part, a.tiles_total);
            if (r == first_round && eset == 0) load_query_to_tmem(a.qb + (size_t)gq * a.Dp * 2, a.Dp, lane_addr);  // the query tile never changes
            sl.reset(q_valid, a, gq, part, (uint32_t)eset);
            // ring: tile t uses slot (t - t0) % 3; tiles t0 and t0 + 1 are in flight before the loop, tile t + 2 is fetched in
            // iteration t (one commit group per iteration, so wait_group<2> always means "tile t has landed")
            if (MT::kBias) {
                __syncwarp();
                side_fetch(t0, 0);
                side_fetch(min(t0 + 1, t1 - 1), 1);
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0 && eset == 0) mbar_arrive_cluster(even_aready);
            if constexpr (!MT::kBias && EPI_ == 2) {
                // Pipelined drain (two epilogue sets, modes without side values — the short tiles of narrow rows, where the
                // TMEM read latency, ~200 cycles per 64 columns, is a third of a tile's budget): this warp's 64 columns of a
                // tile come out of TMEM as two units of 32, A in v[0, 32) and B in v[32, 64).  B's load is in flight under
                // the test of A; the tile's buffer is released as soon as B has landed — before any slow path, so a warp
                // with a hit never holds the accumulator ring up —; the next tile's A is in flight under the whole of B.
                const uint32_t col0 = (uint32_t)eset * 64u;
                const bool scan_disabled = (a.debug_mode & 4) != 0;
                using SL = Shortlist<MODE_, HITS_>;
                uint32_t v[64];
                auto acquire = [&](uint32_t ti) -> bool {   // tile (running index ti) complete in TMEM?
                    const long long ec0 = clock64();
                    if (!mbar_wait(tfull0 + 8u * (ti % NACC), (ti / NACC) & 1u, abort_flag, 5)) return false;
                    e_wait += clock64() - ec0;
                    tcgen05_fence_after();
                    return true;
                };
                auto unit_addr = [&](uint32_t ti, uint32_t uu) { return lane_addr + DCOL + (ti % NACC) * BN + col0 + uu * 32u; };
                auto dump_unit = [&](const uint32_t* src, uint32_t row0) {
                    if (a.dump != nullptr) {
                        float* drow = a.dump + (size_t)gq * ((size_t)a.tiles_total * BN) + row0;
                        for (int i = 0; i < 32; ++i) drow[i] = KO::as_f32(sl.fn.key(src[i], 0u));
                    }
                };
                ok = t1 == t0 || acquire(tile_iter);
                if (ok && t1 > t0) tmem_ld_32x32b_x32(unit_addr(tile_iter, 0u), v);
                for (uint32_t t = t0; t < t1 && ok; ++t, ++tile_iter) {
                    if (!(a.debug_mode & 128)) sl.poll_floor(tile_iter, (uint32_t)a.poll_mask);
                    const uint32_t rowA = t * BN + col0, rowB = rowA + 32u;
                    const long long lc0 = clock64();
                    tmem_ld_wait_32(v);                                            // A has landed
                    tmem_ld_32x32b_x32(unit_addr(tile_iter, 1u), v + 32);          // B on its way
                    const long long sc0 = clock64();
                    typename SL::template ScanState<2> stA, stB;
                    dump_unit(v, rowA);
                    const bool hitA = sl.template scan_fast<2>(v, nullptr, scan_disabled, stA);
                    const long long sc1 = clock64();
                    tmem_ld_wait_32(v + 32);                                       // B has landed: the tile is in registers
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(even_tempty0 + 8u * (tile_iter % NACC));
                    const long long sc2 = clock64();
                    e_ld += (sc0 - lc0) + (sc2 - sc1);
                    if (hitA) sl.template scan_slow<2>(v, nullptr, rowA, a.n_rows, a.allow_bits, stA);
                    if (t + 1 < t1) {                                              // next tile's A under the scan of B
                        ok = acquire(tile_iter + 1u);
                        if (ok) tmem_ld_32x32b_x32(unit_addr(tile_iter + 1u, 0u), v);
                    }
                    dump_unit(v + 32, rowB);
                    if (sl.template scan_fast<2>(v + 32, nullptr, scan_disabled, stB)) sl.template scan_slow<2>(v + 32, nullptr, rowB, a.n_rows, a.allow_bits, stB);
                    const long long sd = (sc1 - sc0) + (clock64() - sc2);
                    e_scan += sd;
                    if (sd > 400) { ++n_slow; e_slow += sd; }
                    if (sd > e_max) e_max = sd;
                }
            } else {
                for (uint32_t t = t0; t < t1; ++t, ++tile_iter) {
                    const uint32_t buf = tile_iter % NACC;
                    if (!(a.debug_mode & 128)) sl.poll_floor(tile_iter, (uint32_t)a.poll_mask);
                    if (MT::kBias) side_fetch(min(t + 2, t1 - 1), (t - t0 + 2) % EPI_SCRATCH_SLOTS);
                    const long long ec0 = clock64();
                    if (!mbar_wait(tfull0 + 8u * buf, (tile_iter / NACC) & 1u, abort_flag, 5)) { ok = false; break; }
                    const long long ec1 = clock64();
                    tcgen05_fence_after();
                    uint32_t v[64];
#pragma unroll
                    for (int h = 0; h < BN_ / 64; ++h) {
                        if (EPI_ == 2 && h != eset) continue;  // the other set's half
                        if (!(a.debug_mode & 2)) {
                            tmem_ld_32x32b_x64(lane_addr + DCOL + buf * BN + h * 64, v);
                            tmem_ld_wait();
                        }
                        if (EPI_ == 2 || h == BN_ / 64 - 1) {
                            tcgen05_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_cluster(even_tempty0 + 8u * buf);  // accumulator is in registers
                            e_wait += ec1 - ec0;
                            e_ld += clock64() - ec1;
                        }
                        if (a.debug_mode & 2) continue;
                        const uint32_t row0 = t * BN + h * 64;
                        const long long sc0 = clock64();
                        if (MT::kBias && (EPI_ == 2 || h == 0)) {
                            cp_async_wait<2>();  // this tile's side values have landed (the two younger groups may be in flight)
                            __syncwarp();
                        }
                        const uint32_t* side = scratch + ((t - t0) % EPI_SCRATCH_SLOTS) * SLOT_WORDS + (EPI_ == 2 ? 0 : h * 64);
                        if (a.dump != nullptr) {
                            float* drow = a.dump + (size_t)gq * ((size_t)a.tiles_total * BN) + row0;
                            for (int i = 0; i < 64; ++i) drow[i] = KO::as_f32(sl.fn.key(v[i], MT::kBias ? side[i] : 0u));
                        }
                        sl.scan64(v, side, row0, a.n_rows, (a.debug_mode & 4) != 0, a.allow_bits);
                        const long long sd = clock64() - sc0;
                        e_scan += sd;
                        if (sd > 400) { ++n_slow; e_slow += sd; }
                        if (sd > e_max) e_max = sd;
                    }
                    if (MT::kBias) __syncwarp();  // every lane is done reading this tile's slot before a later fetch reuses it
                }
            }
            if (MT::kBias) cp_async_wait<0>();
            if (ok) sl.flush(a, gq, part, (uint32_t)eset);
        }
        if (a.prof != nullptr && warp == 2 && lane == 0) {
            unsigned long long* pr = a.prof + (size_t)blockIdx.x * 8;
            pr[6] = (unsigned long long)e_wait;
            pr[7] = (unsigned long long)e_ld;
            if (crank == 1) {  // the odd CTA has no MMA issuer: its slots 0..3 carry the scan statistics of one warp
                pr[0] = (unsigned long long)e_scan;
                pr[1] = (unsigned long long)e_slow;
                pr[2] = (unsigned long long)n_slow;
                pr[3] = (unsigned long long)e_max;
                pr[4] = tile_iter;
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();  // no CTA leaves (or frees TMEM) while the pair may still touch its memory or barriers
    tcgen05_fence_after();
    if (threadIdx.x == 0 && *abort_flag) atomicMax(a.error_flag, *abort_flag);
    if (threadIdx.x == 0 && blockIdx.x == 0) {  // SM clock actually seen by this launch: cycles and nanoseconds of CTA 0
        a.error_flag[2] = (uint32_t)((clock64() - prof_c0) >> 4);
        a.error_flag[3] = (uint32_t)(globaltimer_ns() - prof_t0);
    }
    if (warp == 0) tmem_dealloc_pair(tmem_base, TMEM_COLS);
}

}  // namespace tc
}  // namespace lb
