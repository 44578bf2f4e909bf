// lb_tc1.cuh — tensor-core coarse pass, one CTA per 128 queries (used when a batch has a single query tile).
//
// Persistent CTAs, 192 threads:
//   warp 0      TMA producer: one 32 KiB box of the tiled shadow per stage (4 K blocks x 64 rows), 6-stage ring
//   warp 1      MMA issuer: tcgen05.mma.cta_group::1.kind::f16, M=128 (queries, A in TMEM), N=64 (rows), K=16
//   warps 2..5  epilogue: lane == query, Shortlist (lb_tc.cuh)
// Work: CTA c owns slot c and walks the row partitions c, c + n_slots, ...  With a single query tile the pass is
// HBM-bound (each shadow byte is used by 128 queries), so what matters here is bytes in flight: 148 x 192 KiB.
#pragma once
#include "lb_tc.cuh"

namespace lb {
namespace tc {

constexpr int S_STAGE_BYTES = KPS * 2 * HALF_BLOCK_BYTES;  // 32 KiB
constexpr int S_NSTAGES = SMEM_RING_BYTES / S_STAGE_BYTES;  // 6

template <bool I8>
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    if (I8) umma_ts_i8(d_tmem, a_tmem, b_desc, idesc, accumulate);
    else umma_ts_bf16(d_tmem, a_tmem, b_desc, idesc, accumulate);
}

template <int MODE_ = CM_F32, bool HITS_ = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
coarse_single_kernel(const __grid_constant__ CUtensorMap tmap_full, const __grid_constant__ CUtensorMap tmap_rem, TcArgs a) {
    const int slot = (int)blockIdx.x;
    const int n_rounds = slot < a.n_slots ? a.parts_per_slot : 0;
    constexpr int first_round = 0;
    using MT = ModeTraits<MODE_>;
    using Key = typename MT::Key;
    using KO = KeyOps<Key>;
    constexpr bool I8 = MT::kI8;
    extern __shared__ __align__(16) unsigned char smem_tc1[];
    const uint32_t smem_base = (smem_u32(smem_tc1) + 1023u) & ~1023u;
    unsigned char* smem = smem_tc1 + (smem_base - smem_u32(smem_tc1));
    const uint32_t bar_base = smem_base + SMEM_BAR_OFF;
    const uint32_t full0 = bar_base, empty0 = bar_base + 8u * S_NSTAGES, tfull0 = bar_base + 8u * (2 * S_NSTAGES),
                   tempty0 = tfull0 + 16u, aready_bar = tfull0 + 32u;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + SMEM_BAR_OFF + 8 * (2 * S_NSTAGES + 5));
    volatile uint32_t* abort_flag = reinterpret_cast<volatile uint32_t*>(smem + SMEM_BAR_OFF + 8 * (2 * S_NSTAGES + 5) + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S_NSTAGES; ++s) {
            mbar_init(full0 + 8u * s, 1);
            mbar_init(empty0 + 8u * s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tfull0 + 8u * b, 1);
            mbar_init(tempty0 + 8u * b, 4);  // one arrive per epilogue warp
        }
        mbar_init(aready_bar, 4);
        *abort_flag = 0;
        fence_barrier_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_full) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_rem) : "memory");
    }
    if (warp == 0) {
        __syncwarp();
        tmem_alloc(smem_u32(tmem_ptr_smem), TMEM_COLS);
        tmem_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int nkb = a.Dp / KBLK;
    const int n_full = nkb / KPS, rem = a.rem_kb;
    const int spt = n_full + (rem ? 1 : 0);

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            bool ok = true;
            for (int r = first_round; r < n_rounds && ok; ++r) {
                const uint32_t part = (uint32_t)slot + (uint32_t)r * (uint32_t)a.n_slots;
                if (part >= (uint32_t)a.P) break;
                const uint32_t t0 = part * a.tiles_per_part;
                const uint32_t t1 = min(t0 + a.tiles_per_part, a.tiles_total);
                for (uint32_t t = t0; t < t1 && ok; ++t) {
                    for (int s = 0; s < spt; ++s) {
                        if (!mbar_wait(empty0 + 8u * stage, phase ^ 1u, abort_flag, 1)) { ok = false; break; }
                        const int kbc = s < n_full ? KPS : rem;
                        if (a.debug_mode & 1) {
                            mbar_arrive(full0 + 8u * stage);
                        } else {
                            mbar_arrive_expect_tx(full0 + 8u * stage, (uint32_t)kbc * 2u * HALF_BLOCK_BYTES);
                            tma_load_4d(smem_base + stage * S_STAGE_BYTES, s < n_full ? &tmap_full : &tmap_rem, 0,
                                        (int)(t * (uint32_t)nkb) + s * KPS, full0 + 8u * stage);
                        }
                        if (++stage == S_NSTAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform loops; one elected lane issues) =====================
        // The next stage's barrier and the next tile's accumulator buffer are probed (test_wait, non-blocking) in
        // the shadow of the current stage's MMAs; the blocking wait is only taken when the probe failed.
        const bool leader = elect_one();
        uint32_t stage = 0, phase = 0, tile_iter = 0, item_iter = 0;
        bool ok = true;
        uint32_t full_ready = 0, tempty_ready = 0;
        const uint64_t desc_base = make_b_desc(smem_base);
        const uint32_t S_IDESC = make_idesc<I8>(BM, BN) | (a.idesc_extra != nullptr ? __ldg(a.idesc_extra) : 0u);
        for (int r = first_round; r < n_rounds && ok; ++r, ++item_iter) {
            const uint32_t part = (uint32_t)slot + (uint32_t)r * (uint32_t)a.n_slots;
            if (part >= (uint32_t)a.P) break;
            const uint32_t t0 = part * a.tiles_per_part;
            const uint32_t t1 = min(t0 + a.tiles_per_part, a.tiles_total);
            if (!mbar_wait(aready_bar, item_iter & 1u, abort_flag, 2)) break;
            tcgen05_fence_after();
            for (uint32_t t = t0; t < t1 && ok; ++t, ++tile_iter) {
                const uint32_t buf = tile_iter & 1u;
                if (!tempty_ready && !mbar_wait(tempty0 + 8u * buf, ((tile_iter >> 1) & 1u) ^ 1u, abort_flag, 3)) { ok = false; break; }
                tempty_ready = 0;
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + DCOL + buf * BN;
#pragma unroll 1
                for (int s = 0; s < spt; ++s) {
                    if (!full_ready && !mbar_wait(full0 + 8u * stage, phase, abort_flag, 4)) { ok = false; break; }
                    const int kbc = s < n_full ? KPS : rem;
                    const uint64_t bdesc0 = desc_base + (uint64_t)((stage * S_STAGE_BYTES) >> 4);
                    const uint32_t a0 = tmem_base + (uint32_t)(s * KPS * 4 * 8);
                    const int ks0 = s * KPS * 4;  // first K step of this stage; steps >= n_ksteps are zero padding
                    if (leader) {
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4)
                            if (ks0 + k4 < a.n_ksteps)
                                umma_ts<I8>(d_tmem, a0 + (uint32_t)(k4 * 8), bdesc0 + (uint64_t)(k4 * 2), S_IDESC,
                                             (k4 == 0) ? (s > 0 ? 1u : 0u) : 1u);
                    }
                    {
                        uint32_t ns = stage + 1, nph = phase;
                        if (ns == S_NSTAGES) { ns = 0; nph ^= 1u; }
                        full_ready = mbar_test_wait(full0 + 8u * ns, nph);
                        if (s == spt - 1) {
                            const uint32_t nti = tile_iter + 1;
                            tempty_ready = mbar_test_wait(tempty0 + 8u * (nti & 1u), ((nti >> 1) & 1u) ^ 1u);
                        }
                    }
                    if (leader) {
#pragma unroll
                        for (int kb = 1; kb < KPS; ++kb) {
                            if (kb < kbc) {
#pragma unroll
                                for (int k4 = 0; k4 < 4; ++k4)
                                    if (ks0 + kb * 4 + k4 < a.n_ksteps)
                                        umma_ts<I8>(d_tmem, a0 + (uint32_t)((kb * 4 + k4) * 8),
                                                     bdesc0 + (uint64_t)(kb * ((2 * HALF_BLOCK_BYTES) >> 4) + k4 * 2), S_IDESC, 1u);
                            }
                        }
                        umma_commit(empty0 + 8u * stage);  // frees the stage once these MMAs have read it
                    }
                    __syncwarp();
                    if (++stage == S_NSTAGES) { stage = 0; phase ^= 1u; }
                }
                if (ok && leader) umma_commit(tfull0 + 8u * buf);  // accumulator tile complete
                __syncwarp();
            }
        }
    } else {
        // ===================== epilogue: lane == query =====================
        const int quad = warp & 3;                    // TMEM lane quadrant this warp may access
        const int ql = quad * 32 + lane;              // query within the tile
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
        Shortlist<MODE_, HITS_> sl;
        uint32_t tile_iter = 0;
        bool ok = true;
        const uint32_t gq = (uint32_t)ql;
        const bool q_valid = gq < (uint32_t)a.nq;
        const float qaux = a.qaux != nullptr ? __ldg(a.qaux + gq) : 0.0f;
        // side values of the rows being scanned: cp.async ring, two tiles ahead (see lb_tc2.cuh)
        uint32_t* scratch = reinterpret_cast<uint32_t*>(smem + SMEM_SCRATCH_OFF) + (warp - 2) * (EPI_SCRATCH_SLOTS * EPI_SCRATCH_WORDS);
        auto side_fetch = [&](uint32_t t, uint32_t slot) {
            if (MT::kBias) cp_async_8(smem_u32(scratch + slot * EPI_SCRATCH_WORDS + 2 * lane), a.bias + (size_t)t * BN + 2 * lane);
            cp_async_commit();
        };
        sl.init_floor(qaux);
        for (int r = first_round; r < n_rounds && ok; ++r) {
            const uint32_t part = (uint32_t)slot + (uint32_t)r * (uint32_t)a.n_slots;
            if (part >= (uint32_t)a.P) break;
            const uint32_t t0 = part * a.tiles_per_part;
            const uint32_t t1 = min(t0 + a.tiles_per_part, a.tiles_total);
            if (r == first_round) load_query_to_tmem(a.qb + (size_t)gq * a.Dp * 2, a.Dp, lane_addr);  // the query tile never changes
            sl.reset(q_valid, a, gq, part, 0u);
            if (MT::kBias) {
                __syncwarp();
                side_fetch(t0, 0);
                side_fetch(min(t0 + 1, t1 - 1), 1);
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(aready_bar);
            for (uint32_t t = t0; t < t1; ++t, ++tile_iter) {
                const uint32_t buf = tile_iter & 1u;
                if (!(a.debug_mode & 128)) sl.poll_floor(tile_iter, (uint32_t)a.poll_mask);
                if (MT::kBias) side_fetch(min(t + 2, t1 - 1), (t - t0 + 2) % EPI_SCRATCH_SLOTS);
                if (!mbar_wait(tfull0 + 8u * buf, (tile_iter >> 1) & 1u, abort_flag, 5)) { ok = false; break; }
                tcgen05_fence_after();
                uint32_t v[64];
                if (!(a.debug_mode & 2)) {
                    tmem_ld_32x32b_x64(lane_addr + DCOL + buf * BN, v);
                    tmem_ld_wait();
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty0 + 8u * buf);  // accumulator is in registers: the buffer may be reused
                if (a.debug_mode & 2) continue;
                const uint32_t row0 = t * BN;
                if (MT::kBias) {
                    cp_async_wait<2>();
                    __syncwarp();
                }
                const uint32_t* side = scratch + ((t - t0) % EPI_SCRATCH_SLOTS) * EPI_SCRATCH_WORDS;
                if (a.dump != nullptr) {
                    float* drow = a.dump + (size_t)gq * ((size_t)a.tiles_total * BN) + row0;
                    for (int i = 0; i < 64; ++i) drow[i] = KO::as_f32(sl.fn.key(v[i], MT::kBias ? side[i] : 0u));
                }
                sl.scan64(v, side, row0, a.n_rows, (a.debug_mode & 4) != 0, a.allow_bits);
                if (MT::kBias) __syncwarp();  // every lane is done reading this tile's slot before a later fetch reuses it
            }
            if (MT::kBias) cp_async_wait<0>();
            if (ok) sl.flush(a, gq, part);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    if (threadIdx.x == 0 && *abort_flag) atomicMax(a.error_flag, *abort_flag);
    if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace tc
}  // namespace lb
