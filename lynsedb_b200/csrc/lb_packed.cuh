// lb_packed.cuh — packed one-bit rows of 16 u64 words (1024-bit fingerprints): Hamming / Jaccard / Tanimoto / Dice.
//
// Replaces the reference's packed_binary_search (src/storage/flat_mmap.rs:1345-1409) over
// packed_hamming_u64 / packed_jaccard_u64 / packed_dice_u64 (src/distance/simd.rs:765-801).  Integer
// arithmetic: counts are exact, the final f32 division is IEEE, so ids, order and distances are bit-identical.
//
// Shape (persistent CTAs, 256 threads, one row per thread):
//   * the corpus streams through a ring of 32 KiB stages, one TMA box of 256 rows x 128 B per stage with
//     SWIZZLE_128B, so the 16-byte chunk c of row r lands at chunk c ^ (r & 7) and the per-thread 128-bit
//     reads of "my row" are bank-conflict free; HBM sees whole 32 KiB sequential reads;
//   * a thread keeps its row (16 u64) in registers and walks the queries in tiles of 16 held in shared memory
//     (broadcast reads); 1024-bit popcounts are either 32 POPC or a Harley-Seal carry-save tree (6 POPC);
//   * a (partition, query) pair owns a k-entry list in global memory gated by its current worst key; a thread
//     whose key passes the gate appends it to a per-tile candidate buffer, and one warp per query folds the
//     (rare) candidates into the list — keys order by (distance, row), so the result is the reference's top-k
//     however rows are partitioned.
#pragma once
#include "lb_tc.cuh"

namespace lb {

constexpr int PK_W = 16;                    // u64 words per row
constexpr int PK_ROWS = 256;                // rows per stage == threads per CTA
constexpr int PK_STAGE_BYTES = PK_ROWS * PK_W * 8;  // 32 KiB
constexpr int PK_NSTAGES = 2;
constexpr int PK_TQ = 16;                   // queries per tile
constexpr uint32_t PK_SMEM_CAND = PK_NSTAGES * PK_STAGE_BYTES;                 // candidates [TQ][256] u64
constexpr uint32_t PK_SMEM_Q = PK_SMEM_CAND + PK_TQ * PK_ROWS * 8;             // query words [TQ][16] u64
constexpr uint32_t PK_SMEM_THR = PK_SMEM_Q + PK_TQ * PK_W * 8;                 // gates [TQ] u64
constexpr uint32_t PK_SMEM_CNT = PK_SMEM_THR + PK_TQ * 8;                      // candidate counts [TQ] u32, query popcounts [TQ] u32
constexpr uint32_t PK_SMEM_BAR = PK_SMEM_CNT + PK_TQ * 8;                      // full[NSTAGES]
constexpr uint32_t PK_SMEM_LISTS = PK_SMEM_BAR + 64;                           // optional [TQ][k] keys + [TQ] counts
constexpr uint32_t PK_SMEM_BYTES = PK_SMEM_LISTS + 1024;                       // + the lists when ScanArgs::smem_lists

// carry-save adder on 32-bit lanes: (a + b + c) = sum + 2 * carry, bitwise
#define LB_CSA(h, l, a, b, c)                  \
    do {                                       \
        const uint32_t _u = (a) ^ (b);         \
        (h) = ((a) & (b)) | (_u & (c));        \
        (l) = _u ^ (c);                        \
    } while (0)

// popcount of 32 x 32-bit words x[0..31] with a Harley-Seal tree: 31 CSAs + 6 POPC instead of 32 POPC
__device__ __forceinline__ uint32_t popc1024_hs(const uint32_t* x) {
    uint32_t ones = 0, twos = 0, fours = 0, eights = 0, sixteens = 0, thirtytwos;
    uint32_t twosA, twosB, foursA, foursB, eightsA, eightsB, sixteensA, sixteensB;
    LB_CSA(twosA, ones, ones, x[0], x[1]);
    LB_CSA(twosB, ones, ones, x[2], x[3]);
    LB_CSA(foursA, twos, twos, twosA, twosB);
    LB_CSA(twosA, ones, ones, x[4], x[5]);
    LB_CSA(twosB, ones, ones, x[6], x[7]);
    LB_CSA(foursB, twos, twos, twosA, twosB);
    LB_CSA(eightsA, fours, fours, foursA, foursB);
    LB_CSA(twosA, ones, ones, x[8], x[9]);
    LB_CSA(twosB, ones, ones, x[10], x[11]);
    LB_CSA(foursA, twos, twos, twosA, twosB);
    LB_CSA(twosA, ones, ones, x[12], x[13]);
    LB_CSA(twosB, ones, ones, x[14], x[15]);
    LB_CSA(foursB, twos, twos, twosA, twosB);
    LB_CSA(eightsB, fours, fours, foursA, foursB);
    LB_CSA(sixteensA, eights, eights, eightsA, eightsB);
    LB_CSA(twosA, ones, ones, x[16], x[17]);
    LB_CSA(twosB, ones, ones, x[18], x[19]);
    LB_CSA(foursA, twos, twos, twosA, twosB);
    LB_CSA(twosA, ones, ones, x[20], x[21]);
    LB_CSA(twosB, ones, ones, x[22], x[23]);
    LB_CSA(foursB, twos, twos, twosA, twosB);
    LB_CSA(eightsA, fours, fours, foursA, foursB);
    LB_CSA(twosA, ones, ones, x[24], x[25]);
    LB_CSA(twosB, ones, ones, x[26], x[27]);
    LB_CSA(foursA, twos, twos, twosA, twosB);
    LB_CSA(twosA, ones, ones, x[28], x[29]);
    LB_CSA(twosB, ones, ones, x[30], x[31]);
    LB_CSA(foursB, twos, twos, twosA, twosB);
    LB_CSA(eightsB, fours, fours, foursA, foursB);
    LB_CSA(sixteensB, eights, eights, eightsA, eightsB);
    LB_CSA(thirtytwos, sixteens, sixteens, sixteensA, sixteensB);
    return 32u * __popc(thirtytwos) + 16u * __popc(sixteens) + 8u * __popc(eights) + 4u * __popc(fours) + 2u * __popc(twos) +
           __popc(ones);
}

// one warp folds n candidate keys (shared memory) of one query into its (partition, query) list; the list lives in
// global memory (GLOBAL_LIST: read through L2 with __ldcg) or, for batches of a single query tile, in shared memory
// for the whole kernel (a list update is then tens of cycles instead of two L2 round trips).
// returns the new gate (KEY_NONE while the list is not full)
template <bool GLOBAL_LIST>
__device__ __forceinline__ uint64_t warp_fold_candidates_t(const uint64_t* __restrict__ cand, int n, uint64_t* list, uint32_t* count_p,
                                                           uint64_t* thr_p, int k, int lane) {
    auto ld = [&](const uint64_t* p) -> uint64_t { return GLOBAL_LIST ? __ldcg(p) : *reinterpret_cast<const volatile uint64_t*>(p); };
    const uint32_t cnt0 = *reinterpret_cast<volatile uint32_t*>(count_p);
    uint32_t cnt = cnt0;
    uint64_t thr = 0xFFFFFFFFFFFFFFFFull;  // KEY_NONE
    if (cnt0 >= (uint32_t)k) thr = *reinterpret_cast<volatile uint64_t*>(thr_p);
    __syncwarp();  // every lane has read the list header before lane 0 may rewrite it below
    if (cnt0 == 0 && n >= k && n <= 256) {
        // first block of a partition: the list is empty and every row is a candidate.  Take the k smallest by k
        // warp-wide minimum reductions instead of ~n serial insertions (the dominant cost of a short partition).
        uint64_t c[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) c[j] = (j * 32 + lane < n) ? cand[j * 32 + lane] : 0xFFFFFFFFFFFFFFFFull;
        uint64_t w = 0;
        for (int r = 0; r < k; ++r) {
            uint64_t m = c[0];
#pragma unroll
            for (int j = 1; j < 8; ++j) m = c[j] < m ? c[j] : m;
            w = warp_min_u64(m);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (c[j] == w) c[j] = 0xFFFFFFFFFFFFFFFFull;  // keys are unique: one holder
            if (lane == 0) list[r] = w;
        }
        if (lane == 0) {
            *count_p = (uint32_t)k;
            *thr_p = w;
        }
        __syncwarp();
        return w;
    }
    for (int base = 0; base < n; base += 32) {
        uint64_t key = 0xFFFFFFFFFFFFFFFFull;
        if (base + lane < n) key = cand[base + lane];
        unsigned m = __ballot_sync(0xffffffffu, key < thr);
        while (m) {
            int src = __ffs(m) - 1;
            m &= m - 1;
            uint64_t kk = __shfl_sync(0xffffffffu, key, src);
            if (kk >= thr) continue;  // gate tightened by an earlier insert of this group
            if (cnt < (uint32_t)k) {
                if (lane == 0) list[cnt] = kk;
                ++cnt;
                __syncwarp();
                if (cnt == (uint32_t)k) {
                    uint64_t mx = 0;
                    for (int idx = lane; idx < k; idx += 32) {
                        uint64_t v = ld(list + idx);
                        mx = v > mx ? v : mx;
                    }
                    thr = warp_max_u64(mx);
                }
            } else {
                for (int idx = lane; idx < k; idx += 32)
                    if (ld(list + idx) == thr) list[idx] = kk;  // keys are unique: exactly one slot holds the worst
                __syncwarp();
                uint64_t mx = 0;
                for (int idx = lane; idx < k; idx += 32) {
                    uint64_t v = ld(list + idx);
                    mx = v > mx ? v : mx;
                }
                thr = warp_max_u64(mx);
            }
        }
    }
    if (lane == 0) {
        *count_p = cnt;
        *thr_p = thr;
    }
    __syncwarp();
    return thr;
}
__device__ __forceinline__ uint64_t warp_fold_candidates(const uint64_t* __restrict__ cand, int n, uint64_t* list, uint32_t* count_p,
                                                         uint64_t* thr_p, int k, int lane) {
    return warp_fold_candidates_t<true>(cand, n, list, count_p, thr_p, k, lane);
}

// Lists of a single-tile batch kept in shared memory: [tq][k] keys + [tq] counts, written back once at the end.
struct SmemLists {
    uint64_t* keys;    // [TQ][k]
    uint32_t* counts;  // [TQ]
    __device__ __forceinline__ void init(int tq, int tid, int nthreads) {
        for (int i = tid; i < tq; i += nthreads) counts[i] = 0u;
    }
    __device__ __forceinline__ void write_back(const ScanArgs& a, int part, int tq, const uint64_t* sthr, int warp, int lane, int nwarps) {
        for (int j = warp; j < tq; j += nwarps) {
            const size_t lq = (size_t)part * a.nq + j;
            const uint32_t cnt = counts[j];
            for (int i = lane; i < (int)cnt; i += 32) a.lists[lq * a.k + i] = keys[(size_t)j * a.k + i];
            if (lane == 0) {
                a.counts[lq] = cnt;
                a.thr[lq] = sthr[j];
            }
        }
    }
};

// MODE: 0 = Hamming (popc(x ^ y)), 1 = Jaccard / Tanimoto, 2 = Dice (both from popc(x & y) and the two row popcounts)
template <int MODE, bool HS>
__global__ void __launch_bounds__(PK_ROWS, 2) scan_packed16_kernel(const __grid_constant__ CUtensorMap tmap, ScanArgs a) {
    extern __shared__ __align__(16) unsigned char smem_pk[];
    const uint32_t smem_base = (tc::smem_u32(smem_pk) + 1023u) & ~1023u;
    unsigned char* smem = smem_pk + (smem_base - tc::smem_u32(smem_pk));
    uint64_t* cand = reinterpret_cast<uint64_t*>(smem + PK_SMEM_CAND);   // [TQ][256]
    uint64_t* sq = reinterpret_cast<uint64_t*>(smem + PK_SMEM_Q);        // [TQ][16]
    uint64_t* sthr = reinterpret_cast<uint64_t*>(smem + PK_SMEM_THR);    // [TQ]
    uint32_t* scnt = reinterpret_cast<uint32_t*>(smem + PK_SMEM_CNT);    // [TQ]
    uint32_t* sqpop = scnt + PK_TQ;                                       // [TQ]
    const uint32_t full0 = smem_base + PK_SMEM_BAR;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int part = blockIdx.x;
    SmemLists sl;
    sl.keys = reinterpret_cast<uint64_t*>(smem + PK_SMEM_LISTS);          // [nq <= TQ][k] when a.smem_lists
    sl.counts = reinterpret_cast<uint32_t*>(sl.keys + (size_t)PK_TQ * a.k);
    const uint64_t part_begin = (uint64_t)part * a.rows_per_part;
    uint64_t part_end = part_begin + a.rows_per_part;
    if (part_end > a.n_rows) part_end = a.n_rows;
    const uint32_t n_blocks = part_end > part_begin ? (uint32_t)((part_end - part_begin + PK_ROWS - 1) / PK_ROWS) : 0u;

    if (tid == 0) {
        for (int s = 0; s < PK_NSTAGES; ++s) tc::mbar_init(full0 + 8u * s, 1);
        tc::fence_barrier_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    }
    __syncthreads();
    if (tid == 0) {
        for (uint32_t b = 0; b < (uint32_t)PK_NSTAGES && b < n_blocks; ++b) {
            tc::mbar_arrive_expect_tx(full0 + 8u * b, PK_STAGE_BYTES);
            tc::tma_load_2d(smem_base + b * PK_STAGE_BYTES, &tmap, 0, (int)(part_begin + (uint64_t)b * PK_ROWS), full0 + 8u * b);
        }
    }

    const bool single_tile = a.nq <= PK_TQ;
    if (single_tile) {
        if (tid < a.nq * PK_W) sq[tid] = __ldg(a.qwords + tid);
        if (tid < a.nq) {
            sthr[tid] = KEY_NONE;  // the lists of this launch start empty (counts are zeroed by the host)
            scnt[tid] = 0u;
        }
        if (a.smem_lists) sl.init(a.nq, tid, PK_ROWS);
        __syncthreads();
        if (tid < a.nq) {
            uint32_t p = 0;
#pragma unroll
            for (int w = 0; w < PK_W; ++w) p += __popcll(sq[tid * PK_W + w]);
            sqpop[tid] = p;
        }
        __syncthreads();
    }

    for (uint32_t blk = 0; blk < n_blocks; ++blk) {
        const uint32_t stage = blk % PK_NSTAGES, phase = (blk / PK_NSTAGES) & 1u;
        while (!tc::mbar_try_wait(full0 + 8u * stage, phase)) {
        }
        // my row: 8 x 128-bit reads, chunk c at c ^ (row & 7) (rows past the end of the corpus are zero-filled by TMA)
        uint32_t r[32];
        {
            const uint4* rowp = reinterpret_cast<const uint4*>(smem + stage * PK_STAGE_BYTES + tid * 128);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint4 x = rowp[c ^ (tid & 7)];
                r[4 * c + 0] = x.x; r[4 * c + 1] = x.y; r[4 * c + 2] = x.z; r[4 * c + 3] = x.w;
            }
        }
        const uint64_t row64 = part_begin + (uint64_t)blk * PK_ROWS + tid;
        const bool valid = row64 < part_end && row_allowed(a.allow_bits, (uint32_t)row64);
        const uint32_t row = (uint32_t)row64;
        uint32_t row_pop = 0;
        if (MODE != 0) {
#pragma unroll
            for (int w = 0; w < 32; ++w) row_pop += __popc(r[w]);
        }
        __syncthreads();  // every thread holds its row: the stage can be refilled
        if (tid == 0 && blk + PK_NSTAGES < n_blocks) {
            tc::mbar_arrive_expect_tx(full0 + 8u * stage, PK_STAGE_BYTES);
            tc::tma_load_2d(smem_base + stage * PK_STAGE_BYTES, &tmap, 0, (int)(part_begin + (uint64_t)(blk + PK_NSTAGES) * PK_ROWS),
                            full0 + 8u * stage);
        }
        for (int q0 = 0; q0 < a.nq; q0 += PK_TQ) {
            const int tq = min(PK_TQ, a.nq - q0);
            // query tile, gates and counters.  A batch that fits one tile keeps them in shared memory for the whole
            // kernel (loaded before the block loop, gates updated by the fold): nothing global on the per-block path.
            if (!single_tile) {
                if (tid < tq * PK_W) sq[tid] = __ldg(a.qwords + (size_t)q0 * PK_W + tid);
                if (tid < tq) {
                    const size_t lq = (size_t)part * a.nq + (q0 + tid);
                    sthr[tid] = __ldcg(a.counts + lq) < (uint32_t)a.k ? KEY_NONE : __ldcg(a.thr + lq);
                    scnt[tid] = 0u;
                }
                __syncthreads();
                if (MODE != 0 && tid < tq) {
                    uint32_t p = 0;
#pragma unroll
                    for (int w = 0; w < PK_W; ++w) p += __popcll(sq[tid * PK_W + w]);
                    sqpop[tid] = p;
                }
                if (MODE != 0) __syncthreads();
            }
            for (int j = 0; j < tq; ++j) {
                const uint4* qp = reinterpret_cast<const uint4*>(sq + j * PK_W);
                uint32_t x[32];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint4 qv = qp[c];  // broadcast
                    if (MODE == 0) {
                        x[4 * c + 0] = r[4 * c + 0] ^ qv.x; x[4 * c + 1] = r[4 * c + 1] ^ qv.y;
                        x[4 * c + 2] = r[4 * c + 2] ^ qv.z; x[4 * c + 3] = r[4 * c + 3] ^ qv.w;
                    } else {
                        x[4 * c + 0] = r[4 * c + 0] & qv.x; x[4 * c + 1] = r[4 * c + 1] & qv.y;
                        x[4 * c + 2] = r[4 * c + 2] & qv.z; x[4 * c + 3] = r[4 * c + 3] & qv.w;
                    }
                }
                uint32_t cnt;
                if (HS) {
                    cnt = popc1024_hs(x);
                } else {
                    cnt = 0;
#pragma unroll
                    for (int w = 0; w < 32; ++w) cnt += __popc(x[w]);
                }
                float d;
                if (MODE == 0) {
                    d = (float)cnt;
                } else if (MODE == 1) {
                    const uint32_t uni = row_pop + sqpop[j] - cnt;
                    d = uni == 0 ? 0.0f : 1.0f - (float)cnt / (float)uni;
                } else {
                    const uint32_t tot = row_pop + sqpop[j];
                    d = tot == 0 ? 0.0f : 1.0f - (float)(2u * cnt) / (float)tot;
                }
                const uint64_t key = make_key<true>(d, row);
                if (valid && key < sthr[j]) {
                    const uint32_t pos = atomicAdd(&scnt[j], 1u);
                    cand[j * PK_ROWS + pos] = key;
                }
            }
            __syncthreads();
            for (int j = warp; j < tq; j += PK_ROWS / 32) {
                const int n = (int)scnt[j];
                if (n > 0) {
                    const size_t lq = (size_t)part * a.nq + (q0 + j);
                    const uint64_t g = (single_tile && a.smem_lists)
                                           ? warp_fold_candidates_t<false>(cand + j * PK_ROWS, n, sl.keys + (size_t)j * a.k, sl.counts + j, sthr + j, a.k, lane)
                                           : warp_fold_candidates(cand + j * PK_ROWS, n, a.lists + lq * a.k, a.counts + lq, a.thr + lq, a.k, lane);
                    if (lane == 0) {
                        sthr[j] = g;
                        scnt[j] = 0u;
                    }
                }
            }
            __syncthreads();
        }
    }
    if (single_tile && a.smem_lists) sl.write_back(a, part, a.nq, sthr, warp, lane, PK_ROWS / 32);
}

}  // namespace lb
