// lb_probe.cu — hardware rate probes (tcgen05.mma issue rate, CUDA-core instruction rates).  Built into its own library,
// liblynse_b200_probe.so: the roofline denominators under profiles/ come from here, the product library does not
// carry this code.  Declared in include/lynse_b200_probe.h; run by tools/mma_probe.py and tools/core_peaks.py.
#include <algorithm>
#include <vector>

#include "lb_tc.cuh"
#include "../../include/lynse_b200_probe.h"

namespace lb {
static thread_local std::string g_probe_error;
void set_error(const std::string& msg) { g_probe_error = msg; }
int fail(int status, const std::string& msg) {
    g_probe_error = msg;
    return status;
}
static int probe_env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}
namespace tc {
// ---- diagnostics: tcgen05.mma issue-rate probe ------------------------------------------------------------------
// One warp issues `iters` MMAs (M=128, K=16, bf16) round-robin over `n_acc` independent accumulators of N columns;
// operands are whatever is in shared memory / TMEM (timing only); I8 = kind::i8 (K = 32) instead of kind::f16 (K = 16).  Reports SM cycles from first issue to last commit.
__device__ __forceinline__ void umma_ss_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void umma_ss_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

template <int N, int NACC, bool TS, bool I8>
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
        constexpr uint32_t idesc = make_idesc<I8>(BM, N);
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
                    if (TS && I8)
                        umma_ts_i8(d, tmem_base + (uint32_t)(j * 8), b_desc + (uint64_t)((j & 3) * 2), idesc, acc);
                    else if (TS)
                        umma_ts_bf16(d, tmem_base + (uint32_t)(j * 8), b_desc + (uint64_t)((j & 3) * 2), idesc, acc);
                    else if (I8)
                        umma_ss_i8(d, a_desc + (uint64_t)((j & 3) * 2), b_desc + (uint64_t)((j & 3) * 2), idesc, acc);
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

using namespace lb;

// ---- CUDA-core instruction-rate probe (roofline denominators of the non-tensor kernels) ---------------------------------------------
// Every thread runs `iters` rounds of eight independent dependency chains of one instruction; 2 CTAs x 1024 threads per SM
// keep every scheduler full.  Result: instructions per clock per SM, from the SM's own cycle counter.
template <int OP>
static __global__ void __launch_bounds__(1024, 2) core_rate_kernel(int iters, uint32_t seed, unsigned long long* cycles, uint32_t* sink) {
    uint32_t x[8];
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        x[i] = seed * (threadIdx.x + 1) + i * 0x9E3779B9u;
        f[i] = (float)(x[i] & 0xffff) * 1e-5f;
    }
    const uint32_t a = seed | 1u, b = seed * 3u + 7u;
    const float fa = 1.0000001f, fb = 1e-9f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) x[i] = __popc(x[i]) + a;                          // POPC + IADD: the add keeps the chain data-dependent
            else if (OP == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a), "r"(b));
            else if (OP == 2) f[i] = __fmaf_rn(f[i], fa, fb);
            else if (OP == 3) x[i] = x[i] + a;                               // IADD alone (to subtract from OP 0)
            else x[i] = max(max(x[i], a), b + (uint32_t)it);                  // VIMNMX3
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc ^= x[i] ^ __float_as_uint(f[i]);
    if (acc == 0x12345678u) sink[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

extern "C" {

const char* lb_probe_last_error(void) { return g_probe_error.c_str(); }

int lb_debug_mma_rate(int n, int n_acc, int iters, int a_in_tmem, int i8, int grid, uint64_t* cycles_total, uint64_t* cycles_issue) {
    if (iters < 16 || grid < 1) return fail(LB_INVALID_ARGUMENT, "bad probe arguments");
    unsigned long long* d = nullptr;
    LB_CUDA_TRY(cudaMalloc(&d, (size_t)grid * 16));
    const size_t smem = 49152 + 64 + 1024;
    cudaError_t e = cudaSuccess;
    bool found = false;
#define LB_PROBE(NN, NA, TSV, I8V)                                                                                         \
    if (!found && n == NN && n_acc == NA && (a_in_tmem != 0) == TSV && (i8 != 0) == I8V) {                                 \
        found = true;                                                                                                      \
        e = cudaFuncSetAttribute(tc::mma_rate_kernel<NN, NA, TSV, I8V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e == cudaSuccess) {                                                                                            \
            tc::mma_rate_kernel<NN, NA, TSV, I8V><<<grid, 64, smem>>>(iters / 16, probe_env_int("LYNSE_B200_PROBE_COMMIT", 0), d); \
            e = cudaDeviceSynchronize();                                                                                   \
        }                                                                                                                  \
    }
    LB_PROBE(64, 1, true, false)
    LB_PROBE(64, 2, true, false)
    LB_PROBE(128, 1, true, false)
    LB_PROBE(128, 2, true, false)
    LB_PROBE(64, 2, false, false)
    LB_PROBE(128, 2, false, false)
    LB_PROBE(256, 2, false, false)
    LB_PROBE(64, 2, true, true)
    LB_PROBE(128, 1, true, true)
    LB_PROBE(128, 2, true, true)
    LB_PROBE(128, 2, false, true)
    LB_PROBE(256, 2, false, true)
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


int lb_debug_core_rate(int op, int iters, double* inst_per_clk_per_sm) {
    if (op < 0 || op > 4 || iters < 1 || !inst_per_clk_per_sm) return fail(LB_INVALID_ARGUMENT, "bad probe arguments");
    int dev = 0, sms = 0;
    LB_CUDA_TRY(cudaGetDevice(&dev));
    LB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = sms * 2;
    unsigned long long* d_cycles = nullptr;
    uint32_t* d_sink = nullptr;
    LB_CUDA_TRY(cudaMalloc(&d_cycles, (size_t)grid * 8));
    LB_CUDA_TRY(cudaMalloc(&d_sink, 4));
    for (int rep = 0; rep < 2; ++rep) {
        switch (op) {
            case 0: core_rate_kernel<0><<<grid, 1024>>>(iters, 12345u, d_cycles, d_sink); break;
            case 1: core_rate_kernel<1><<<grid, 1024>>>(iters, 12345u, d_cycles, d_sink); break;
            case 2: core_rate_kernel<2><<<grid, 1024>>>(iters, 12345u, d_cycles, d_sink); break;
            case 3: core_rate_kernel<3><<<grid, 1024>>>(iters, 12345u, d_cycles, d_sink); break;
            default: core_rate_kernel<4><<<grid, 1024>>>(iters, 12345u, d_cycles, d_sink); break;
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<unsigned long long> h(grid);
    if (e == cudaSuccess) e = cudaMemcpy(h.data(), d_cycles, (size_t)grid * 8, cudaMemcpyDeviceToHost);
    cudaFree(d_cycles);
    cudaFree(d_sink);
    if (e != cudaSuccess) return fail(LB_CUDA, std::string("core rate probe: ") + cudaGetErrorString(e));
    std::sort(h.begin(), h.end());
    const double cyc = (double)h[grid / 2];
    // two resident CTAs of 1024 threads per SM, 8 instructions per thread and round
    *inst_per_clk_per_sm = 2.0 * 1024.0 * 8.0 * (double)iters / cyc;
    return LB_OK;
}

}  // extern "C"
