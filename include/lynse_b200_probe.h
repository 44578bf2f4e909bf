/* lynse_b200_probe.h — C ABI of liblynse_b200_probe.so: hardware rate probes used to measure the roofline
 * denominators committed under profiles/ (tools/mma_probe.py, tools/core_peaks.py).  Diagnostics only: nothing in the
 * product library or the Python package depends on it, and it replaces no reference interface. */
#ifndef LYNSE_B200_PROBE_H
#define LYNSE_B200_PROBE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* message of the last failed probe call on this thread */
const char* lb_probe_last_error(void);

/* tcgen05.mma issue-rate probe (M=128; kind::f16 K=16 bf16, or kind::i8 K=32 when i8 != 0): `iters` MMAs round-robin
 * over n_acc accumulators of n columns, A from TMEM (a_in_tmem=1) or shared memory; returns SM cycles (max over CTAs)
 * to completion and to end of issue. */
int lb_debug_mma_rate(int n, int n_acc, int iters, int a_in_tmem, int i8, int grid, uint64_t* cycles_total, uint64_t* cycles_issue);

/* CUDA-core instruction-rate probe: thread-level instructions per clock per SM of op 0 = popc.b32 + add (one chain step),
 * 1 = lop3.b32, 2 = fp32 fma, 3 = integer add, 4 = three-input integer max; 2 x 1024 threads per SM, 8 independent chains
 * per thread.  The denominators of the popcount / CUDA-core rooflines in profiles/ come from it. */
int lb_debug_core_rate(int op, int iters, double* inst_per_clk_per_sm);

#ifdef __cplusplus
}
#endif
#endif /* LYNSE_B200_PROBE_H */
