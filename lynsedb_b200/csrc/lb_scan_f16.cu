// lb_scan_f16.cu — the dense exact scans over binary16 rows (float16 indexes): lb_scan_dense.cuh with RT = __half.
#include "lb_scan_dense.cuh"

namespace lb {

template int dense_scan_launch<__half>(lb_index*, const ScanRequest&, ScanArgs&, const ScanPlan&, bool, bool, bool);

}  // namespace lb
