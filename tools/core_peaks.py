"""Measures the CUDA-core instruction rates the non-tensor rooflines use and the tcgen05 MMA cadence of both operand
kinds; writes profiles/r2_cuda_core_peaks.json (run on a B200)."""
import ctypes as C
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from lynsedb_b200 import _native as N  # noqa: E402

lib = N.probe_lib()   # liblynse_b200_probe.so (include/lynse_b200_probe.h)
out = {}
names = {0: "popc_plus_add", 1: "lop3_b32", 2: "fp32_fma", 3: "iadd", 4: "vimnmx3"}
for op, name in names.items():
    v = C.c_double(0)
    if lib.lb_debug_core_rate(op, 4096, C.byref(v)) != 0:
        raise RuntimeError(lib.lb_probe_last_error().decode())
    out[name + "_per_clk_per_sm"] = v.value
# popc alone: a chain step is popc + add; rate(popc) = 1 / (1/rate(step) - 1/rate(add)) when the two pipes do not overlap,
# and rate(step) when they do: report both bounds, use the conservative (overlapped) one as the ceiling
step, add = out["popc_plus_add_per_clk_per_sm"], out["iadd_per_clk_per_sm"]
out["popc_b32_per_clk_per_sm"] = step
out["popc_b32_per_clk_per_sm_if_serialised_with_add"] = 1.0 / max(1.0 / step - 1.0 / add, 1e-9)
info = N.device_info(0)
out["sm_count"] = info["sm_count"]
out["device"] = info["name"]
for i8 in (0, 1):
    t, i = C.c_uint64(0), C.c_uint64(0)
    if lib.lb_debug_mma_rate(128, 2, 4096, 1, i8, info["sm_count"], C.byref(t), C.byref(i)) != 0:
        raise RuntimeError(lib.lb_probe_last_error().decode())
    out["tcgen05_" + ("i8" if i8 else "f16") + "_m128_n128_cycles_per_mma"] = t.value / 4096
out["note"] = ("thread-level instructions per clock per SM (2 x 1024 threads per SM, 8 chains per thread); tcgen05: cycles per "
               "M=128 x N=128 MMA (kind::f16 K=16, kind::i8 K=32), A in TMEM, all SMs busy")
path = ROOT / "gpurun_out" / "r2_cuda_core_peaks.json"
path.parent.mkdir(exist_ok=True)
path.write_text(json.dumps(out, indent=1))
print(json.dumps(out, indent=1))
