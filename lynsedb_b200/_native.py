"""ctypes binding of liblynse_b200.so (include/lynse_b200.h).

This is the whole native boundary of the package: plain pointers and sizes in,
status codes out.  There is no CPU fallback — if the shared library is missing
the import fails loudly, and every compute call fails with ``RuntimeError``
when no B200 is visible.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("LYNSE_B200_LIB", _HERE / "liblynse_b200.so"))

LB_OK, LB_INVALID_ARGUMENT, LB_DIMENSION_MISMATCH, LB_IO, LB_CUDA, LB_NCCL, LB_UNSUPPORTED, LB_INTERNAL = range(8)
LB_F32, LB_PACKED_U64, LB_F16 = 0, 1, 2
LB_PLAN_AUTO, LB_PLAN_EXACT = 0, 1
ROW_NONE = 0xFFFFFFFF


class SearchStats(C.Structure):
    _fields_ = [
        ("plan_used", C.c_uint32),
        ("n_fallback", C.c_uint32),
        ("n_partitions", C.c_uint32),
        ("kernels_launched", C.c_uint32),
        ("ms_dominant", C.c_float),
        ("ms_total", C.c_float),
        ("algorithmic_bytes", C.c_uint64),
        ("algorithmic_flops", C.c_uint64),
        ("coarse_operand", C.c_uint32),
        ("coarse_hit_mode", C.c_uint32),
        ("coarse_sm_mhz", C.c_float),
        ("reserved", C.c_uint32),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


_f32p, _u32p, _u64p, _u8p = C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint8)
_vp, _vpp = C.c_void_p, C.POINTER(C.c_void_p)

# name -> (restype, argtypes); kept in one table so tests can check it against the header
SIGNATURES = {
    "lb_last_error": (C.c_char_p, []),
    "lb_version": (C.c_char_p, []),
    "lb_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "lb_device_info": (C.c_int, [C.c_int, C.c_char_p, C.c_int, _u64p, _u64p, C.POINTER(C.c_int)]),
    "lb_compute_distance": (C.c_int, [_f32p, _f32p, C.c_uint32, C.c_int, _f32p]),
    "lb_top_k_search": (C.c_int, [_f32p, _f32p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, _u32p, _f32p, _u32p]),
    "lb_index_create": (C.c_int, [_vpp, C.c_uint32, C.c_int, C.c_int]),
    "lb_index_destroy": (None, [_vp]),
    "lb_index_reserve": (C.c_int, [_vp, C.c_uint64]),
    "lb_index_set_segment_target": (C.c_int, [_vp, C.c_uint64]),
    "lb_index_new_segment": (C.c_int, [_vp]),
    "lb_index_append_f32": (C.c_int, [_vp, _f32p, C.c_uint64]),
    "lb_index_append_packed": (C.c_int, [_vp, _u64p, C.c_uint64]),
    "lb_index_append_f16": (C.c_int, [_vp, C.POINTER(C.c_uint16), C.c_uint64]),
    "lb_index_append_synthetic": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_uint64]),
    "lb_index_len": (C.c_uint64, [_vp]),
    "lb_index_dim": (C.c_uint32, [_vp]),
    "lb_index_segments": (C.c_int, [_vp, _u64p, C.c_int, C.POINTER(C.c_int)]),
    "lb_index_read_rows_f32": (C.c_int, [_vp, C.c_uint64, C.c_uint64, _f32p]),
    "lb_index_prepare": (C.c_int, [_vp, C.c_int]),
    "lb_index_set_plan": (C.c_int, [_vp, C.c_int]),
    "lb_index_search": (C.c_int, [_vp, C.c_int, _f32p, C.c_uint32, C.c_uint32, _u64p, C.c_uint64, _u32p, _f32p, _u32p]),
    "lb_index_search_pairwise": (C.c_int, [_vp, C.c_int, _f32p, C.c_uint32, C.c_uint32, _u64p, C.c_uint64, _u32p, _f32p, _u32p]),
    "lb_index_search_f16_rows": (C.c_int, [_vp, C.c_int, _f32p, C.c_uint32, C.c_uint32, _u64p, C.c_uint64, _u32p, _f32p, _u32p]),
    "lb_index_search_packed": (C.c_int, [_vp, C.c_int, _u64p, C.c_uint32, C.c_uint32, _u32p, _f32p, _u32p]),
    "lb_index_search_device": (C.c_int, [_vp, C.c_int, _vp, C.c_uint32, C.c_uint32, _vp, _vp, _vp]),
    "lb_index_set_timing": (C.c_int, [_vp, C.c_int]),
    "lb_index_last_stats": (C.c_int, [_vp, C.POINTER(SearchStats)]),
    "lb_ivf_train": (C.c_int, [_vp, C.c_int, C.c_uint32, C.c_uint32, _vpp]),
    "lb_ivf_create": (C.c_int, [_vp, C.c_int, _f32p, C.c_uint32, _u32p, _vpp]),
    "lb_ivf_destroy": (None, [_vp]),
    "lb_ivf_info": (C.c_int, [_vp, _u32p, _u64p]),
    "lb_ivf_centroids": (C.c_int, [_vp, _f32p]),
    "lb_ivf_assignments": (C.c_int, [_vp, _u32p]),
    "lb_ivf_search": (C.c_int, [_vp, _f32p, C.c_uint32, C.c_uint32, C.c_uint32, _u64p, C.c_uint64, _u32p, _f32p, _u32p]),
    "lb_ivf_flat_search": (C.c_int, [_vp, _f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, _u32p, _f32p, _u32p]),
    "lb_device_malloc": (C.c_int, [C.c_int, C.c_uint64, _vpp]),
    "lb_device_free": (C.c_int, [C.c_int, _vp]),
    "lb_host_malloc": (C.c_int, [C.c_uint64, _vpp]),
    "lb_host_free": (C.c_int, [_vp]),
    "lb_memcpy_h2d": (C.c_int, [C.c_int, _vp, _vp, C.c_uint64]),
    "lb_memcpy_d2h": (C.c_int, [C.c_int, _vp, _vp, C.c_uint64]),
    "lb_device_synchronize": (C.c_int, [C.c_int]),
    "lb_device_memset": (C.c_int, [C.c_int, _vp, C.c_int, C.c_uint64]),
    "lb_nccl_unique_id": (C.c_int, [_u8p]),
    "lb_comm_create": (C.c_int, [_vpp, C.c_int, C.c_int, C.c_int, _u8p]),
    "lb_comm_destroy": (None, [_vp]),
    "lb_comm_allgather": (C.c_int, [_vp, _vp, _vp, C.c_uint64]),
    "lb_comm_barrier": (C.c_int, [_vp]),
    "lb_sharded_search": (C.c_int, [_vp, _vp, C.c_int, _f32p, C.c_uint32, C.c_uint32, C.c_uint64, _u64p, _f32p, _u32p]),
    "lb_sharded_search_filtered": (C.c_int, [_vp, _vp, C.c_int, _f32p, C.c_uint32, C.c_uint32, C.c_uint64, _u64p, C.c_uint64, _u64p, _f32p, _u32p]),
    "lb_merge_shard_blocks": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, _u32p, _f32p, _u32p, _u64p, _u64p, _f32p, _u32p]),
    "lb_sharded_search_packed": (C.c_int, [_vp, _vp, C.c_int, _u64p, C.c_uint32, C.c_uint32, C.c_uint64, _u64p, _f32p, _u32p]),
    "lb_sharded_search_device": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_uint32, C.c_uint32, C.c_uint64, _vp, _vp, _vp]),
    "lb_index_event_record": (C.c_int, [_vp, C.c_int]),
    "lb_index_event_elapsed_ms": (C.c_int, [_vp, C.c_int, C.c_int, _f32p]),
    "lb_debug_tc_scores": (C.c_int, [_f32p, C.c_uint32, _f32p, C.c_uint32, C.c_uint32, C.c_int, _f32p]),
}

_lib = None


def lib():
    """Load liblynse_b200.so (once).  Missing library -> ImportError, never a silent fallback."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C lynsedb_b200/csrc` (nvcc, sm_100a). lynsedb_b200 has no CPU fallback."
            )
        L = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


PROBE_LIB_PATH = LIB_PATH.with_name("liblynse_b200_probe.so")
PROBE_SIGNATURES = {
    "lb_probe_last_error": (C.c_char_p, []),
    "lb_debug_mma_rate": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _u64p, _u64p]),
    "lb_debug_core_rate": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_double)]),
}
_probe_lib = None


def probe_lib():
    """The diagnostic rate probes (include/lynse_b200_probe.h), a library of their own: tools/ only."""
    global _probe_lib
    if _probe_lib is None:
        if not PROBE_LIB_PATH.exists():
            raise ImportError(f"{PROBE_LIB_PATH} is missing; build it with `make -C lynsedb_b200/csrc`")
        L = C.CDLL(str(PROBE_LIB_PATH))
        for name, (res, args) in PROBE_SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _probe_lib = L
    return _probe_lib


def last_error() -> str:
    msg = lib().lb_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(status: int) -> None:
    """Map lb_status to the exception the reference raises (src/error.rs:54-72)."""
    if status == LB_OK:
        return
    msg = last_error()
    if status in (LB_INVALID_ARGUMENT, LB_DIMENSION_MISMATCH):
        raise ValueError(msg)
    if status == LB_IO:
        raise IOError(msg)
    raise RuntimeError(msg)


def fptr(a: np.ndarray):
    return a.ctypes.data_as(_f32p)


def u32ptr(a: np.ndarray):
    return a.ctypes.data_as(_u32p)


def u64ptr(a: np.ndarray):
    return a.ctypes.data_as(_u64p)


def device_count() -> int:
    n = C.c_int(0)
    status = lib().lb_device_count(C.byref(n))
    return n.value if status == LB_OK else 0


def device_info(device: int = 0) -> dict:
    name = C.create_string_buffer(256)
    total, free, sms = C.c_uint64(0), C.c_uint64(0), C.c_int(0)
    check(lib().lb_device_info(device, name, 256, C.byref(total), C.byref(free), C.byref(sms)))
    return {"name": name.value.decode(), "total_bytes": total.value, "free_bytes": free.value, "sm_count": sms.value}
