"""Host reproduction of the device-side synthetic corpus generator.

``lb_index_append_synthetic`` fills rows on the GPU with a counter-based hash
(lynsedb_b200/csrc/lb_common.cuh: ``synth_u64`` / ``synth_f32``) so that 10M+
row corpora never exist in host memory.  These numpy functions return the very
same values for any subset of rows, which is what lets the CPU oracle check
GPU results at full benchmark sizes (shape of the data follows the reference's
benchmarks/flat_search_bench.py:49-79: U[0,1) floats, Bernoulli(1/2) bits).
"""
from __future__ import annotations

import numpy as np

_M1 = np.uint64(0xFF51AFD7ED558CCD)
_M2 = np.uint64(0xC4CEB9FE1A85EC53)
_GOLD = np.uint64(0x9E3779B97F4A7C15)
_SEED_SALT = np.uint64(0x632BE59BD9B4E019)
_S33 = np.uint64(33)


def _mix64(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        x ^= x >> _S33
        x *= _M1
        x ^= x >> _S33
        x *= _M2
        x ^= x >> _S33
    return x


def synth_u64(seed: int, index: np.ndarray) -> np.ndarray:
    index = np.asarray(index, dtype=np.uint64)
    with np.errstate(over="ignore"):
        salt = _mix64(np.asarray([np.uint64(seed) + _SEED_SALT], dtype=np.uint64))[0]
        return _mix64(index * _GOLD + salt)


def synth_f32(seed: int, index: np.ndarray) -> np.ndarray:
    return ((synth_u64(seed, index) >> np.uint64(40)).astype(np.float32)) * np.float32(1.0 / 16777216.0)


def rows_f32(seed: int, rows: np.ndarray, dim: int) -> np.ndarray:
    """f32 rows ``rows`` (global row numbers) of the synthetic corpus with dimension ``dim``."""
    rows = np.asarray(rows, dtype=np.uint64).reshape(-1, 1)
    with np.errstate(over="ignore"):
        idx = rows * np.uint64(dim) + np.arange(dim, dtype=np.uint64).reshape(1, -1)
    return synth_f32(seed, idx)


def rows_packed(seed: int, rows: np.ndarray, n_words: int) -> np.ndarray:
    rows = np.asarray(rows, dtype=np.uint64).reshape(-1, 1)
    with np.errstate(over="ignore"):
        idx = rows * np.uint64(n_words) + np.arange(n_words, dtype=np.uint64).reshape(1, -1)
    return synth_u64(seed, idx)
