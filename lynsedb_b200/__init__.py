"""lynsedb_b200 — B200-native batched vector-distance + top-k path behind LynseDB's Python surface.

Only the search hot path of LynseDB is implemented here; the native side is
hand-written CUDA for sm_100a loaded through ctypes (``lynsedb_b200/liblynse_b200.so``).
"""
from . import metrics
from ._backend import FlatIndex, IvfFlatIndex, compute_distance, top_k_search
from .client import Collection, Database, VectorDBClient
from .index import DeviceIndex, ShardedDeviceIndex, make_allow_bits
from .ivf import IVFIndex
from .result_view import ResultView

import logging as _logging

logger = _logging.getLogger("LynseDB")   # the reference's logger name (python/lynse/utils/utils.py)

__version__ = "0.1.0"
__all__ = ["VectorDBClient", "Database", "Collection", "DeviceIndex", "ShardedDeviceIndex", "IVFIndex", "FlatIndex", "IvfFlatIndex", "ResultView", "compute_distance",
           "top_k_search", "make_allow_bits", "metrics"]
