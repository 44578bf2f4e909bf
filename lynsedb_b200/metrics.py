"""Metric names, aliases and index-mode parsing.

Restates, as data, the reference's ``DistanceMetric`` surface
(reference src/distance/mod.rs:19-189) and the display names of
python/lynse/result_view.py:14-72.  Ids are the enum order of the reference and
of ``lb_metric`` in include/lynse_b200.h.
"""
from __future__ import annotations

from typing import Optional, Tuple

IP, L2, COSINE, HAMMING, JACCARD, MANHATTAN, HAVERSINE, CORRELATION, HELLINGER, WASSERSTEIN, DICE, TANIMOTO, \
    JENSEN_SHANNON, CHEBYSHEV, CANBERRA, BRAY_CURTIS = range(16)

# DistanceMetric::name (distance/mod.rs:118-137)
NAMES = ["ip", "l2", "cosine", "hamming", "jaccard", "l1", "haversine", "correlation", "hellinger", "wasserstein",
         "dice", "tanimoto", "jensen_shannon", "chebyshev", "canberra", "bray_curtis"]

# DistanceMetric::from_str (distance/mod.rs:39-67)
_ALIASES = {
    IP: ("ip", "inner_product", "inner", "dot"),
    L2: ("l2", "l2sq", "l2_squared", "euclidean"),
    COSINE: ("cosine", "cos", "cosine_distance"),
    HAMMING: ("hamming",),
    JACCARD: ("jaccard",),
    MANHATTAN: ("l1", "manhattan", "cityblock"),
    HAVERSINE: ("haversine", "haversine_m", "haversine-m", "geo"),
    CORRELATION: ("correlation", "pearson"),
    HELLINGER: ("hellinger",),
    WASSERSTEIN: ("wasserstein", "wasserstein1d", "wasserstein_1d", "wasserstein-1d", "emd"),
    DICE: ("dice", "sorensen", "sorensen_dice", "sorensen-dice"),
    TANIMOTO: ("tanimoto",),
    JENSEN_SHANNON: ("jensen_shannon", "jensen-shannon", "jensenshannon", "js"),
    CHEBYSHEV: ("chebyshev", "chebychev", "linf", "l_inf", "l-infinity"),
    CANBERRA: ("canberra",),
    BRAY_CURTIS: ("bray_curtis", "bray-curtis", "braycurtis"),
}
_FROM_STR = {alias: metric for metric, aliases in _ALIASES.items() for alias in aliases}

# DistanceMetric::flat_index_mode (distance/mod.rs:139-158)
FLAT_INDEX_MODE = ["FLAT-IP", "FLAT-L2", "FLAT-COS", "FLAT-HAMMING-BINARY", "FLAT-JACCARD-BINARY", "FLAT-L1",
                   "FLAT-HAVERSINE", "FLAT-CORRELATION", "FLAT-HELLINGER", "FLAT-WASSERSTEIN", "FLAT-DICE-BINARY",
                   "FLAT-TANIMOTO-BINARY", "FLAT-JENSEN-SHANNON", "FLAT-CHEBYSHEV", "FLAT-CANBERRA",
                   "FLAT-BRAY-CURTIS"]


def from_str(name) -> Optional[int]:
    """``DistanceMetric::from_str``: case-insensitive alias lookup, ``None`` when unknown."""
    if isinstance(name, int) and not isinstance(name, bool):
        return name if 0 <= name < 16 else None
    return _FROM_STR.get(str(name).lower())


def require(name) -> int:
    metric = from_str(name)
    if metric is None:
        raise ValueError(f"Unknown metric: {name}")
    return metric


def from_index_mode(mode: str) -> Optional[int]:
    """``DistanceMetric::from_index_mode`` with the reference's precedence (distance/mod.rs:69-108)."""
    tokens = mode.upper().split("-")
    has = tokens.__contains__
    if has("JENSENSHANNON") or has("JS") or (has("JENSEN") and has("SHANNON")):
        return JENSEN_SHANNON
    if has("CHEBYSHEV") or has("CHEBYCHEV") or has("LINF"):
        return CHEBYSHEV
    if has("CANBERRA"):
        return CANBERRA
    if has("BRAYCURTIS") or (has("BRAY") and has("CURTIS")):
        return BRAY_CURTIS
    if has("TANIMOTO"):
        return TANIMOTO
    if has("JACCARD"):
        return JACCARD
    if has("HAMMING"):
        return HAMMING
    if has("DICE") or has("SORENSEN"):
        return DICE
    if has("HAVERSINE") or has("GEO"):
        return HAVERSINE
    if has("CORRELATION") or has("PEARSON"):
        return CORRELATION
    if has("HELLINGER"):
        return HELLINGER
    if has("WASSERSTEIN") or has("WASSERSTEIN1D") or has("EMD"):
        return WASSERSTEIN
    if has("L1") or has("MANHATTAN") or has("CITYBLOCK"):
        return MANHATTAN
    if has("L2") or has("L2SQ"):
        return L2
    if has("COS") or has("COSINE"):
        return COSINE
    if has("IP"):
        return IP
    return None


def is_ascending(metric: int) -> bool:
    """Lower is better for everything except inner product (distance/mod.rs:111-116)."""
    return metric != IP


def is_binary(metric: int) -> bool:
    """Metrics evaluated on packed one-bit rows in the flat path (distance/mod.rs:161-166)."""
    return metric in (HAMMING, JACCARD, DICE, TANIMOTO)


def accepts_dimension(metric: int, dimension: int) -> bool:
    """distance/mod.rs:169-174"""
    return dimension == 2 if metric == HAVERSINE else dimension > 0


def parse_index_mode(index_mode: Optional[str]) -> Tuple[str, str]:
    """(index type, metric display name) of an index-mode string — the contract of
    python/lynse/result_view.py:14-72 (``_parse_index_mode``)."""
    if not index_mode:
        return ("Flat", "IP")
    parts = index_mode.upper().split("-")
    idx_type = {"FLAT": "Flat", "IVF": "IVF", "SPANN": "SPANN", "HNSW": "HNSW", "DISKANN": "DiskANN"}.get(parts[0], parts[0])
    full = "-".join(parts[1:])
    table = [
        (("TANIMOTO",), "Tanimoto"), (("JACCARD",), "Jaccard"), (("HAMMING",), "Hamming"),
        (("DICE", "SORENSEN"), "Dice"), (("HAVERSINE", "GEO"), "Haversine"),
        (("CORRELATION", "PEARSON"), "Correlation"), (("HELLINGER",), "Hellinger"),
        (("WASSERSTEIN", "EMD"), "Wasserstein-1D"),
    ]
    for needles, label in table:
        if any(n in full for n in needles):
            return idx_type, label
    if "JENSEN" in full or full == "JS":
        return idx_type, "Jensen-Shannon"
    table2 = [
        (("CHEBYSHEV", "CHEBYCHEV", "LINF"), "Chebyshev"), (("CANBERRA",), "Canberra"), (("BRAY",), "Bray-Curtis"),
        (("L1", "MANHATTAN", "CITYBLOCK"), "L1"), (("L2",), "L2"), (("COS",), "Cosine"),
    ]
    for needles, label in table2:
        if any(n in full for n in needles):
            return idx_type, label
    return idx_type, "IP"
