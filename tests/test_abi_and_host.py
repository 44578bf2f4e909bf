"""CPU-side checks: the C ABI library loads and exports every symbol the header declares, the host-side
mirror of the reference's metric/alias/index-mode tables is right, and the product fails loudly without a GPU."""
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def native():
    import __graft_entry__ as g

    if not (ROOT / "lynsedb_b200" / "liblynse_b200.so").exists():
        g.build()
    from lynsedb_b200 import _native

    return _native


def _header_functions():
    text = (ROOT / "include" / "lynse_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(native):
    declared = _header_functions()
    assert len(declared) >= 35
    out = subprocess.run(["nm", "-D", "--defined-only", str(native.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (lb_[a-z0-9_]+)", out))
    missing = [f for f in declared if f not in exported]
    assert not missing, f"declared in include/lynse_b200.h but not exported: {missing}"
    unbound = [f for f in declared if f not in native.SIGNATURES]
    assert not unbound, f"declared but not bound in _native.SIGNATURES: {unbound}"
    native.lib()  # resolves every bound symbol


def test_library_is_sm100a_with_tcgen05_and_tma(native):
    sass = subprocess.run(["cuobjdump", "-sass", str(native.LIB_PATH)], capture_output=True, text=True).stdout
    if not sass:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "STTM"):
        assert mnemonic in sass, f"{mnemonic} missing: the tensor-core path is not in the build"


def test_no_gpu_means_loud_failure_not_a_cpu_fallback(native):
    if native.device_count() > 0:
        pytest.skip("a CUDA device is present")
    import lynsedb_b200 as L

    with pytest.raises(RuntimeError):
        L.compute_distance(np.ones(4, np.float32), np.ones(4, np.float32), "ip")
    with pytest.raises(RuntimeError):
        L.DeviceIndex(8)
    with pytest.raises(RuntimeError):
        L.top_k_search(np.ones(4, np.float32), np.ones((3, 4), np.float32), "l2", 2)


def test_argument_errors_come_before_any_device_work(native):
    import lynsedb_b200 as L

    with pytest.raises(ValueError, match="Unknown metric"):
        L.compute_distance(np.ones(4, np.float32), np.ones(4, np.float32), "nope")
    with pytest.raises(ValueError, match="dimensions must match"):
        L.compute_distance(np.ones(4, np.float32), np.ones(3, np.float32), "ip")
    with pytest.raises(ValueError, match="haversine requires two values"):
        L.compute_distance(np.ones(4, np.float32), np.ones(4, np.float32), "haversine")
    with pytest.raises(ValueError, match="Query dimension must match"):
        L.top_k_search(np.ones(4, np.float32), np.ones((3, 5), np.float32), "ip", 2)


def test_product_package_never_imports_the_oracle():
    for path in (ROOT / "lynsedb_b200").rglob("*.py"):
        text = path.read_text()
        assert not re.search(r"^\s*(import oracle|from oracle)", text, flags=re.M), f"{path} imports the oracle"
    for path in (ROOT / "lynsedb_b200" / "csrc").glob("*"):
        if path.suffix in (".cu", ".cuh", ".cpp", ".h"):
            assert "oracle" not in path.read_text().lower().replace("the cpu oracle", ""), f"{path} references the oracle"


# ---- metric tables (reference src/distance/mod.rs:645-704, src/index/mod.rs:838-873) --------------------------
def test_metric_aliases():
    from lynsedb_b200 import metrics as M

    table = {"ip": M.IP, "dot": M.IP, "inner_product": M.IP, "INNER": M.IP, "l2": M.L2, "euclidean": M.L2, "l2sq": M.L2,
             "cos": M.COSINE, "cosine_distance": M.COSINE, "hamming": M.HAMMING, "jaccard": M.JACCARD, "cityblock": M.MANHATTAN,
             "l1": M.MANHATTAN, "geo": M.HAVERSINE, "haversine-m": M.HAVERSINE, "pearson": M.CORRELATION,
             "hellinger": M.HELLINGER, "emd": M.WASSERSTEIN, "wasserstein-1d": M.WASSERSTEIN, "sorensen-dice": M.DICE,
             "tanimoto": M.TANIMOTO, "js": M.JENSEN_SHANNON, "jensen-shannon": M.JENSEN_SHANNON, "linf": M.CHEBYSHEV,
             "chebychev": M.CHEBYSHEV, "canberra": M.CANBERRA, "bray-curtis": M.BRAY_CURTIS, "braycurtis": M.BRAY_CURTIS}
    for name, want in table.items():
        assert M.from_str(name) == want, name
    assert M.from_str("nope") is None
    assert [M.NAMES[M.from_str(n)] for n in M.NAMES] == M.NAMES
    assert not M.is_ascending(M.IP) and all(M.is_ascending(m) for m in range(1, 16))
    assert [m for m in range(16) if M.is_binary(m)] == [M.HAMMING, M.JACCARD, M.DICE, M.TANIMOTO]
    assert M.accepts_dimension(M.HAVERSINE, 2) and not M.accepts_dimension(M.HAVERSINE, 3) and not M.accepts_dimension(M.IP, 0)


def test_index_mode_parsing():
    from lynsedb_b200 import metrics as M

    for metric, mode in enumerate(M.FLAT_INDEX_MODE):
        assert M.from_index_mode(mode) == metric, mode
    assert M.from_index_mode("HNSW-CORRELATION") == M.CORRELATION
    assert M.from_index_mode("IVF-JENSEN-SHANNON") == M.JENSEN_SHANNON
    assert M.from_index_mode("flat-tanimoto-binary") == M.TANIMOTO
    assert M.from_index_mode("IVF-IP-SQ8") == M.IP
    assert M.from_index_mode("FLAT") is None
    assert M.parse_index_mode("FLAT-IP") == ("Flat", "IP")
    assert M.parse_index_mode("FLAT-COS-SQ8") == ("Flat", "Cosine")
    assert M.parse_index_mode("IVF-HAMMING-BINARY") == ("IVF", "Hamming")
    assert M.parse_index_mode("SPANN-L2") == ("SPANN", "L2")
    assert M.parse_index_mode("FLAT-JENSEN-SHANNON") == ("Flat", "Jensen-Shannon")
    assert M.parse_index_mode("FLAT-WASSERSTEIN") == ("Flat", "Wasserstein-1D")
    assert M.parse_index_mode(None) == ("Flat", "IP")


def test_synthetic_generator_is_stable():
    from lynsedb_b200 import synthetic

    a = synthetic.rows_f32(42, [0, 1, 10**7], 4)
    assert a.dtype == np.float32 and a.shape == (3, 4)
    assert np.all((a >= 0) & (a < 1))
    np.testing.assert_array_equal(a[0], np.array([0.20437068, 0.9639019, 0.64448464, 0.45936424], np.float32))
    assert np.array_equal(synthetic.rows_f32(42, [10**7], 4)[0], a[2])
    w = synthetic.rows_packed(1, [5], 2)
    assert w.dtype == np.uint64 and int(w[0, 0]) == 11770942570365255535


def test_result_view_contract():
    from lynsedb_b200 import ResultView

    ids, d = np.array([3, 1], np.int64), np.array([0.5, 0.7], np.float32)
    rv = ResultView(ids=ids, distances=d, k=2, distance="IP", index="Flat")
    a, b, f = rv
    assert a is ids and b is d and f == []
    assert len(rv) == 2 and bool(rv) and rv["ids"] is ids and rv["distance"] is d and rv["measure"] == "IP" and rv["n"] == 2
    assert rv.to_tuple()[0] is ids and rv.to_dict()["ids"] == [3, 1]
    assert rv == ResultView(ids=ids.copy(), distances=d.copy())
    with pytest.raises(KeyError):
        rv["nope"]
    with pytest.raises(TypeError):
        rv[0]
    assert not ResultView(ids=np.array([], np.int64), distances=np.array([], np.float32))
