"""Fixtures for the reference's own search-path tests (test_backend.py, test_search.py, copied by
tools/vendor_ref_tests.py) running against the `lynse` alias of lynsedb_b200.

The fixture names and the data they build are those of the reference's tests/standard_tests/conftest.py (np.random.seed(42),
20 vectors of 8 dims with fields {"tag", "group"}; query from seed 0) — the tests depend on them.  Every test here needs
the GPU.  Tests of subsystems outside the hot path are skipped by name, with the reason, in OUT_OF_SCOPE.
"""
import numpy as np
import pytest

DIM = 8
N = 20

OUT_OF_SCOPE = {
    "test_bm25_search_returns_result_view": "BM25 text search (src/bm25) is outside the distance + top-k path",
    "test_hybrid_search_returns_fused_results": "hybrid BM25 + vector fusion is outside the path",
    "test_search_reranker_reorders_results": "external rerankers run in the reference's Python layer above the path",
    "test_search_reranker_can_read_fields_without_returning_them": "external rerankers",
    "test_bm25_search_reranker_accepts_scores": "BM25 + rerankers",
    "test_hybrid_search_reranker_can_return_id_score_pairs": "hybrid search + rerankers",
    "test_named_vector_field_search": "named vector fields (multi-vector collections) are outside the path",
    "test_named_vector_field_approx_rounds_distances": "named vector fields",
    "test_named_vector_field_search_with_filter_and_fields": "named vector fields",
    "test_sparse_search_returns_inner_product_results": "sparse vectors are a separate index (src/sparse)",
    "test_sparse_search_with_filter_and_fields": "sparse vectors",
    "test_batch_search_reranker_applies_per_query": "external rerankers",
    "test_filtered_search_respects_where_for_quantized_and_graph_indexes": "SQ8 / PQ quantized index modes are outside the path (DESIGN.md §7)",
    "test_query_filter_ids_empty_list": "Collection.query is the SQL metadata query API",
    "test_query_filter_ids_subset_returns_only_those": "Collection.query is the SQL metadata query API",
}


def pytest_collection_modifyitems(config, items):
    for item in items:
        if "ref_tests" not in str(item.fspath):
            continue
        item.add_marker(pytest.mark.gpu)
        base = item.name.split("[")[0]
        if base in OUT_OF_SCOPE:
            item.add_marker(pytest.mark.skip(reason="out of scope: " + OUT_OF_SCOPE[base]))


@pytest.fixture(scope="function")
def tmp_root(tmp_path):
    yield str(tmp_path)


@pytest.fixture(scope="function")
def client(tmp_root):
    import lynse

    c = lynse.VectorDBClient(uri=tmp_root)
    yield c
    c.close()


@pytest.fixture(scope="function")
def db(client):
    yield client.create_database("test_db", drop_if_exists=True)


@pytest.fixture(scope="function")
def collection(db):
    yield db.require_collection("test_col", dim=DIM, drop_if_exists=True)


@pytest.fixture(scope="function")
def populated_collection(collection):
    np.random.seed(42)
    vectors = [np.random.rand(DIM).astype(np.float32) for _ in range(N)]
    fields = [{"tag": f"item_{i}", "group": i % 3} for i in range(N)]
    with collection.insert_session() as session:
        session.add(ids=list(range(N)), vectors=vectors, fields=fields)
    yield collection


@pytest.fixture(scope="function")
def query_vec():
    np.random.seed(0)
    return np.random.rand(DIM).astype(np.float32)
