"""``lynse`` — the reference package's import name, served by ``lynsedb_b200``.

LynseDB users import ``lynse`` (``lynse.VectorDBClient``, ``lynse._backend.compute_distance`` ...).  This alias package
lets the same imports — and the reference's own test-suite for the search path (tests/golden/ref_tests/) — run against
the B200 library unchanged.  Everything here is a re-export; the implementation lives in ``lynsedb_b200``.
"""
from lynsedb_b200 import *  # noqa: F401,F403
from lynsedb_b200 import __version__, metrics  # noqa: F401
from lynsedb_b200.client import Collection, Database, VectorDBClient  # noqa: F401
from lynsedb_b200.result_view import ResultView  # noqa: F401
