"""``lynse.result_view`` (python/lynse/result_view.py of the reference)."""
from lynsedb_b200.result_view import *  # noqa: F401,F403
from lynsedb_b200.result_view import ResultView, _parse_index_mode  # noqa: F401
