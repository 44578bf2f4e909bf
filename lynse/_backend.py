"""``lynse._backend`` (python/lynse/_backend.py of the reference): the stateless operators and the raw index classes."""
from lynsedb_b200._backend import *  # noqa: F401,F403
from lynsedb_b200._backend import FlatIndex, IvfFlatIndex, compute_distance, top_k_search  # noqa: F401
