"""safepy_b200 -- SAFE's neighborhood + enrichment path on one B200 per process (hand-written sm_100a CUDA behind a
C ABI, see include/safe_b200.h).  Importing never touches the GPU; the first call that computes does, and raises
SafeB200Error if the library or device is missing (there is no CPU fallback)."""
from ._lib import SafeB200Error, load_library  # noqa: F401
from .neighborhood_matrix import PackedNeighborhoods  # noqa: F401
from .safe import SAFE, SafeB200Mixin, accelerate, get_context  # noqa: F401
from .safe_extras import compute_neighborhood_score, run_permutations  # noqa: F401

__all__ = ["SAFE", "SafeB200Mixin", "accelerate", "get_context", "compute_neighborhood_score", "run_permutations",
           "PackedNeighborhoods", "SafeB200Error", "load_library"]
