"""Drop-in replacements for the two free functions of safepy/safe_extras.py, running on the B200.

Same names, argument order and return conventions as the reference:
    compute_neighborhood_score(neighborhood2node, node2attribute, neighborhood_score_type)   safe_extras.py:6
    run_permutations(arg_tuple, **kwargs)                                                      safe_extras.py:36
"""
import numpy as np

from . import _lib
from .neighborhood_matrix import as_packed
from .permutations import make_perm_rows
from .safe import get_context


def compute_neighborhood_score(neighborhood2node, node2attribute, neighborhood_score_type):
    """[N, M] float64 neighborhood scores ('sum' or 'z-score'); NaN attribute values count as missing."""
    if neighborhood_score_type not in _lib.SCORE_TYPES:
        neighborhood_score_type = "sum"  # the reference computes the plain sum for any other string
    ctx = get_context()
    plan = _lib.Enrichment(as_packed(neighborhood2node).on_device(ctx), node2attribute)
    try:
        return plan.score(neighborhood_score_type)
    finally:
        plan.close()


def run_permutations(arg_tuple, **kwargs):
    """(counts_neg, counts_pos) as float64 arrays holding integers, exactly like the reference; the 5-tuple is
    (neighborhood2node, node2attribute, neighborhood_score_type, num_permutations, random_seed).
    Keyword `verbose` is accepted for compatibility (there is no per-iteration progress to show), `engine`
    selects 'auto' | 'tc' | 'simt'."""
    neighborhood2node, node2attribute, neighborhood_score_type, num_permutations, random_seed = arg_tuple
    if neighborhood_score_type not in _lib.SCORE_TYPES:
        neighborhood_score_type = "sum"
    node2attribute = np.asarray(node2attribute)
    rows = make_perm_rows(node2attribute, int(num_permutations), random_seed)
    ctx = get_context()
    plan = _lib.Enrichment(as_packed(neighborhood2node).on_device(ctx), node2attribute)
    try:
        cneg, cpos = plan.perm_counts(rows, neighborhood_score_type, kwargs.get("engine", "auto"))
    finally:
        plan.close()
    return cneg.astype(np.float64), cpos.astype(np.float64)
