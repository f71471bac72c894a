"""Host-side permutation stream for the randomization null.

The reference draws one `np.random.permutation(indx_vals)` per iteration from the legacy global MT19937 stream
after `np.random.seed(random_seed)` and applies it IN PLACE to the already permuted attribute matrix
(safepy/safe_extras.py:46-58).  The shuffle is inherently sequential, so it stays on the host; what goes to the GPU
is the composition of those shuffles as plain gather indices: rows[p, t] is the row of the ORIGINAL matrix that node
t holds during permutation p.  Rows without any data are never moved (safe_extras.py:51).
"""
import numpy as np


def rows_with_data(node2attribute):
    """indx_vals of safe_extras.py:51."""
    return np.nonzero(np.sum(~np.isnan(node2attribute), axis=1))[0]


def make_perm_rows(node2attribute, num_permutations, random_seed, out=None):
    """Replay the reference's RNG calls and compose them. Returns int32 [num_permutations, n].

    Consumes the global NumPy RNG exactly like the reference does, so interleaving with reference code keeps both
    streams aligned.  random_seed=None seeds from OS entropy (np.random.seed(None)), as upstream."""
    n = node2attribute.shape[0]
    np.random.seed(random_seed)
    indx_vals = rows_with_data(node2attribute)
    rows = out if out is not None else np.empty((num_permutations, n), dtype=np.int32)
    cur = np.arange(n, dtype=np.int32)
    for p in range(num_permutations):
        # n2a[indx_vals, :] = n2a[np.random.permutation(indx_vals), :]
        cur[indx_vals] = cur[np.random.permutation(indx_vals)]
        rows[p] = cur
    return rows


def shard_bounds(num_permutations, world_size, rank):
    """Contiguous permutation range [lo, hi) owned by `rank` (the last ranks get the short shards)."""
    per = -(-num_permutations // world_size)
    lo = min(num_permutations, rank * per)
    return lo, min(num_permutations, lo + per)
