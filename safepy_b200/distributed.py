"""Multi-GPU plumbing for the permutation null: one process per GPU, permutations sharded, ONE sum all-reduce.

The randomization null is embarrassingly parallel over permutations once the host has generated the (inherently
sequential) index stream: rank r scores permutations [lo, hi) against the full neighborhood matrix and the counts
add up (safepy/safe.py:518-519 does the same reduction with np.sum over its worker results).  This module holds the
backend-independent part so that it can be exercised with gloo on CPU; bench.py uses it with NCCL on device buffers.
"""
import sys

import numpy as np

from .permutations import make_perm_rows, shard_bounds


def active_group(enabled=True):
    """torch.distributed if the calling program runs one process per GPU -- it has imported torch itself, initialised
    a process group and the world is larger than one -- else None.  (This package never imports torch on its own.)"""
    td = sys.modules.get("torch.distributed")
    if not enabled or td is None or not td.is_available() or not td.is_initialized() or td.get_world_size() < 2:
        return None
    return td


class DeviceArray:
    """Zero-copy view of library-owned device memory for torch.as_tensor (the CUDA array interface, version 2)."""

    def __init__(self, ptr, count, typestr="<i4"):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


def shard_stream(stream, num_permutations, world_size, rank, add):
    """Drive one rank's share of a permutation stream: permutations before the shard are drawn and dropped (the RNG
    cannot jump), `add(stream, count)` consumes the shard, the rest is drawn so that every rank leaves the generator
    where the reference would.  Returns (lo, hi)."""
    lo, hi = shard_bounds(num_permutations, world_size, rank)
    stream.skip(lo)
    if hi > lo:
        add(stream, hi - lo)
    stream.skip(num_permutations - hi)
    return lo, hi


def broadcast_row_shards(dist, rows_tensor, n):
    """All ranks end up with every rank's block of rows of `rows_tensor` ([n, ld], in place): one broadcast per shard
    (shards are unequal when world does not divide n, which all_gather_into_tensor cannot express in place)."""
    world = dist.get_world_size()
    for src in range(world):
        r0, r1 = row_shard(n, world, src)
        if r1 > r0:
            dist.broadcast(rows_tensor[r0:r1], src=src)


def row_shard(n, world_size, rank):
    """Source rows [r0, r1) of stage 1 owned by `rank`: equal blocks of ceil(n / world) rows (the last may be short
    or empty), so that the shards line up with an all-gather buffer of world * ceil(n / world) rows."""
    per = -(-n // world_size)
    r0 = min(n, rank * per)
    return r0, min(n, r0 + per)


def local_perm_rows(node2attribute, num_permutations, random_seed, world_size, rank):
    """Gather rows of this rank's shard.  Every rank replays the whole RNG stream (cumulative shuffles cannot be
    skipped ahead) and keeps its slice; returns (rows[lo:hi], lo, hi)."""
    rows = make_perm_rows(node2attribute, num_permutations, random_seed)
    lo, hi = shard_bounds(num_permutations, world_size, rank)
    return np.ascontiguousarray(rows[lo:hi]), lo, hi


def sharded_perm_counts(count_fn, node2attribute, num_permutations, random_seed, dist=None, device=None):
    """counts_neg, counts_pos over all permutations.

    count_fn(rows) -> (counts_neg, counts_pos) for a block of gather rows (e.g. Enrichment.perm_counts);
    dist: an initialised torch.distributed module (or None for a single process)."""
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    rows, lo, hi = local_perm_rows(node2attribute, num_permutations, random_seed, world, rank)
    n, m = node2attribute.shape
    if hi > lo:
        cneg, cpos = count_fn(rows)
    else:
        cneg = np.zeros((n, m), dtype=np.int64)
        cpos = np.zeros((n, m), dtype=np.int64)
    if dist is None or world == 1:
        return np.asarray(cneg), np.asarray(cpos)
    import torch
    both = torch.from_numpy(np.stack([np.asarray(cneg, dtype=np.int64), np.asarray(cpos, dtype=np.int64)]))
    if device is not None:
        both = both.to(device)
    dist.all_reduce(both)  # the single collective of stage 2
    both = both.cpu().numpy()
    return both[0], both[1]
