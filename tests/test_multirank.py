"""World-size-2 run of the permutation sharding + count all-reduce on CPU (gloo).  The per-shard counts come from
the oracle here (no GPU in this test); the GPU path plugs Enrichment.perm_counts into the same function."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, golden, perms, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch.distributed as dist
    import safe_oracle as orc
    from safepy_b200.distributed import sharded_perm_counts
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    with np.load(golden) as z:
        dense = z["neighborhoods_dense"] if "neighborhoods_dense" in z.files else None
        attrs = z["attr_normal32"]
        packed = z["neighborhoods"]
    if dense is None:
        from safepy_b200._lib import unpack_packed
        dense = unpack_packed(packed, attrs.shape[0]).astype(np.int64)

    def count_fn(rows):
        return orc.perm_counts_from_rows(dense, attrs, "sum", rows)

    cneg, cpos = sharded_perm_counts(count_fn, attrs, perms, 7, dist=dist)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), cneg=cneg, cpos=cpos)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("perms", [11, 2])
def test_two_ranks_reproduce_the_single_process_counts(tmp_path, perms):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import safe_oracle as orc
    from safepy_b200._lib import unpack_packed
    golden = os.path.join(ROOT, "tests", "golden", "stage2_small.npz")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, golden, perms, str(tmp_path)), nprocs=2, join=True)
    with np.load(golden) as z:
        attrs = z["attr_normal32"]
        dense = unpack_packed(z["neighborhoods"], attrs.shape[0]).astype(np.int64)
    ref_neg, ref_pos = orc.run_permutations(dense, attrs, "sum", perms, 7)
    for r in range(2):
        with np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) as z:
            assert np.array_equal(z["cneg"], ref_neg.astype(np.int64))
            assert np.array_equal(z["cpos"], ref_pos.astype(np.int64))


def test_shards_partition_the_stream():
    from safepy_b200.distributed import local_perm_rows
    from safepy_b200.permutations import make_perm_rows
    rng = np.random.default_rng(3)
    b = rng.standard_normal((50, 3)).astype(np.float32)
    b[::7] = np.nan
    full = make_perm_rows(b, 13, 5)
    got = [local_perm_rows(b, 13, 5, 4, r) for r in range(4)]
    assert np.array_equal(np.concatenate([g[0] for g in got]), full)
    assert [(g[1], g[2]) for g in got] == [(0, 4), (4, 8), (8, 12), (12, 13)]


def test_row_shards_cover_all_sources():
    from safepy_b200.distributed import row_shard
    for n, w in ((20000, 8), (3971, 4), (10, 4), (7, 8)):
        per = -(-n // w)
        got = [row_shard(n, w, r) for r in range(w)]
        assert got[0][0] == 0 and max(b for _, b in got) == n
        assert all(a == min(n, r * per) for r, (a, _) in enumerate(got))
        covered = sorted(i for a, b in got for i in range(a, b))
        assert covered == list(range(n))


def _stream_worker(rank, world, port, golden, perms, out_dir):
    """The SAFE-class form of the sharding: a native permutation stream per rank (earlier permutations drawn and
    dropped), this rank's share counted, ONE all-reduce -- here with the oracle standing in for the device."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch.distributed as dist
    import safe_oracle as orc
    from safepy_b200._lib import unpack_packed
    from safepy_b200.distributed import active_group, broadcast_row_shards, row_shard, shard_stream
    from safepy_b200.permutations import perm_stream
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    assert active_group() is None                      # no process group yet
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert active_group() is dist and active_group(False) is None
    with np.load(golden) as z:
        attrs = z["attr_normal32"]
        packed = z["neighborhoods"]
    n = attrs.shape[0]
    # stage 1 exchange: every rank owns a block of packed rows, one broadcast per block
    r0, r1 = row_shard(n, world, rank)
    mine = torch.zeros(packed.shape, dtype=torch.int32)
    mine[r0:r1] = torch.from_numpy(packed.view(np.int32)[r0:r1])
    broadcast_row_shards(dist, mine, n)
    assert np.array_equal(mine.numpy().view(np.uint32), packed)
    dense = unpack_packed(packed, n).astype(np.int64)
    counts = np.zeros((2, n, attrs.shape[1]), dtype=np.int64)

    def add(stream, count):
        cneg, cpos = orc.perm_counts_from_rows(dense, attrs, "sum", stream.next(count))
        counts[0] += cneg.astype(np.int64)
        counts[1] += cpos.astype(np.int64)

    stream = perm_stream(attrs, 7)
    lo, hi = shard_stream(stream, perms, world, rank, add)
    assert stream.state()[2] == perms                  # every rank drew the whole stream
    both = torch.from_numpy(counts)
    dist.all_reduce(both)
    np.savez(os.path.join(out_dir, "srank%d.npz" % rank), cneg=both[0].numpy(), cpos=both[1].numpy(), lo=lo, hi=hi)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("perms", [11, 1])
def test_two_ranks_stream_sharding(tmp_path, perms):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import safe_oracle as orc
    from safepy_b200._lib import unpack_packed
    golden = os.path.join(ROOT, "tests", "golden", "stage2_small.npz")
    mp.spawn(_stream_worker, args=(2, _free_port(), golden, perms, str(tmp_path)), nprocs=2, join=True)
    with np.load(golden) as z:
        attrs = z["attr_normal32"]
        dense = unpack_packed(z["neighborhoods"], attrs.shape[0]).astype(np.int64)
    ref_neg, ref_pos = orc.run_permutations(dense, attrs, "sum", perms, 7)
    spans = []
    for r in range(2):
        with np.load(os.path.join(str(tmp_path), "srank%d.npz" % r)) as z:
            assert np.array_equal(z["cneg"], ref_neg.astype(np.int64))
            assert np.array_equal(z["cpos"], ref_pos.astype(np.int64))
            spans.append((int(z["lo"]), int(z["hi"])))
    assert spans[0][0] == 0 and spans[0][1] == spans[1][0] and spans[1][1] == perms
