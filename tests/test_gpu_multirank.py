"""GPU, two ranks (NCCL): the SAFE class under torch.distributed -- source rows sharded in define_neighborhoods with
one exchange of packed rows, permutations sharded in compute_pvalues with ONE all-reduce of the device counts -- gives
every rank exactly the single-GPU results.  Skipped on a one-GPU box (run it with `gpurun --gpus 2`)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from safepy_b200 import SAFE, synthetic as syn
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    with np.load(os.path.join(ROOT, "tests", "golden", "stage2_small.npz")) as z:
        g = {k: z[k] for k in z.files}
    n = g["x"].shape[0]
    indptr, indices, csr_len = syn.edges_to_csr(n, g["edges"][:, 0], g["edges"][:, 1], g["length"])
    net = dict(n=n, x=g["x"], y=g["y"], edges=g["edges"], length=g["length"], indptr=indptr, indices=indices,
               csr_length=csr_len)
    sf = SAFE(verbose=False, device=rank)
    sf.load_network(graph=syn.to_networkx(net))
    sf.random_seed = int(g["seed"])
    sf.define_neighborhoods(neighborhood_radius=float(g["radius"]))
    assert np.array_equal(sf.neighborhoods.words, g["neighborhoods"])
    sf.define_neighborhoods(node_distance_metric="euclidean", neighborhood_radius=0.1)
    euclid = sf.neighborhoods.words.copy()
    sf.define_neighborhoods(node_distance_metric="shortpath_weighted_layout", neighborhood_radius=float(g["radius"]))
    sf.load_attributes(attribute_file=g["attr_normal32"].copy())
    P = int(g["num_permutations"])
    sf.compute_pvalues(how="randomization", num_permutations=P, verbose=False)
    assert np.array_equal(sf.pvalues_neg, g["rand_pneg_normal32"], equal_nan=True)
    assert np.array_equal(sf.pvalues_pos, g["rand_ppos_normal32"], equal_nan=True)
    assert np.array_equal(sf.nes, g["rand_nes_normal32"], equal_nan=True)
    assert np.array_equal(sf.nes_binary, g["rand_nesbin_normal32"])
    assert np.array_equal(sf.attributes["num_neighborhoods_enriched"].values, g["rand_enriched_normal32"])
    st = sf.last_enrichment_stats
    share = st["decided"] + st["fixups"]
    assert 0 < share < n * 6 * P                      # this rank counted only its shard
    sf.results_rank = 1                               # only rank 1 receives the [N, M] arrays
    sf.compute_pvalues(how="randomization", num_permutations=P, verbose=False)
    assert np.array_equal(sf.attributes["num_neighborhoods_enriched"].values, g["rand_enriched_normal32"])
    if rank == 1:
        assert np.array_equal(sf.nes, g["rand_nes_normal32"], equal_nan=True)
    else:
        assert sf.nes is None and sf.nes_binary is None and sf.pvalues_pos is None
    sf.results_rank = None
    sf.multi_gpu = False                              # opt out: the rank does everything itself
    sf.define_neighborhoods(node_distance_metric="euclidean", neighborhood_radius=0.1)
    assert np.array_equal(sf.neighborhoods.words, euclid)
    sf.random_seed = None
    sf.multi_gpu = True
    try:
        sf.compute_pvalues(how="randomization", num_permutations=P, verbose=False)
        raise AssertionError("a shared stream needs a seed")
    except ValueError:
        pass
    np.save(os.path.join(out_dir, "ok%d.npy" % rank), np.array([share]))
    dist.barrier()
    dist.destroy_process_group()


def test_safe_class_on_two_gpus(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    shares = [int(np.load(os.path.join(str(tmp_path), "ok%d.npy" % r))[0]) for r in range(2)]
    assert sum(shares) == 400 * 6 * 60
