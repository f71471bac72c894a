"""GPU: stage-1 kernels (through the C ABI) against the reference goldens and the oracle.  Bit-exact."""
import numpy as np
import pytest

import safe_oracle as orc
from conftest import net_from_golden
from safepy_b200 import _lib, synthetic as syn
from safepy_b200._lib import unpack_packed

pytestmark = pytest.mark.gpu


def run_shortpath(ctx, net, cost, cutoff, row0=0, row1=None):
    nb = _lib.Neighborhoods(ctx, net["n"])
    nb.shortpath(net["indptr"], net["indices"], cost, cutoff, row0, row1)
    return nb


@pytest.mark.parametrize("which", ["stage1_small", "stage1_mid"])
def test_golden_neighborhoods_bit_exact(ctx, which, request):
    g = request.getfixturevalue(which)
    net = net_from_golden(g)
    n = net["n"]
    extent = np.max(net["x"]) - np.min(net["x"])
    nb = run_shortpath(ctx, net, net["csr_length"], float(g["r_layout"]) * extent)
    assert np.array_equal(nb.packed(), g["nb_layout"])
    assert np.array_equal(nb.rowsums(), unpack_packed(g["nb_layout"], n).sum(axis=1))
    assert np.array_equal(nb.dense(dtype=np.int64), unpack_packed(g["nb_layout"], n))
    assert np.array_equal(run_shortpath(ctx, net, None, float(g["r_hops"])).packed(), g["nb_hops"])
    _, _, wts = syn.edges_to_csr(n, g["edges"][:, 0], g["edges"][:, 1], g["edge_weight"])
    assert np.array_equal(run_shortpath(ctx, net, wts, 3.0).packed(), g["nb_weighted_hops"])
    eu = _lib.Neighborhoods(ctx, n).euclid(net["x"], net["y"], float(g["r_euclid"]) * extent)
    assert np.array_equal(eu.packed(), g["nb_euclid"])


def test_row_ranges_compose(ctx, stage1_mid):
    """Source sharding: computing disjoint row ranges into one matrix equals the single-shot result."""
    g = stage1_mid
    net = net_from_golden(g)
    n = net["n"]
    extent = np.max(net["x"]) - np.min(net["x"])
    nb = _lib.Neighborhoods(ctx, n)
    for r0, r1 in ((0, 1), (1, 700), (700, 700), (700, n)):
        nb.shortpath(net["indptr"], net["indices"], net["csr_length"], float(g["r_layout"]) * extent, r0, r1)
    assert np.array_equal(nb.packed(), g["nb_layout"])
    eu = _lib.Neighborhoods(ctx, n)
    for r0, r1 in ((0, 129), (129, n)):
        eu.euclid(net["x"], net["y"], float(g["r_euclid"]) * extent, r0, r1)
    assert np.array_equal(eu.packed(), g["nb_euclid"])


def test_edge_cases(ctx):
    # isolated nodes, zero-length edge, duplicate coordinates, tiny n, radius 0 and huge radius
    x = np.array([0.0, 0.0, 1.0, 2.0, 5.0])
    y = np.array([0.0, 0.0, 0.0, 0.0, 5.0])
    edges = np.array([[0, 1], [1, 2], [2, 3]])
    length = np.array([0.0, 1.0, 1.0])
    indptr, indices, w = syn.edges_to_csr(5, edges[:, 0], edges[:, 1], length)
    import networkx as nx
    g = nx.Graph()
    g.add_nodes_from(range(5))
    g.add_weighted_edges_from([(0, 1, 0.0), (1, 2, 1.0), (2, 3, 1.0)], weight="length")
    for cutoff in (0.0, 0.5, 1.0, 2.0, 1e9):
        nb = _lib.Neighborhoods(ctx, 5).shortpath(indptr, indices, w, cutoff)
        assert np.array_equal(nb.dense(), orc.neighborhoods_shortpath_nx(g, cutoff, "length")), cutoff
    for nr in (0.0, 1e-300, 1.0, 1.0000000000000002, 1e9):
        eu = _lib.Neighborhoods(ctx, 5).euclid(x, y, nr)
        assert np.array_equal(eu.dense(), orc.neighborhoods_euclidean(x, y, nr)), nr
    one = _lib.Neighborhoods(ctx, 1).euclid(np.zeros(1), np.zeros(1), 1.0)
    assert one.dense().tolist() == [[1]]
    with pytest.raises(_lib.SafeB200Error):
        _lib.Neighborhoods(ctx, 5).shortpath(indptr, indices, -w - 1, 1.0)
    with pytest.raises(_lib.SafeB200Error):
        _lib.Neighborhoods(ctx, 5).shortpath(indptr, indices + 9, w, 1.0)


def test_euclid_strict_threshold_ties(ctx):
    """Points on an integer lattice: many distances equal the radius exactly; `<` must exclude them."""
    gx, gy = np.meshgrid(np.arange(40.0), np.arange(25.0))
    x, y = gx.ravel(), gy.ravel()
    for nr in (5.0, np.sqrt(2.0), 13.0):
        eu = _lib.Neighborhoods(ctx, x.size).euclid(x, y, nr)
        assert np.array_equal(eu.dense(), orc.neighborhoods_euclidean(x, y, nr))


def test_c1_shape_network_sampled_rows(ctx):
    """Example-1 shape (3971 nodes / 28k edges): every row against csgraph Dijkstra, bit-exact."""
    c = syn.make_config("C1")
    net = c["net"]
    nr = c["radius"] * (np.max(net["x"]) - np.min(net["x"]))
    nb = run_shortpath(ctx, net, net["csr_length"], nr)
    ref = orc.neighborhoods_shortpath_csr(net["indptr"], net["indices"], net["csr_length"], nr)
    got = nb.dense()
    assert np.array_equal(got, ref)
    assert got.diagonal().all()
    # shortest-path neighborhoods of an undirected graph are symmetric up to fp64 path-order rounding;
    # at this cutoff no pair flips, so symmetry doubles as a structural check
    assert (got != got.T).sum() <= 4


def test_full_size_euclid_properties(ctx):
    """C4 shape (100k points): sampled rows against the row-wise oracle plus symmetry / diagonal / row-sum
    identities that need no reference."""
    c = syn.make_config("C4")
    x, y = c["net"]["x"], c["net"]["y"]
    n = x.size
    nr = c["radius"] * (np.max(x) - np.min(x))
    eu = _lib.Neighborhoods(ctx, n).euclid(x, y, nr)
    rows = np.random.default_rng(5).choice(n, 64, replace=False)
    ref = orc.neighborhoods_euclidean_rows(x, y, nr, rows)
    for k, r in enumerate(rows):
        assert np.array_equal(eu.dense(int(r), int(r) + 1)[0], ref[k])
    rs = eu.rowsums()
    blk = eu.dense(0, 2048)
    assert np.array_equal(blk.sum(axis=1), rs[:2048])
    assert np.array_equal(blk[:, :2048], blk[:, :2048].T)       # (dx*dx + dy*dy) is symmetric bit-for-bit
    assert blk[np.arange(2048), np.arange(2048)].all()
    w = eu.packed(0, 4)                                        # padding bits stay zero
    nw = (n + 31) // 32
    assert not w[:, nw:].any()
    if n % 32:
        assert not np.any(w[:, nw - 1] >> np.uint32(n % 32))
