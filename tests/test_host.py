"""CPU: host-side logic of the drop-in layer (settings, permutation stream, packed matrix view, CSR extraction)."""
import pickle

import numpy as np
import pandas as pd
import pytest

import safe_oracle as orc
from conftest import net_from_golden
from safepy_b200 import SAFE, PackedNeighborhoods, synthetic as syn
from safepy_b200._lib import pack_dense, unpack_packed
from safepy_b200.neighborhood_matrix import as_packed
from safepy_b200.permutations import iter_perm_rows, make_perm_rows, shard_bounds
from safepy_b200.safe import graph_csr


def test_defaults_match_reference_ini():
    sf = SAFE(verbose=False)
    assert sf.node_distance_metric == "shortpath_weighted_layout"
    assert sf.neighborhood_radius == 0.1 and sf.neighborhood_radius_type == "diameter"
    assert sf.background == "attribute_file" and sf.attribute_sign == "both"
    assert sf.num_permutations == 1000 and sf.neighborhood_score_type == "sum"
    assert sf.enrichment_type == "auto" and sf.enrichment_threshold == 0.05 and sf.random_seed is None


@pytest.mark.parametrize("attr,bad,default", [
    ("node_distance_metric", "manhattan", "shortpath_weighted_layout"),
    ("background", "genome", "attribute_file"),
    ("attribute_sign", "up", "both"),
    ("num_permutations", 5, 1000),
    ("enrichment_threshold", 2.0, 0.05),
])
def test_validate_config_raises_and_restores_default(attr, bad, default):
    """safepy/safe.py:190-235: ValueError and the setting snaps back to its default."""
    sf = SAFE(verbose=False)
    setattr(sf, attr, bad)
    with pytest.raises(ValueError):
        sf.validate_config()
    assert getattr(sf, attr) == default


def test_kwargs_are_sticky_and_validated_before_any_gpu_work():
    sf = SAFE(verbose=False)
    with pytest.raises(ValueError):
        sf.define_neighborhoods(node_distance_metric="bogus")
    assert sf.node_distance_metric == "shortpath_weighted_layout"
    with pytest.raises(ValueError):
        sf.compute_pvalues(background="bogus")


@pytest.mark.parametrize("kind", ["normal32", "single", "binary"])
def test_perm_rows_replay_the_reference_stream(stage2_small, kind):
    """Gather rows composed from the legacy RNG reproduce the reference's in-place cumulative shuffles: applying
    them to the original matrix and scoring on the CPU gives the reference's counts."""
    g = stage2_small
    attrs = g["attr_" + kind]
    P, seed = int(g["num_permutations"]), int(g["seed"])
    rows = make_perm_rows(attrs, P, seed)
    assert rows.dtype == np.int32 and rows.shape == (P, attrs.shape[0])
    assert np.array_equal(rows, orc.perm_gather_rows(attrs, P, seed))
    # each row is a permutation and leaves all-NaN rows in place
    nodata = np.all(np.isnan(attrs), axis=1)
    for p in (0, P - 1):
        assert np.array_equal(np.sort(rows[p]), np.arange(attrs.shape[0]))
        assert np.array_equal(rows[p][nodata], np.nonzero(nodata)[0])
    nb = unpack_packed(g["neighborhoods"], attrs.shape[0]).astype(np.int64)
    cneg, cpos = orc.perm_counts_from_rows(nb, attrs, "sum", rows)
    assert np.array_equal(cneg, g["cneg_%s_sum" % kind]) and np.array_equal(cpos, g["cpos_%s_sum" % kind])


def test_shard_bounds_cover_all_permutations():
    for P in (10, 1000, 1001):
        for ws in (1, 2, 3, 8):
            spans = [shard_bounds(P, ws, r) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == P
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_packed_neighborhoods_behaves_like_the_dense_matrix(stage1_small):
    g = stage1_small
    n = g["x"].shape[0]
    dense = unpack_packed(g["nb_layout"], n).astype(np.int64)
    pk = PackedNeighborhoods(g["nb_layout"], n)
    assert pk.shape == (n, n)
    assert np.array_equal(np.sum(pk, axis=1), dense.sum(axis=1))      # reference tests' usage
    assert np.array_equal(np.asarray(pk), dense)
    v = np.arange(n, dtype=np.float64)
    assert np.array_equal(np.dot(pk, v), np.dot(dense, v))
    assert pk[3, 3] == 1 and np.array_equal(pk[5], dense[5])
    clone = pickle.loads(pickle.dumps(pk))
    assert np.array_equal(clone.words, pk.words) and clone._device is None
    assert np.array_equal(as_packed(dense).words, pk.words)
    with pytest.raises(ValueError):
        as_packed(dense * 2)


def test_graph_csr_follows_networkx_weight_rule(stage1_small):
    net = net_from_golden(stage1_small)
    g = syn.to_networkx(net)
    ip, ix, w = graph_csr(g, "length")
    assert np.array_equal(ip, net["indptr"]) and np.array_equal(ix, net["indices"])
    assert np.array_equal(w, net["csr_length"])
    ip2, ix2, w2 = graph_csr(g, "weight")          # attribute absent -> every edge costs 1
    assert np.array_equal(ix2, ix) and np.all(w2 == 1.0)
    ipo, ixo, wo = orc.graph_to_csr(g, "length")
    assert np.array_equal(ipo, ip) and np.array_equal(ixo, ix) and np.array_equal(wo, w)


def test_streamed_perm_rows_equal_the_one_shot_replay():
    rng = np.random.default_rng(5)
    attrs = rng.standard_normal((57, 3)).astype(np.float32)
    attrs[rng.random(57) < 0.2] = np.nan
    full = make_perm_rows(attrs, 23, 11)
    state_after = np.random.get_state()[1].copy()
    pieces = list(iter_perm_rows(attrs, 23, 11, piece=8, depth=2, first=2))
    assert [p.shape[0] for p in pieces] == [2, 4, 8, 8, 1]
    assert np.array_equal(np.concatenate(pieces), full)
    assert np.array_equal(np.random.get_state()[1], state_after)      # the global stream ends where upstream's does
    # abandoning the iterator must not leave the producer thread blocked
    it = iter_perm_rows(attrs, 1000, 11, piece=2, depth=1)
    next(it)
    it.close()
    assert not it._worker.is_alive()
    iter_perm_rows(attrs, 1000, 11, piece=2, depth=1).close()        # never iterated


def test_packed_row_sums_on_the_host(stage1_small):
    g = stage1_small
    n = g["x"].shape[0]
    pk = PackedNeighborhoods(g["nb_layout"], n)
    assert np.array_equal(pk.row_sums(), unpack_packed(g["nb_layout"], n).sum(axis=1))


def test_loaders_accept_arrays_and_frames(stage1_small):
    net = net_from_golden(stage1_small)
    sf = SAFE(verbose=False)
    # explicit lengths: nothing to compute (without them the lengths come from the device, tests/test_next_rows.py)
    sf.load_network(edges=net["edges"], x=net["x"], y=net["y"], length=net["length"])
    assert sf.graph.number_of_nodes() == net["n"] and sf.graph.number_of_edges() == len(net["edges"])
    _, _, w = graph_csr(sf.graph, "length")
    assert np.array_equal(w, net["csr_length"])
    frame = pd.DataFrame(np.arange(2 * net["n"], dtype=float).reshape(net["n"], 2), index=[str(i) for i in range(net["n"])],
                         columns=["a", "b"])
    sf.load_attributes(attribute_file=frame)
    assert sf.node2attribute.shape == (net["n"], 2) and list(sf.attributes["name"]) == ["a", "b"]


def test_synthetic_configs_are_deterministic():
    a = syn.make_config("C1", scale=0.1)
    b = syn.make_config("C1", scale=0.1)
    assert np.array_equal(a["net"]["edges"], b["net"]["edges"]) and np.array_equal(a["net"]["x"], b["net"]["x"])
    assert np.array_equal(a["attributes"], b["attributes"], equal_nan=True)
    assert a["attributes"].dtype == np.float32


def test_kd_order_and_relabelling():
    from safepy_b200.ordering import kd_order, graph_order
    c0 = syn.make_config("C1", 0.15)
    c1 = syn.make_config("C1", 0.15, shuffle=True)
    n = c0["n"]
    for c in (c0, c1):
        o = kd_order(c["net"]["x"], c["net"]["y"])
        assert o.dtype == np.int32 and np.array_equal(np.sort(o), np.arange(n))
    # relabelling keeps the geometry: same multiset of coordinates and edge lengths, same attribute rows
    assert np.allclose(np.sort(c0["net"]["x"]), np.sort(c1["net"]["x"]))
    assert np.allclose(np.sort(c0["net"]["length"]), np.sort(c1["net"]["length"]))
    assert np.array_equal(np.sort(np.nan_to_num(c0["attributes"][:, 0])), np.sort(np.nan_to_num(c1["attributes"][:, 0])))
    e = c1["net"]["edges"]
    d = np.hypot(c1["net"]["x"][e[:, 0]] - c1["net"]["x"][e[:, 1]], c1["net"]["y"][e[:, 0]] - c1["net"]["y"][e[:, 1]])
    assert np.allclose(d, c1["net"]["length"])
    g = graph_order(c1["net"]["indptr"], c1["net"]["indices"], n)
    assert np.array_equal(np.sort(g), np.arange(n))


@pytest.mark.parametrize("n,perms,seed", [(57, 23, 11), (1, 3, 0), (2, 5, 5), (5, 4, 2 ** 32 - 1), (300, 10, 7),
                                          (5000, 6, 20240), (700, 9, None)])
def test_native_perm_stream_replays_numpy(n, perms, seed):
    """sb_perm_stream_* (MT19937 + legacy shuffle in C++) against the oracle's NumPy replay of safe_extras.py:46-58:
    same gather rows, and the same generator state afterwards."""
    from safepy_b200 import _lib
    rng = np.random.default_rng(n)
    attrs = rng.standard_normal((n, 2)).astype(np.float32)
    attrs[rng.random(n) < 0.25] = np.nan
    idx = np.nonzero(np.sum(~np.isnan(attrs), axis=1))[0]
    stream = _lib.PermStream(n, idx, seed)
    if seed is None:                                   # OS entropy: only the structure can be checked
        rows = stream.next(perms)
        assert all(sorted(r) == list(range(n)) for r in rows)
        moved = np.setdiff1d(np.arange(n), idx)
        assert np.array_equal(rows[:, moved], np.tile(moved, (perms, 1)))
        return
    ref = orc.perm_gather_rows(attrs, perms, seed)
    state = np.random.get_state()
    first = stream.next(perms // 2)
    stream.skip(1)                                     # a skipped permutation still advances the stream
    rest = stream.next(perms - perms // 2 - 1)
    assert np.array_equal(first, ref[:perms // 2]) and np.array_equal(rest, ref[perms // 2 + 1:])
    key, pos, drawn = stream.state()
    assert drawn == perms and pos == state[2] and np.array_equal(key, state[1])
    np.random.seed(12345)
    stream.sync_numpy()
    assert np.array_equal(np.random.get_state()[1], state[1]) and np.random.get_state()[2] == state[2]
    # make_perm_rows goes through the same stream and leaves NumPy's generator where upstream would
    np.random.seed(999)
    assert np.array_equal(make_perm_rows(attrs, perms, seed), ref)
    assert np.array_equal(np.random.get_state()[1], state[1])


def test_perm_stream_read_ahead():
    """sb_perm_stream_prefetch: the rows drawn in the background are the stream's next rows, the generator ends where
    the synchronous replay ends, and a request that does not match the read-ahead is refused."""
    from safepy_b200 import _lib
    n, perms, seed = 900, 37, 5
    rng = np.random.default_rng(3)
    attrs = rng.standard_normal((n, 2)).astype(np.float32)
    attrs[rng.random(n) < 0.2] = np.nan
    idx = np.nonzero(np.sum(~np.isnan(attrs), axis=1))[0]
    ref = orc.perm_gather_rows(attrs, perms, seed)
    state = np.random.get_state()
    stream = _lib.PermStream(n, idx, seed).prefetch(perms)
    assert np.array_equal(stream.next(perms), ref)
    key, pos, drawn = stream.state()
    assert drawn == perms and pos == state[2] and np.array_equal(key, state[1])
    stream.close()
    # a rank's share of a dealt null: the state afterwards is that of ALL permutations (foreign pieces are drawn too)
    stream = _lib.PermStream(n, idx, seed).prefetch(perms, world=2, rank=1)
    with pytest.raises(_lib.SafeB200Error, match="another request"):
        stream.next(perms)
    key, pos, drawn = stream.state()                   # waits for the read-ahead
    assert drawn == perms and pos == state[2] and np.array_equal(key, state[1])
    stream.close()
    _lib.PermStream(n, idx, seed).prefetch(perms).close()      # closing with a read-ahead in flight joins it


def test_perm_stream_rejects_bad_input():
    from safepy_b200 import _lib
    with pytest.raises(ValueError):
        _lib.PermStream(10, np.arange(10), -1)
    with pytest.raises(_lib.SafeB200Error, match="out of range"):
        _lib.PermStream(10, np.array([3, 10]), 1)
    from safepy_b200.permutations import native_seed
    assert native_seed(None) and native_seed(7) and native_seed(np.int64(7))
    assert not native_seed(2 ** 32) and not native_seed([1, 2]) and not native_seed(True)
