"""CPU: the oracle (oracle/safe_oracle.py) reproduces what the unmodified reference returned (tests/golden/,
written by oracle/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

import safe_oracle as orc
from conftest import load_golden, net_from_golden
from safepy_b200 import synthetic as syn
from safepy_b200._lib import pack_dense, unpack_packed

KINDS = ["normal32", "dyadic", "binary", "normal64", "single"]


@pytest.mark.parametrize("which", ["stage1_small", "stage1_mid"])
def test_stage1_oracle_matches_reference(which, request):
    g = request.getfixturevalue(which)
    net = net_from_golden(g)
    n = net["n"]
    graph = syn.to_networkx(net)
    # shortpath_weighted_layout: networkx restatement and the csgraph twin, both bit-exact
    nr = orc.neighborhood_radius(net["x"], float(g["r_layout"]), "shortpath_weighted_layout")
    ref = unpack_packed(g["nb_layout"], n)
    if n <= 400:
        assert np.array_equal(orc.neighborhoods_shortpath_nx(graph, nr, "length"), ref)
    assert np.array_equal(orc.neighborhoods_shortpath_csr(net["indptr"], net["indices"], net["csr_length"], nr), ref)
    # the csr extraction used by the oracle agrees with the generator's
    ip, ix, w = orc.graph_to_csr(graph, "length")
    assert np.array_equal(ip, net["indptr"]) and np.array_equal(ix, net["indices"]) and np.array_equal(w, net["csr_length"])
    # shortpath (hop count)
    ref = unpack_packed(g["nb_hops"], n)
    assert np.array_equal(orc.neighborhoods_shortpath_csr(net["indptr"], net["indices"], None, float(g["r_hops"])), ref)
    # shortpath with a 'weight' edge attribute
    _, _, wts = syn.edges_to_csr(n, g["edges"][:, 0], g["edges"][:, 1], g["edge_weight"])
    ref = unpack_packed(g["nb_weighted_hops"], n)
    assert np.array_equal(orc.neighborhoods_shortpath_csr(net["indptr"], net["indices"], wts, 3.0), ref)
    # euclidean: full pdist restatement and the row-wise twin
    nr = orc.neighborhood_radius(net["x"], float(g["r_euclid"]), "euclidean")
    ref = unpack_packed(g["nb_euclid"], n)
    assert np.array_equal(orc.neighborhoods_euclidean(net["x"], net["y"], nr), ref)
    rows = np.arange(0, n, 7)
    assert np.array_equal(orc.neighborhoods_euclidean_rows(net["x"], net["y"], nr, rows), ref[rows])
    assert ref.diagonal().all()


@pytest.mark.parametrize("kind", KINDS)
def test_scores_and_counts_match_reference(stage2_small, kind):
    g = stage2_small
    n = g["x"].shape[0]
    nb = unpack_packed(g["neighborhoods"], n).astype(np.int64)
    attrs = g["attr_" + kind]
    P, seed = int(g["num_permutations"]), int(g["seed"])
    for stype, tag in (("sum", "sum"), ("z-score", "z")):
        ns = orc.compute_neighborhood_score(nb, attrs, stype)
        assert np.array_equal(ns, g["ns_%s_%s" % (kind, tag)], equal_nan=True)
        cneg, cpos = orc.run_permutations(nb, attrs, stype, P, seed)
        assert np.array_equal(cneg, g["cneg_%s_%s" % (kind, tag)])
        assert np.array_equal(cpos, g["cpos_%s_%s" % (kind, tag)])
        # explicit gather indices give the same counts as the in-place cumulative shuffle
        rows = orc.perm_gather_rows(attrs, P, seed)
        cneg2, cpos2 = orc.perm_counts_from_rows(nb, attrs, stype, rows)
        assert np.array_equal(cneg2, g["cneg_%s_%s" % (kind, tag)])
        assert np.array_equal(cpos2, g["cpos_%s_%s" % (kind, tag)])


@pytest.mark.parametrize("kind", KINDS)
def test_randomization_nes_matches_reference(stage2_small, kind):
    g = stage2_small
    P = int(g["num_permutations"])
    pneg, ppos, nes = orc.randomization_nes(g["ns_%s_sum" % kind], g["cneg_%s_sum" % kind], g["cpos_%s_sum" % kind],
                                            P, "both")
    assert np.array_equal(pneg, g["rand_pneg_" + kind], equal_nan=True)
    assert np.array_equal(ppos, g["rand_ppos_" + kind], equal_nan=True)
    assert np.array_equal(nes, g["rand_nes_" + kind], equal_nan=True)
    nb = orc.nes_binary(nes, 0.05)
    assert np.array_equal(nb, g["rand_nesbin_" + kind])
    assert np.array_equal(nb.sum(axis=0), g["rand_enriched_" + kind])


def test_hypergeom_matches_reference(stage2_small):
    g = stage2_small
    n = g["x"].shape[0]
    nb = unpack_packed(g["neighborhoods"], n).astype(np.int64)
    p, nes = orc.hypergeom_pvalues(nb, g["attr_binary"])
    assert np.array_equal(p, g["hyper_p"], equal_nan=True)
    assert np.array_equal(nes, g["hyper_nes"], equal_nan=True)
    assert np.array_equal(orc.nes_binary(nes, 0.05), g["hyper_nesbin"])
    b0 = np.where(np.isnan(g["attr_binary"]), 0, g["attr_binary"])
    p, nes = orc.hypergeom_pvalues(nb, b0)
    assert np.array_equal(p, g["hyper_bgnet_p"], equal_nan=True)


def test_hypergeom_block_equals_the_full_statement(stage2_small):
    """The sampled-block form used for the full-size checks (bench.py parity gate, C2 test) reproduces the reference's
    cells bit for bit, also when a sampled column subset alone would misjudge which nodes carry data."""
    g = stage2_small
    n = g["x"].shape[0]
    nb = unpack_packed(g["neighborhoods"], n).astype(np.int64)
    rows = np.array([0, 3, 17, n - 1])
    cols = np.array([1, 4, 5])
    p, nes = orc.hypergeom_pvalues_block(nb[rows], g["attr_binary"], cols)
    assert np.array_equal(p, g["hyper_p"][np.ix_(rows, cols)], equal_nan=True)
    assert np.array_equal(nes, g["hyper_nes"][np.ix_(rows, cols)], equal_nan=True)
    attrs = g["attr_binary"].copy()
    attrs[:, cols] = np.where(np.arange(n)[:, None] % 7 == 0, np.nan, attrs[:, cols])   # NaN only in these columns
    pf, _ = orc.hypergeom_pvalues(nb, attrs)
    pb, _ = orc.hypergeom_pvalues_block(nb[rows], attrs, cols)
    assert np.array_equal(pb, pf[np.ix_(rows, cols)], equal_nan=True)


def test_sparse_twin_equals_dense_dot(stage2_small):
    g = stage2_small
    n = g["x"].shape[0]
    nb = unpack_packed(g["neighborhoods"], n)
    for kind in ("normal32", "dyadic", "binary"):
        s = orc.score_sum_csr(nb, g["attr_" + kind])
        assert np.array_equal(s, g["ns_%s_sum" % kind])


def test_pack_roundtrip():
    rng = np.random.default_rng(0)
    for n in (1, 31, 32, 33, 127, 130):
        d = (rng.uniform(size=(n, n)) < 0.3).astype(np.int64)
        w = pack_dense(d)
        assert w.shape[1] % 4 == 0
        assert np.array_equal(unpack_packed(w, n), d)


@pytest.mark.parametrize("kind", ["normal32", "binary", "single"])
def test_randomization_tail_matches_reference(stage2_small, kind):
    """safe.py:528-554 + 466-472 restated (oracle.randomization_tail) against the reference's recorded outputs."""
    g = stage2_small
    pn, pp, nes, nb, enriched = orc.randomization_tail(
        g["ns_%s_sum" % kind], g["cneg_%s_sum" % kind], g["cpos_%s_sum" % kind], int(g["num_permutations"]), "both",
        False, 0.05)
    assert np.array_equal(pn, g["rand_pneg_" + kind], equal_nan=True)
    assert np.array_equal(pp, g["rand_ppos_" + kind], equal_nan=True)
    assert np.array_equal(nes, g["rand_nes_" + kind], equal_nan=True)
    assert np.array_equal(nb, g["rand_nesbin_" + kind])
    assert np.array_equal(enriched, g["rand_enriched_" + kind])


def test_fdr_restatement_is_benjamini_hochberg():
    """statsmodels is not installed here (parity unpinned, see oracle.fdrcorrection): check the restatement against
    the textbook definition, adj_(k) = min_{k' >= k} p_(k') * m / k', its tie and NaN behaviour."""
    rng = np.random.default_rng(3)
    p = rng.uniform(size=(7, 40))
    p[1, 5:9] = p[1, 4]                                  # ties
    adj = orc.fdr_rows(p)
    for r in range(p.shape[0]):
        order = np.argsort(p[r], kind="stable")
        expect = np.empty(40)
        run = 1.0
        for rank in range(40, 0, -1):
            run = min(run, p[r][order[rank - 1]] * 40 / rank)
            expect[order[rank - 1]] = run
        assert np.allclose(adj[r], expect, rtol=0, atol=1e-15)
    assert np.all(adj >= p - 1e-18) and np.all(adj <= 1)
    assert len(set(adj[1, 4:9])) == 1
    # an independent implementation in an installed library: SciPy's Benjamini-Hochberg (p * m / k instead of
    # statsmodels' p / (k / m): equal up to the last bit)
    from scipy.stats import false_discovery_control
    big = rng.uniform(size=(20, 300))
    big[rng.uniform(size=big.shape) < 0.2] = 0.0
    big = np.where(rng.uniform(size=big.shape) < 0.3, np.round(big, 2), big)
    scipy_bh = np.stack([false_discovery_control(r, method="bh") for r in big])
    assert np.allclose(orc.fdr_rows(big), scipy_bh, rtol=1e-15, atol=1e-16)
    q = p.copy()
    q[3, 7] = np.nan
    adj2 = orc.fdr_rows(q)
    assert np.isnan(adj2[3]).all() and np.array_equal(np.delete(adj2, 3, 0), np.delete(adj, 3, 0))


def test_domains_oracle_matches_reference():
    g = load_golden("domains_small.npz")
    top = g["top"]
    assert np.array_equal(orc.jaccard_condensed(g["nes_binary"], np.flatnonzero(top)), g["jaccard"])
    domain, ids, counts, primary, primary_nes = orc.define_domains(g["nes"], g["nes_binary"], top,
                                                                   float(g["threshold"]))
    assert np.array_equal(domain, g["domain"]) and np.array_equal(ids, g["domain_ids"])
    assert np.array_equal(counts, g["node2domain"])
    assert np.array_equal(primary, g["primary_domain"])
    assert np.array_equal(primary_nes, g["primary_nes"])
