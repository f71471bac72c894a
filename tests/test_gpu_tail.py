"""GPU: what follows the counts inside compute_pvalues (streaming null, fused p-value / NES / nes_binary tail,
row-wise FDR) and the Jaccard distances of define_domains, through the C ABI, against the reference goldens and the
oracle.  Without FDR every output is bit-identical (p-values and NES of the possible counts are host-made tables);
with FDR the adjusted p-values are bit-identical to the oracle's restatement of statsmodels' formula and the NES,
a device log10 of them, is held to 1e-12 relative (the north star allows 1e-6)."""
import numpy as np
import pandas as pd
import pytest

import safe_oracle as orc
from conftest import load_golden, net_from_golden
from safepy_b200 import SAFE, _lib, synthetic as syn
from safepy_b200._lib import unpack_packed
from safepy_b200.permutations import make_perm_rows

pytestmark = pytest.mark.gpu

KINDS = ["normal32", "dyadic", "binary", "normal64", "single"]


@pytest.fixture(scope="module")
def small(ctx, stage2_small):
    g = stage2_small
    n = g["x"].shape[0]
    nb = _lib.Neighborhoods(ctx, n).upload_packed(g["neighborhoods"])
    return g, n, nb


def stream_null(plan, rows, pieces, score_type="sum", engine="auto"):
    plan.null_begin(score_type, engine)
    for part in np.array_split(rows, pieces):
        plan.null_add(part)
    return plan


@pytest.mark.parametrize("kind", KINDS)
def test_streamed_null_and_tail_match_reference(small, kind):
    g, n, nb = small
    attrs = g["attr_" + kind]
    P = int(g["num_permutations"])
    rows = make_perm_rows(attrs, P, int(g["seed"]))
    plan = stream_null(_lib.Enrichment(nb, attrs), rows, 3)
    num, cneg, cpos = plan.null_counts()
    assert num == P
    assert np.array_equal(cneg, g["cneg_%s_sum" % kind]) and np.array_equal(cpos, g["cpos_%s_sum" % kind])
    st = plan.stats()
    assert st["decided"] + st["fixups"] == n * attrs.shape[1] * P          # accumulated over the pieces
    out = plan.null_finalize(P, "both", 0.05)
    if kind == "normal64":
        assert np.allclose(out["ns"], g["ns_%s_sum" % kind], rtol=1e-13, atol=1e-13, equal_nan=True)
    else:
        assert np.array_equal(out["ns"], g["ns_%s_sum" % kind], equal_nan=True)
    assert np.array_equal(out["pvalues_neg"], g["rand_pneg_" + kind], equal_nan=True)
    assert np.array_equal(out["pvalues_pos"], g["rand_ppos_" + kind], equal_nan=True)
    assert np.array_equal(out["nes"], g["rand_nes_" + kind], equal_nan=True)
    assert np.array_equal(np.signbit(out["nes"]), np.signbit(g["rand_nes_" + kind]))   # -0.0 where numpy has it
    assert np.array_equal(out["nes_binary"], g["rand_nesbin_" + kind])
    assert np.array_equal(out["num_neighborhoods_enriched"], g["rand_enriched_" + kind])


@pytest.mark.parametrize("sign", ["highest", "lowest", "both"])
@pytest.mark.parametrize("score_type", ["sum", "z-score"])
def test_tail_signs_and_nan_scores(small, sign, score_type):
    """z-scores are NaN for small / constant neighborhoods: counts there become NaN p-values and nes_binary 0."""
    g, n, nb = small
    attrs = np.ascontiguousarray(g["attr_binary"][:, 2:5])        # the middle column is all 0: its z-score is NaN
    rows = make_perm_rows(attrs, 30, 5)
    plan = stream_null(_lib.Enrichment(nb, attrs), rows, 2, score_type)
    _, cneg, cpos = plan.null_counts()
    out = plan.null_finalize(30, sign, 0.1)
    pn, pp, nes, nbin, enriched = orc.randomization_tail(out["ns"], cneg, cpos, 30, sign, False, 0.1)
    if score_type == "z-score":
        assert np.isnan(out["ns"][:, 1]).all() and np.isnan(out["nes"][:, 1]).all()
        assert not out["nes_binary"][:, 1].any()
    for key, ref in (("pvalues_neg", pn), ("pvalues_pos", pp), ("nes", nes), ("nes_binary", nbin)):
        assert np.array_equal(out[key], ref, equal_nan=True), key
    assert np.array_equal(out["num_neighborhoods_enriched"], enriched)


@pytest.mark.parametrize("kind,perms", [("normal32", 60), ("single", 300), ("binary", 7)])
def test_null_from_native_stream(small, kind, perms):
    """sb_enrich_null_add_stream (RNG replay on a producer thread inside the C call) against the explicit rows."""
    from safepy_b200.permutations import perm_stream
    g, n, nb = small
    attrs = g["attr_" + kind]
    plan = _lib.Enrichment(nb, attrs)
    cneg0, cpos0 = plan.perm_counts(make_perm_rows(attrs, perms, 3))
    plan.null_begin("sum", "auto")
    stream = perm_stream(attrs, 3)
    plan.null_add_stream(stream, perms // 3)
    plan.null_add_stream(stream, perms - perms // 3)          # a stream can be drained in several calls
    num, cneg, cpos = plan.null_counts()
    assert num == perms and stream.state()[2] == perms
    assert np.array_equal(cneg, cneg0) and np.array_equal(cpos, cpos0)
    other = _lib.PermStream(n + 1, np.arange(n + 1), 3)
    with pytest.raises(_lib.SafeB200Error, match="the stream permutes"):
        plan.null_add_stream(other, 1)


@pytest.mark.parametrize("world,perms", [(2, 60), (3, 100), (8, 37)])
def test_round_robin_shards_add_up(small, world, perms):
    """sb_enrich_null_add_stream_shard: every rank draws the whole stream and counts the pieces dealt to it; here the
    ranks run one after the other into the same count arrays."""
    from safepy_b200.permutations import perm_stream
    g, n, nb = small
    attrs = g["attr_normal32"]
    plan = _lib.Enrichment(nb, attrs)
    cneg0, cpos0 = plan.perm_counts(make_perm_rows(attrs, perms, 5))
    plan.null_begin("sum", "auto")
    shares = []
    for rank in range(world):
        stream = perm_stream(attrs, 5)
        before = plan.null_counts(False)[0]
        plan.null_add_stream(stream, perms, world, rank)
        shares.append(plan.null_counts(False)[0] - before)
        assert stream.state()[2] == perms                       # the whole stream was drawn
    assert sum(shares) == perms and (world > perms // 8 or min(shares) > 0)
    _, cneg, cpos = plan.null_counts()
    assert np.array_equal(cneg, cneg0) and np.array_equal(cpos, cpos0)


def test_tail_needs_a_table_entry_per_count(small):
    g, n, nb = small
    attrs = g["attr_dyadic"]
    plan = stream_null(_lib.Enrichment(nb, attrs), make_perm_rows(attrs, 12, 1), 1)
    with pytest.raises(_lib.SafeB200Error, match="permutations were counted"):
        plan.null_finalize(10)
    with pytest.raises(_lib.SafeB200Error, match="null_begin"):
        _lib.Enrichment(nb, attrs).null_add(make_perm_rows(attrs, 2, 1))


def test_large_outputs_go_through_the_pinned_ring(ctx):
    """12 MB per output array: several 4 MB pieces through the worker threads; compared with the one-shot host call."""
    c = syn.make_config("C3", scale=0.075)                      # 1500 nodes
    net = c["net"]
    n = net["n"]
    attrs = np.random.default_rng(4).standard_normal((n, 1000)).astype(np.float32)
    attrs[::17] = np.nan
    nr = 0.1 * (net["x"].max() - net["x"].min())
    nb = _lib.Neighborhoods(ctx, n).shortpath(net["indptr"], net["indices"], net["csr_length"], nr)
    rows = make_perm_rows(attrs, 20, 9)
    plan = _lib.Enrichment(nb, attrs)
    cneg0, cpos0 = plan.perm_counts(rows)
    stream_null(plan, rows, 4)
    _, cneg, cpos = plan.null_counts()
    assert np.array_equal(cneg, cneg0) and np.array_equal(cpos, cpos0)
    out = plan.null_finalize(20, "both", 0.05)
    pn, pp, nes, nbin, enriched = orc.randomization_tail(out["ns"], cneg, cpos, 20, "both", False, 0.05)
    assert np.array_equal(out["ns"], plan.score("sum"))
    for key, ref in (("pvalues_neg", pn), ("pvalues_pos", pp), ("nes", nes), ("nes_binary", nbin)):
        assert np.array_equal(out[key], ref, equal_nan=True), key
    assert np.array_equal(out["num_neighborhoods_enriched"], enriched)


# ------------------------------------------------------------------------------------------------ FDR
@pytest.mark.parametrize("shape", [(1, 1), (3, 2), (40, 257), (17, 5000), (300, 64)])
def test_fdr_rows_bit_exact(ctx, shape):
    rng = np.random.default_rng(shape[1])
    p = rng.uniform(size=shape)
    p[rng.uniform(size=shape) < 0.2] = 0.0                       # ties at 0 (randomization p-values are discrete)
    p = np.where(rng.uniform(size=shape) < 0.3, np.round(p, 2), p)
    if shape[0] > 2:
        p[1, shape[1] // 2] = np.nan                             # one NaN poisons its row
        p[2] = 1.0
    adj = _lib.fdr_rows(ctx, p)
    assert np.array_equal(adj, orc.fdr_rows(p), equal_nan=True)


def test_randomization_with_fdr(small):
    g, n, nb = small
    attrs = g["attr_binary"]                                     # 30 attributes: a row worth adjusting
    P = int(g["num_permutations"])
    plan = stream_null(_lib.Enrichment(nb, attrs), make_perm_rows(attrs, P, int(g["seed"])), 2)
    _, cneg, cpos = plan.null_counts()
    out = plan.null_finalize(P, "both", 0.05, multiple_testing=True)
    pn, pp, nes, nbin, enriched = orc.randomization_tail(out["ns"], cneg, cpos, P, "both", True, 0.05)
    assert np.array_equal(out["pvalues_neg"], pn, equal_nan=True)
    assert np.array_equal(out["pvalues_pos"], pp, equal_nan=True)
    assert np.allclose(out["nes"], nes, rtol=1e-12, atol=1e-14, equal_nan=True)
    clear = np.abs(np.abs(nes) - -np.log10(0.05)) > 1e-9
    assert np.array_equal(out["nes_binary"][clear], nbin[clear])
    assert np.abs(out["num_neighborhoods_enriched"] - enriched).max() <= np.count_nonzero(~clear)


@pytest.mark.parametrize("fdr", [False, True])
def test_hypergeom_tail(small, fdr):
    g, n, nb = small
    attrs = g["attr_binary"]
    out = _lib.Enrichment(nb, attrs).hypergeom_finalize(0.05, multiple_testing=fdr)
    p0, _ = _lib.Enrichment(nb, attrs).hypergeom()
    if fdr:
        assert np.array_equal(out["pvalues_pos"], orc.fdr_rows(p0), equal_nan=True)
        with np.errstate(divide="ignore"):
            ref_nes = -np.log10(out["pvalues_pos"])
        assert np.allclose(out["nes"], ref_nes, rtol=1e-12, atol=1e-14, equal_nan=True)
    else:
        assert np.array_equal(out["pvalues_pos"], p0, equal_nan=True)
        ref_nes = g["hyper_nes"]
        ok = np.isfinite(ref_nes) & (np.abs(ref_nes) >= 1e-3)
        assert np.all(np.abs(out["nes"][ok] - ref_nes[ok]) <= 1e-6 * np.abs(ref_nes[ok]))
        assert np.array_equal(out["nes_binary"], g["hyper_nesbin"])
        assert np.array_equal(out["num_neighborhoods_enriched"], g["hyper_enriched"])
    assert np.array_equal(out["nes_binary"], orc.nes_binary(out["nes"], 0.05))
    assert np.array_equal(out["num_neighborhoods_enriched"], out["nes_binary"].sum(axis=0))


# ------------------------------------------------------------------------------------------------ domains
def test_jaccard_matches_scipy(ctx):
    g = load_golden("domains_small.npz")
    cols = np.flatnonzero(g["top"])
    assert np.array_equal(_lib.jaccard(ctx, g["nes_binary"], cols), g["jaccard"])
    rng = np.random.default_rng(8)
    for n, m in ((1, 3), (33, 5), (1000, 70), (4097, 9)):
        nbin = (rng.uniform(size=(n, m)) < 0.2).astype(np.float64)
        nbin[:, 0] = 0                                           # two empty columns: distance 0 by definition
        if m > 3:
            nbin[:, 3] = 0
        cols = rng.permutation(m)[: max(2, m - 1)]
        assert np.array_equal(_lib.jaccard(ctx, nbin, cols), orc.jaccard_condensed(nbin, cols))
    assert _lib.jaccard(ctx, nbin, [2]).shape == (0,)


def test_define_domains_matches_reference():
    g = load_golden("domains_small.npz")
    m = g["nes"].shape[1]
    sf = SAFE(verbose=False)
    sf.nes, sf.nes_binary = g["nes"], g["nes_binary"]
    sf.attributes = pd.DataFrame({"id": np.arange(m), "name": [str(j) for j in range(m)], "top": g["top"]})
    sf.define_domains()
    assert np.array_equal(sf.attributes["domain"].values, g["domain"])
    ids = list(g["domain_ids"])
    assert list(sf.node2domain.columns) == ids + ["primary_domain", "primary_nes"]
    assert np.array_equal(sf.node2domain[ids].values, g["node2domain"])
    assert np.array_equal(sf.node2domain["primary_domain"].values, g["primary_domain"])
    assert np.array_equal(sf.node2domain["primary_nes"].values, g["primary_nes"])
    sf.define_domains(attribute_distance_threshold=0.5)          # sticky override, different cut
    assert sf.attribute_distance_threshold == 0.5
    assert sf.attributes["domain"].max() >= g["domain"].max()


def test_safe_api_with_multiple_testing(stage2_small):
    g = stage2_small
    net = net_from_golden(g)
    sf = SAFE(verbose=False)
    sf.load_network(graph=syn.to_networkx(net))
    sf.random_seed = int(g["seed"])
    sf.define_neighborhoods(neighborhood_radius=float(g["radius"]))
    assert np.array_equal(np.sum(sf.neighborhoods, axis=1),
                          unpack_packed(g["neighborhoods"], net["n"]).sum(axis=1))      # device row sums
    assert np.array_equal(sf.neighborhoods.words, g["neighborhoods"])                   # fetched on first use
    sf.load_attributes(attribute_file=g["attr_binary"].copy())
    P = int(g["num_permutations"])
    sf.compute_pvalues(how="randomization", num_permutations=P, multiple_testing=True, verbose=False)
    pn, pp, nes, nbin, enriched = orc.randomization_tail(g["ns_binary_sum"], g["cneg_binary_sum"],
                                                         g["cpos_binary_sum"], P, "both", True, 0.05)
    assert np.array_equal(sf.pvalues_pos, pp, equal_nan=True) and np.array_equal(sf.pvalues_neg, pn, equal_nan=True)
    assert np.allclose(sf.nes, nes, rtol=1e-12, atol=1e-14, equal_nan=True)
    assert np.array_equal(sf.attributes["num_neighborhoods_enriched"].values, sf.nes_binary.sum(axis=0))
    sf.compute_pvalues(how="hypergeometric", multiple_testing=True, verbose=False)
    assert np.array_equal(sf.pvalues_pos, orc.fdr_rows(g["hyper_p"]), equal_nan=True) or \
        np.allclose(sf.pvalues_pos, orc.fdr_rows(g["hyper_p"]), rtol=1e-6, equal_nan=True)


@pytest.mark.parametrize("kind", ["normal32", "binary", "normal64"])
def test_attr_summary(small, kind):
    """safe.py:453-458 on the device: NaNs per column, values other than 0 / 1 / NaN."""
    g, n, nb = small
    attrs = g["attr_" + kind]
    nans, other = _lib.Enrichment(nb, attrs).attr_summary()
    mask = np.isnan(attrs)
    assert np.array_equal(nans, mask.sum(axis=0))
    assert other == np.sum(~mask & ~np.isin(attrs, [0, 1]))
