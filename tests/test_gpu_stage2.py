"""GPU: stage-2 kernels (through the C ABI) against the reference goldens and the oracle.
Counts are bit-exact for identical host-generated permutation indices; hypergeometric NES within 1e-6 relative
(absolute 1e-12 below 1e-3, same inf / NaN / -0.0 positions) -- SURVEY.md section 8d."""
import numpy as np
import pytest

import safe_oracle as orc
from conftest import net_from_golden
from safepy_b200 import _lib, synthetic as syn
from safepy_b200._lib import unpack_packed
from safepy_b200.permutations import make_perm_rows

pytestmark = pytest.mark.gpu

KINDS = ["normal32", "dyadic", "binary", "normal64", "single"]


def assert_nes_close(got, ref):
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    big = np.isinf(ref) | (ref > 300)
    assert np.array_equal(np.isinf(got) | (got > 300), big)
    ok = ~np.isnan(ref) & ~big
    g, r = got[ok], ref[ok]
    hi = np.abs(r) >= 1e-3
    assert np.all(np.abs(g[hi] - r[hi]) <= 1e-6 * np.abs(r[hi]))
    assert np.all(np.abs(g[~hi] - r[~hi]) <= 1e-12)


@pytest.fixture(scope="module")
def small(ctx, stage2_small):
    g = stage2_small
    n = g["x"].shape[0]
    nb = _lib.Neighborhoods(ctx, n).upload_packed(g["neighborhoods"])
    return g, n, nb


@pytest.mark.parametrize("kind", KINDS)
def test_scores_match_reference(small, kind):
    g, n, nb = small
    plan = _lib.Enrichment(nb, g["attr_" + kind])
    s = plan.score("sum")
    ref = g["ns_%s_sum" % kind]
    if kind == "normal64":
        assert np.allclose(s, ref, rtol=1e-13, atol=1e-13)
    else:
        assert np.array_equal(s, ref)
    z = plan.score("z-score")
    refz = g["ns_%s_z" % kind]
    assert np.array_equal(np.isnan(z), np.isnan(refz))
    assert np.allclose(z, refz, rtol=1e-9, atol=1e-12, equal_nan=True)


@pytest.mark.parametrize("engine", ["simt", "tc"])
@pytest.mark.parametrize("kind", KINDS)
def test_perm_counts_bit_exact(small, kind, engine):
    g, n, nb = small
    attrs = g["attr_" + kind]
    rows = make_perm_rows(attrs, int(g["num_permutations"]), int(g["seed"]))
    plan = _lib.Enrichment(nb, attrs)
    cneg, cpos = plan.perm_counts(rows, "sum", engine)
    assert np.array_equal(cneg, g["cneg_%s_sum" % kind])
    assert np.array_equal(cpos, g["cpos_%s_sum" % kind])
    if engine == "tc":
        st = plan.stats()
        assert st["decided"] + st["fixups"] == n * attrs.shape[1] * rows.shape[0]
        if kind in ("binary", "dyadic"):
            assert st["fixups"] == 0          # exactly representable data never needs the fp64 fix-up


@pytest.mark.parametrize("kind", ["normal32", "binary", "single"])
def test_zscore_counts(small, kind):
    g, n, nb = small
    attrs = g["attr_" + kind]
    rows = make_perm_rows(attrs, int(g["num_permutations"]), int(g["seed"]))
    cneg, cpos = _lib.Enrichment(nb, attrs).perm_counts(rows, "z-score", "auto")
    # z-score comparisons are not summation-order exact: allow a handful of rounding-level flips
    assert np.abs(cneg.astype(int) - g["cneg_%s_z" % kind]).sum() <= 3
    assert np.abs(cpos.astype(int) - g["cpos_%s_z" % kind]).sum() <= 3


def test_hypergeom_matches_reference(small):
    g, n, nb = small
    p, nes = _lib.Enrichment(nb, g["attr_binary"]).hypergeom()
    assert_nes_close(nes, g["hyper_nes"])
    assert np.array_equal(np.isnan(p), np.isnan(g["hyper_p"]))
    assert np.array_equal(p == 1.0, g["hyper_p"] == 1.0)
    assert np.array_equal(orc.nes_binary(nes, 0.05), g["hyper_nesbin"])
    b0 = np.where(np.isnan(g["attr_binary"]), 0, g["attr_binary"]).astype(np.float32)
    _, nes0 = _lib.Enrichment(nb, b0).hypergeom()
    assert_nes_close(nes0, g["hyper_bgnet_nes"])


def test_hypergeom_invalid_parameters_give_nan(small):
    """Non-integer group sizes fail scipy's argcheck -> NaN (safe.py:596 through rv_discrete.sf)."""
    g, n, nb = small
    attrs = g["attr_normal32"]
    dense = unpack_packed(g["neighborhoods"], n).astype(np.int64)
    pref, nref = orc.hypergeom_pvalues(dense, attrs)
    p, nes = _lib.Enrichment(nb, attrs).hypergeom()
    assert np.array_equal(np.isnan(p), np.isnan(pref))


def test_mid_size_counts_against_oracle(ctx, stage1_mid):
    """1500 nodes, 70 float32 attributes (two column groups incl. a ragged one), 40 permutations:
    tensor-core counts == SIMT counts == NumPy oracle."""
    g = stage1_mid
    n = g["x"].shape[0]
    nb = _lib.Neighborhoods(ctx, n).upload_packed(g["nb_layout"])
    attrs = syn.make_attributes(n, 70, 5, "normal32")
    rows = make_perm_rows(attrs, 40, 3)
    plan = _lib.Enrichment(nb, attrs)
    tneg, tpos = plan.perm_counts(rows, "sum", "tc")
    st = plan.stats()
    sneg, spos = plan.perm_counts(rows, "sum", "simt")
    dense = unpack_packed(g["nb_layout"], n).astype(np.int64)
    oneg, opos = orc.perm_counts_from_rows(dense, attrs, "sum", rows)
    assert np.array_equal(sneg, oneg) and np.array_equal(spos, opos)
    assert np.array_equal(tneg, oneg) and np.array_equal(tpos, opos)
    assert st["digits"] == 3 and st["a_tiles"] <= st["a_tiles_dense"]


def test_fixed_point_scale_at_the_edges_of_the_digit_range(ctx, stage1_mid):
    """The scale of an inexact column comes from its exact maximum (one bit more than the exponent alone would give)
    unless that maximum would leave the three balanced digits (|q| <= 8355711): columns whose largest magnitude sits
    just below / above that limit, at a power of two, and negative, for the 'sum' and the z-score null (whose squares
    have their own two-digit scale)."""
    g = stage1_mid
    n = g["x"].shape[0]
    nb = _lib.Neighborhoods(ctx, n).upload_packed(g["nb_layout"])
    rng = np.random.default_rng(11)
    attrs = rng.uniform(-0.45, 0.45, size=(n, 64)).astype(np.float32)
    tops = [8355711 / 2 ** 23, 8355712 / 2 ** 23, 0.998, 0.9999999, 1.0, 0.5, 0.50000006, 1.9921, 3.0, 1e-3, 7.96875]
    for j, top in enumerate(tops):
        attrs[rng.integers(n), j] = np.float32(top) * (-1 if j % 2 else 1)
        attrs[rng.integers(n), j + 16] = np.float32(np.sqrt(top))     # the same edges for the squares
    attrs[rng.uniform(size=attrs.shape) < 0.01] = np.nan
    rows = make_perm_rows(attrs, 30, 5)
    plan = _lib.Enrichment(nb, attrs)
    for score in ("sum", "z-score"):
        tneg, tpos = plan.perm_counts(rows, score, "tc")
        sneg, spos = plan.perm_counts(rows, score, "simt")
        assert np.array_equal(tneg, sneg) and np.array_equal(tpos, spos), score
    plan.close()


@pytest.mark.parametrize("m", [1, 2, 3, 16, 33, 64, 65, 128])
def test_column_group_shapes(ctx, stage1_small, m):
    """Every column-packing regime of the GEMM (several permutations per 64-column slot for m < 64, ragged last
    group for m % 64 != 0), binary and continuous data."""
    g = stage1_small
    n = g["x"].shape[0]
    nb = _lib.Neighborhoods(ctx, n).upload_packed(g["nb_layout"])
    for kind in ("normal32", "binary"):
        attrs = syn.make_attributes(n, m, 100 + m, kind)
        rows = make_perm_rows(attrs, 37, 11)
        plan = _lib.Enrichment(nb, attrs)
        tneg, tpos = plan.perm_counts(rows, "sum", "tc")
        sneg, spos = plan.perm_counts(rows, "sum", "simt")
        assert np.array_equal(tneg, sneg) and np.array_equal(tpos, spos), (m, kind)


def test_counts_properties_without_reference(ctx, stage1_small):
    """Size-independent identities: counts_neg + counts_pos = P + #ties >= P; the identity permutation ties
    everywhere; all-NaN attribute rows never move."""
    g = stage1_small
    n = g["x"].shape[0]
    nb = _lib.Neighborhoods(ctx, n).upload_packed(g["nb_layout"])
    attrs = syn.make_attributes(n, 20, 77, "normal32")
    ident = np.tile(np.arange(n, dtype=np.int32), (12, 1))
    plan = _lib.Enrichment(nb, attrs)
    for engine in ("tc", "simt"):
        cneg, cpos = plan.perm_counts(ident, "sum", engine)
        assert np.all(cneg == 12) and np.all(cpos == 12)
    rows = make_perm_rows(attrs, 25, 5)
    cneg, cpos = plan.perm_counts(rows, "sum", "tc")
    assert np.all(cneg.astype(int) + cpos.astype(int) >= 25)
    assert np.all(cneg <= 25) and np.all(cpos <= 25)


@pytest.mark.parametrize("m", [3, 70])
def test_node_order_hint_does_not_change_counts(ctx, stage1_mid, m):
    """Any internal node order (identity, random, k-d tree of the layout) must give the same counts
    (m = 3 exercises the packed small-M slots, m = 70 two column groups)."""
    from safepy_b200.ordering import kd_order
    from safepy_b200.permutations import make_perm_rows
    g = stage1_mid
    n = g["x"].shape[0]
    rng = np.random.default_rng(5)
    attrs = rng.standard_normal((n, m)).astype(np.float32)
    attrs[rng.uniform(size=n) < 0.05] = np.nan
    nb = _lib.Neighborhoods(ctx, n).upload_packed(g["nb_layout"])
    rows = make_perm_rows(attrs, 70 if m < 64 else 9, 3)
    ref = _lib.Enrichment(nb, attrs).perm_counts(rows, "sum", "simt")
    for order in (None, rng.permutation(n), kd_order(g["x"], g["y"])):
        plan = _lib.Enrichment(nb, attrs)
        plan.set_node_order(order)
        got = plan.perm_counts(rows, "sum", "tc")
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])
        plan.close()
    with pytest.raises(_lib.SafeB200Error):
        _lib.Enrichment(nb, attrs).set_node_order(np.zeros(n, dtype=np.int32))


def test_full_size_c3_tensor_core_equals_exact_engine(ctx):
    """BASELINE.json configs[2] at full size (20k nodes, 2000 float32 attributes, shuffled node numbering + k-d
    order hint): the tensor-core null must agree with the exact fp64 SIMT engine cell for cell, and obey the
    size-independent identities.  (The SIMT engine itself is pinned against the oracle / reference at small sizes.)"""
    from safepy_b200.ordering import kd_order
    cfg = syn.make_config("C3", shuffle=True)
    net, n, m = cfg["net"], cfg["n"], cfg["m"]
    attrs = cfg["attributes"]
    nr = cfg["radius"] * (np.max(net["x"]) - np.min(net["x"]))
    nb = _lib.Neighborhoods(ctx, n).shortpath(net["indptr"], net["indices"], net["csr_length"], nr)
    # stage 1 at full size: sampled rows against scipy's Dijkstra
    rows_s = np.random.default_rng(1).choice(n, 64, replace=False)
    ref = orc.neighborhoods_shortpath_csr(net["indptr"], net["indices"], net["csr_length"], nr, rows=rows_s)
    got = unpack_packed(nb.packed(), n)[rows_s]
    assert np.array_equal(got, ref)
    P = 5
    rows = make_perm_rows(attrs, P, 7)
    plan = _lib.Enrichment(nb, attrs).set_node_order(kd_order(net["x"], net["y"]))
    tneg, tpos = plan.perm_counts(rows, "sum", "tc")
    st = plan.stats()
    sneg, spos = plan.perm_counts(rows, "sum", "simt")
    assert np.array_equal(tneg, sneg) and np.array_equal(tpos, spos)
    assert np.all(tneg.astype(int) + tpos.astype(int) >= P) and tneg.max() <= P and tpos.max() <= P
    assert st["a_tiles"] * 4 < st["a_tiles_dense"], "the order hint must leave most tiles empty"
    # rows without data never move, so a node whose whole neighborhood has no data ties in every permutation
    plan.close()
    nb.close()
