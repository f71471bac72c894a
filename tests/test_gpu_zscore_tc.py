"""GPU: the z-score permutation null on the tensor cores (value / square / non-NaN digit planes in one accumulation,
comparison in the epilogue + exact fix-ups) against the exact SIMT engine, which is pinned against the reference's recorded z-score counts
in tests/test_gpu_stage2.py::test_zscore_counts.  The two engines must agree cell for cell: decided comparisons are
rigorous with respect to the exact engine's value, undecided ones are re-evaluated by the exact engine's own code."""
import numpy as np
import pytest

import safe_oracle as orc
from safepy_b200 import _lib, synthetic as syn
from safepy_b200._lib import unpack_packed
from safepy_b200.ordering import kd_order
from safepy_b200.permutations import make_perm_rows

pytestmark = pytest.mark.gpu


def _attrs(n, m, kind, seed):
    rng = np.random.default_rng(seed)
    if kind == "normal64":
        b = rng.standard_normal((n, m))
        b[rng.uniform(size=(n, m)) < 0.02] = np.nan
        return b
    if kind == "integer":
        b = rng.integers(-5, 6, size=(n, m)).astype(np.float32)
        b[rng.uniform(size=n) < 0.05] = np.nan
        return b
    if kind.endswith("_full"):  # no NaN anywhere
        return np.nan_to_num(_attrs(n, m, kind[:-5], seed), nan=0.25)
    if kind == "mixed":       # exact and inexact columns side by side in one column group
        b = syn.make_attributes(n, m, seed, "normal32")
        b[:, ::3] = syn.make_attributes(n, m, seed + 1, "binary")[:, ::3]
        b[:, 1::7] = syn.make_attributes(n, m, seed + 2, "dyadic")[:, 1::7]
        return b
    return syn.make_attributes(n, m, seed, kind)


@pytest.mark.parametrize("kind", ["normal32", "binary", "dyadic", "integer", "normal64", "mixed", "normal32_full",
                                  "integer_full"])
@pytest.mark.parametrize("m", [64, 70, 130])
def test_zscore_tensor_core_equals_exact_engine(ctx, stage1_mid, kind, m):
    g = stage1_mid
    n = g["x"].shape[0]
    nb = _lib.Neighborhoods(ctx, n).upload_packed(g["nb_layout"])
    attrs = _attrs(n, m, kind, 31 + m)
    rows = make_perm_rows(attrs, 21, 3)
    plan = _lib.Enrichment(nb, attrs).set_node_order(kd_order(g["x"], g["y"]))
    tneg, tpos = plan.perm_counts(rows, "z-score", "tc")
    st = plan.stats()
    sneg, spos = plan.perm_counts(rows, "z-score", "simt")
    assert np.array_equal(tneg, sneg) and np.array_equal(tpos, spos), (kind, m)
    assert st["decided"] + st["fixups"] == n * m * rows.shape[0]
    if kind in ("binary", "integer", "integer_full"):
        assert st["fixups"] == 0            # exactly representable values and squares: no error band, no fix-ups
    else:
        assert st["fixups"] < 0.01 * n * m * rows.shape[0]
    aneg, apos = plan.perm_counts(rows, "z-score", "auto")        # 'auto' takes the tensor path for 64+ attributes
    assert np.array_equal(aneg, sneg) and np.array_equal(apos, spos)
    plan.close()


def test_zscore_against_the_oracle(ctx, stage1_mid):
    """z-score counts of the tensor path against the oracle's statement-for-statement restatement (np.dot in fp64):
    z-score comparisons are not summation-order exact, so a handful of rounding-level flips is allowed,
    exactly as for the SIMT engine against the reference's recorded counts."""
    g = stage1_mid
    n = g["x"].shape[0]
    nb = _lib.Neighborhoods(ctx, n).upload_packed(g["nb_layout"])
    attrs = _attrs(n, 64, "normal32", 5)
    rows = make_perm_rows(attrs, 12, 3)
    cneg, cpos = _lib.Enrichment(nb, attrs).perm_counts(rows, "z-score", "tc")
    dense = unpack_packed(g["nb_layout"], n).astype(np.int64)
    oneg, opos = orc.perm_counts_from_rows(dense, attrs, "z-score", rows)
    assert np.abs(cneg.astype(int) - oneg).sum() <= 6 and np.abs(cpos.astype(int) - opos).sum() <= 6


def test_zscore_small_attribute_counts_take_the_exact_engine(ctx, stage1_small):
    g = stage1_small
    n = g["x"].shape[0]
    nb = _lib.Neighborhoods(ctx, n).upload_packed(g["nb_layout"])
    attrs = _attrs(n, 5, "normal32", 9)
    rows = make_perm_rows(attrs, 9, 3)
    plan = _lib.Enrichment(nb, attrs)
    a = plan.perm_counts(rows, "z-score", "auto")
    s = plan.perm_counts(rows, "z-score", "simt")
    assert np.array_equal(a[0], s[0]) and np.array_equal(a[1], s[1])
    with pytest.raises(_lib.SafeB200Error):
        plan.perm_counts(rows, "z-score", "tc")


def test_zscore_full_size_c3_sampled(ctx):
    """configs[2] at full size with neighborhood_score_type='z-score': tensor path on 3 permutations against the exact
    engine on sampled columns (the SIMT engine on all 2000 columns would take minutes)."""
    cfg = syn.make_config("C3", shuffle=True)
    net, n, m, attrs = cfg["net"], cfg["n"], cfg["m"], cfg["attributes"]
    nr = cfg["radius"] * (np.max(net["x"]) - np.min(net["x"]))
    nb = _lib.Neighborhoods(ctx, n).shortpath(net["indptr"], net["indices"], net["csr_length"], nr)
    rows = make_perm_rows(attrs, 3, 7)
    plan = _lib.Enrichment(nb, attrs).set_node_order(kd_order(net["x"], net["y"]))
    tneg, tpos = plan.perm_counts(rows, "z-score", "tc")
    st = plan.stats()
    plan.close()
    cols = np.sort(np.random.default_rng(9).choice(m, 64, replace=False))
    # the permutation stream of the column subset must be the one of the full matrix: same rows, sampled columns
    sub = _lib.Enrichment(nb, np.ascontiguousarray(attrs[:, cols]))
    sneg, spos = sub.perm_counts(rows, "z-score", "simt")
    sub.close()
    assert np.array_equal(tneg[:, cols], sneg) and np.array_equal(tpos[:, cols], spos)
    assert st["fixups"] < 1e-3 * n * m * 3
    nb.close()
