"""GPU: the SAFE class surface (define_neighborhoods / compute_pvalues) end to end against the reference's recorded
outputs, plus the two drop-in free functions."""
import pickle

import numpy as np
import pandas as pd
import pytest

import safepy_b200
from conftest import net_from_golden
from safepy_b200 import SAFE, synthetic as syn
from safepy_b200._lib import unpack_packed

pytestmark = pytest.mark.gpu


def make_sf(g, **settings):
    net = net_from_golden(g)
    sf = SAFE(verbose=False)
    sf.load_network(graph=syn.to_networkx(net))
    for k, v in settings.items():
        setattr(sf, k, v)
    return sf, net


def test_define_neighborhoods_three_metrics(stage1_small):
    g = stage1_small
    sf, net = make_sf(g)
    n = net["n"]
    sf.define_neighborhoods(node_distance_metric="shortpath_weighted_layout", neighborhood_radius=float(g["r_layout"]))
    assert np.array_equal(sf.neighborhoods.words, g["nb_layout"])
    # the reference tests read it like this (tests/test_neighborhoods.py:21-23)
    num_neighbors = np.sum(sf.neighborhoods, axis=1)
    assert np.array_equal(num_neighbors, unpack_packed(g["nb_layout"], n).sum(axis=1))
    sf.define_neighborhoods(node_distance_metric="shortpath", neighborhood_radius=int(g["r_hops"]))
    assert np.array_equal(sf.neighborhoods.words, g["nb_hops"])
    sf.define_neighborhoods(node_distance_metric="euclidean", neighborhood_radius=float(g["r_euclid"]))
    assert np.array_equal(sf.neighborhoods.words, g["nb_euclid"])
    assert sf.node_distance_metric == "euclidean"          # kwargs are sticky (safe.py:374-381)
    pickle.loads(pickle.dumps(sf.neighborhoods))            # SAFE.save must keep working


@pytest.mark.parametrize("kind", ["normal32", "dyadic", "single", "normal64"])
def test_compute_pvalues_randomization(stage2_small, kind):
    g = stage2_small
    sf, net = make_sf(g, random_seed=int(g["seed"]))
    sf.define_neighborhoods(neighborhood_radius=float(g["radius"]))
    sf.load_attributes(attribute_file=g["attr_" + kind].copy())
    sf.compute_pvalues(how="randomization", num_permutations=int(g["num_permutations"]), verbose=False)
    assert sf.num_permutations == int(g["num_permutations"])
    if kind == "normal64":
        assert np.allclose(sf.ns, g["ns_%s_sum" % kind], rtol=1e-13, atol=1e-13, equal_nan=True)
    else:
        assert np.array_equal(sf.ns, g["ns_%s_sum" % kind], equal_nan=True)
    assert np.array_equal(sf.pvalues_neg, g["rand_pneg_" + kind], equal_nan=True)
    assert np.array_equal(sf.pvalues_pos, g["rand_ppos_" + kind], equal_nan=True)
    assert np.array_equal(sf.nes, g["rand_nes_" + kind], equal_nan=True)      # host arithmetic on equal counts
    assert np.array_equal(sf.nes_binary, g["rand_nesbin_" + kind])
    assert np.array_equal(sf.attributes["num_neighborhoods_enriched"].values, g["rand_enriched_" + kind])


def test_compute_pvalues_auto_picks_hypergeometric(stage2_small):
    g = stage2_small
    sf, net = make_sf(g)
    sf.define_neighborhoods(neighborhood_radius=float(g["radius"]))
    sf.load_attributes(attribute_file=g["attr_binary"].copy())
    sf.compute_pvalues(verbose=False)
    ref = g["hyper_nes"]
    ok = np.isfinite(ref) & (np.abs(ref) >= 1e-3)
    assert np.array_equal(np.isnan(sf.nes), np.isnan(ref))
    assert np.all(np.abs(sf.nes[ok] - ref[ok]) <= 1e-6 * np.abs(ref[ok]))
    assert np.array_equal(sf.nes_binary, g["hyper_nesbin"])
    assert np.array_equal(sf.attributes["num_neighborhoods_enriched"].values, g["hyper_enriched"])
    # background='network' rewrites NaN to 0 in place first (safe.py:449-451)
    sf2, _ = make_sf(g, background="network")
    sf2.define_neighborhoods(neighborhood_radius=float(g["radius"]))
    sf2.load_attributes(attribute_file=g["attr_binary"].copy())
    sf2.compute_pvalues(verbose=False)
    assert not np.isnan(sf2.node2attribute).any()
    ref = g["hyper_bgnet_nes"]
    ok = np.isfinite(ref) & (np.abs(ref) >= 1e-3)
    assert np.all(np.abs(sf2.nes[ok] - ref[ok]) <= 1e-6 * np.abs(ref[ok]))


def test_attribute_sign_and_processes_rounding(stage2_small):
    g = stage2_small
    sf, net = make_sf(g, random_seed=3, attribute_sign="highest")
    sf.define_neighborhoods(neighborhood_radius=float(g["radius"]))
    sf.load_attributes(attribute_file=g["attr_normal32"].copy())
    sf.compute_pvalues(how="randomization", num_permutations=25, processes=4, verbose=False)
    assert sf.num_permutations == 28                       # rounded up to a multiple of `processes` (safe.py:503-504)
    assert np.array_equal(sf.nes, -np.log10(np.where(sf.pvalues_pos == 0, 1 / 28, sf.pvalues_pos)), equal_nan=True)


def test_host_outputs_limits_what_comes_back(stage2_small):
    """sf.host_outputs: only the named [N, M] results are copied to the host, the others are None; the per-attribute
    sums (computed on the device in the same pass) are unaffected."""
    g = stage2_small
    sf, net = make_sf(g, random_seed=int(g["seed"]))
    sf.define_neighborhoods(neighborhood_radius=float(g["radius"]))
    sf.load_attributes(attribute_file=g["attr_normal32"].copy())
    sf.host_outputs = ("nes",)
    sf.compute_pvalues(how="randomization", num_permutations=int(g["num_permutations"]), verbose=False)
    assert np.array_equal(sf.nes, g["rand_nes_normal32"], equal_nan=True)
    assert sf.ns is None and sf.pvalues_neg is None and sf.pvalues_pos is None and sf.nes_binary is None
    assert np.array_equal(sf.attributes["num_neighborhoods_enriched"].values, g["rand_enriched_normal32"])
    sf.host_outputs = ("nes", "bogus")
    with pytest.raises(ValueError):
        sf.compute_pvalues(how="randomization", num_permutations=5, verbose=False)


def test_free_function_drop_ins(stage2_small):
    g = stage2_small
    n = g["x"].shape[0]
    dense = unpack_packed(g["neighborhoods"], n).astype(np.int64)     # what reference callers hold
    attrs = g["attr_normal32"]
    s = safepy_b200.compute_neighborhood_score(dense, attrs, "sum")
    assert np.array_equal(s, g["ns_normal32_sum"])
    cneg, cpos = safepy_b200.run_permutations((dense, attrs, "sum", int(g["num_permutations"]), int(g["seed"])),
                                              verbose=False)
    assert cneg.dtype == np.float64
    assert np.array_equal(cneg, g["cneg_normal32_sum"]) and np.array_equal(cpos, g["cpos_normal32_sum"])


def test_dataframe_attributes_align_to_node_labels(stage2_small):
    g = stage2_small
    sf, net = make_sf(g, random_seed=int(g["seed"]))
    sf.define_neighborhoods(neighborhood_radius=float(g["radius"]))
    n = net["n"]
    labels = ["n%d" % i for i in range(n)]
    shuffled = np.random.default_rng(0).permutation(n)
    frame = pd.DataFrame(g["attr_normal32"][shuffled], index=[labels[i] for i in shuffled])
    sf.load_attributes(attribute_file=frame)
    assert np.array_equal(sf.node2attribute, g["attr_normal32"].astype(np.float64), equal_nan=True) or \
        np.array_equal(sf.node2attribute, g["attr_normal32"], equal_nan=True)
    sf.compute_pvalues(how="randomization", num_permutations=int(g["num_permutations"]), verbose=False)
    assert np.array_equal(sf.nes, g["rand_nes_normal32"], equal_nan=True)
