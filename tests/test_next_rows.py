"""SURVEY.md section 8(f) rows built so far: edge lengths / CSR extraction and the connectivity test of
define_top_attributes.  CPU part: the oracle restatements against goldens recorded from the unmodified reference
(oracle/make_golden_next.py); GPU part: the kernels through the C ABI against the same goldens."""
import os

import numpy as np
import pandas as pd
import pytest

import safe_oracle as orc
from safepy_b200 import synthetic as syn

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


def test_oracle_edge_lengths_match_reference():
    g = load("graph_small.npz")
    got = orc.edge_lengths(g["x"], g["y"], g["eu"], g["ev"], g["weight"])
    assert np.array_equal(np.isnan(got), np.isnan(g["length"]))
    ok = ~np.isnan(got)
    assert np.array_equal(got[ok], g["length"][ok])      # bit-exact
    assert np.isnan(g["length"]).sum() == 5 and g["length"][-1] == 0.0   # zero weights, self loop


def test_oracle_top_attributes_match_reference():
    g = load("top_small.npz")
    n = g["x"].shape[0]
    indptr, indices, _ = syn.edges_to_csr(n, g["edges"][:, 0], g["edges"][:, 1])
    top, ncc, nlarge, sizes = orc.top_attributes(indptr, indices, g["nes_binary"], int(g["min_size"]))
    assert np.array_equal(top, g["top"])
    assert np.array_equal(ncc, g["num_cc"]) and np.array_equal(nlarge, g["num_large_cc"])
    for j, s in enumerate(sizes):
        ref = g["cc_sizes"][j]
        ref = ref[ref > 0]
        assert (s is None and len(ref) == 0) or np.array_equal(s, ref)


@pytest.mark.gpu
def test_gpu_edge_lengths_bit_exact(ctx):
    from safepy_b200 import _lib
    g = load("graph_small.npz")
    got = _lib.edge_lengths(ctx, g["x"], g["y"], g["eu"], g["ev"], g["weight"])
    assert np.array_equal(np.isnan(got), np.isnan(g["length"]))
    ok = ~np.isnan(got)
    assert np.array_equal(got[ok], g["length"][ok])
    net = syn.make_config("C1", shuffle=True)["net"]
    e = net["edges"]
    assert np.array_equal(_lib.edge_lengths(ctx, net["x"], net["y"], e[:, 0], e[:, 1]), net["length"])


@pytest.mark.gpu
def test_gpu_csr_build_matches_host(ctx):
    from safepy_b200 import _lib
    net = syn.make_config("C1", shuffle=True)["net"]
    e = net["edges"]
    indptr, indices, val = _lib.build_csr(ctx, net["n"], e[:, 0], e[:, 1], net["length"])
    assert np.array_equal(indptr, net["indptr"]) and np.array_equal(indices, net["indices"])
    assert np.array_equal(val, net["csr_length"])
    # self loop stored once, isolated nodes, no values
    indptr, indices, val = _lib.build_csr(ctx, 5, [0, 2, 2], [1, 2, 0])
    assert indptr.tolist() == [0, 2, 3, 5, 5, 5] and indices.tolist() == [1, 2, 0, 0, 2] and val is None


@pytest.mark.gpu
def test_gpu_components_match_reference(ctx):
    from safepy_b200 import _lib
    g = load("top_small.npz")
    n, m = g["nes_binary"].shape
    indptr, indices, _ = syn.edges_to_csr(n, g["edges"][:, 0], g["edges"][:, 1])
    min_size = int(g["min_size"])
    cand = np.nonzero(g["nes_binary"].sum(axis=0) >= min_size)[0]
    ncc, nlarge, labels = _lib.components(ctx, indptr, indices, g["nes_binary"], cand, min_size, want_labels=True)
    assert np.array_equal(ncc, g["num_cc"][cand]) and np.array_equal(nlarge, g["num_large_cc"][cand])
    for k, j in enumerate(cand):
        lab = labels[k]
        assert np.array_equal(lab >= 0, g["nes_binary"][:, j] > 0)
        s = np.sort(np.bincount(lab[lab >= 0]))[::-1]
        ref = g["cc_sizes"][j]
        assert np.array_equal(s[s > 0], ref[ref > 0])


@pytest.mark.gpu
def test_gpu_define_top_attributes_api(ctx):
    from safepy_b200 import SAFE
    g = load("top_small.npz")
    n, m = g["nes_binary"].shape
    sf = SAFE(verbose=False)
    sf.load_network(edges=g["edges"], x=g["x"], y=g["y"])
    assert np.array_equal(np.array([d["length"] for _, _, d in sf.graph.edges(data=True)]),
                          orc.edge_lengths(g["x"], g["y"], g["edges"][:, 0], g["edges"][:, 1]))
    sf.nes_binary = g["nes_binary"]
    sf.attributes = pd.DataFrame({"id": np.arange(m), "name": [str(j) for j in range(m)]})
    sf.attributes["num_neighborhoods_enriched"] = g["nes_binary"].sum(axis=0)
    sf.define_top_attributes()
    assert np.array_equal(sf.attributes["top"].values.astype(bool), g["top"])
    assert np.array_equal(sf.attributes["num_connected_components"].values, g["num_cc"])
    assert np.array_equal(sf.attributes["num_large_connected_components"].values, g["num_large_cc"])


@pytest.mark.gpu
def test_gpu_network_from_arrays_uses_the_device_csr(ctx, stage1_small):
    """load_network(edges, x, y): edge lengths and the CSR come from the device; same neighborhoods as the reference
    computed from its networkx graph."""
    from safepy_b200 import SAFE
    from safepy_b200.safe import graph_csr
    g = stage1_small
    sf = SAFE(verbose=False)
    sf.load_network(edges=g["edges"], x=g["x"], y=g["y"])
    walked = graph_csr(sf.graph, "length")                       # host walk over the graph object
    sf.graph.graph.pop("_safe_b200_csr")
    built = graph_csr(sf.graph, "length", ctx)                   # sb_graph_csr on the stored arrays
    for a, b in zip(walked, built):
        assert np.array_equal(a, b)
    sf.define_neighborhoods(neighborhood_radius=float(g["r_layout"]))
    assert np.array_equal(sf.neighborhoods.words, g["nb_layout"])
