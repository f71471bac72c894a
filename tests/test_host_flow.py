"""CPU: the host orchestration of the SAFE methods (dispatcher, sticky kwargs, plan sharing, tail hand-off, domains,
top attributes) with the device calls replaced by oracle-backed stand-ins, against the reference's recorded outputs.
The stand-ins exist only here: the product has no such path (tests/test_gpu_*.py run the same flows on the device)."""
import numpy as np
import pandas as pd
import pytest

import safe_oracle as orc
from conftest import load_golden, net_from_golden
from safepy_b200 import SAFE, PackedNeighborhoods, _lib, synthetic as syn
from safepy_b200 import safe as safe_mod


class OraclePlan:
    """What safe.py needs from _lib.Enrichment, computed by the oracle."""
    created = 0

    def __init__(self, dense, attrs):
        OraclePlan.created += 1
        self.dense, self.attrs = dense, np.asarray(attrs)
        self.n, self.m = self.attrs.shape
        self.closed = False
        self.calls = []

    def attr_summary(self):
        mask = np.isnan(self.attrs)
        return mask.sum(axis=0), int(np.sum(~mask & ~np.isin(self.attrs, [0, 1])))

    def null_begin(self, score_type, engine):
        self.score_type = score_type
        self.cneg = np.zeros((self.n, self.m))
        self.cpos = np.zeros((self.n, self.m))
        self.calls.append("begin")

    def null_add(self, rows):
        cneg, cpos = orc.perm_counts_from_rows(self.dense, self.attrs, self.score_type, rows)
        self.cneg += cneg
        self.cpos += cpos

    def null_add_stream(self, stream, num_perm, world=1, rank=0):
        self.calls.append(("stream", num_perm, world, rank))
        self.null_add(stream.next(num_perm))

    def null_finalize(self, num_permutations, attribute_sign, enrichment_threshold, multiple_testing, want=None):
        ns = orc.compute_neighborhood_score(self.dense, self.attrs, self.score_type)
        pn, pp, nes, nb, enriched = orc.randomization_tail(ns, self.cneg, self.cpos, num_permutations, attribute_sign,
                                                           multiple_testing, enrichment_threshold)
        return {"ns": ns, "pvalues_neg": pn, "pvalues_pos": pp, "nes": nes, "nes_binary": nb,
                "num_neighborhoods_enriched": enriched}

    def hypergeom_finalize(self, enrichment_threshold, multiple_testing):
        self.calls.append("hypergeom")
        p, nes = orc.hypergeom_pvalues(self.dense, self.attrs)
        if multiple_testing:
            p = orc.fdr_rows(p)
            with np.errstate(divide="ignore"):
                nes = -np.log10(p)
        nb = orc.nes_binary(nes, enrichment_threshold)
        return {"pvalues_pos": p, "nes": nes, "nes_binary": nb, "num_neighborhoods_enriched": nb.sum(axis=0)}

    def stats(self):
        return {}

    def close(self):
        self.closed = True


@pytest.fixture
def sf_small(stage2_small, monkeypatch):
    g = stage2_small
    net = net_from_golden(g)
    n = net["n"]
    sf = SAFE(verbose=False)
    sf.load_network(graph=syn.to_networkx(net))
    sf.neighborhoods = PackedNeighborhoods(g["neighborhoods"], n)
    dense = _lib.unpack_packed(g["neighborhoods"], n).astype(np.int64)
    plans = []

    def fake_plan(self):
        plans.append(OraclePlan(dense, self.node2attribute))
        return plans[-1]

    monkeypatch.setattr(safe_mod.SafeB200Mixin, "_enrichment_plan", fake_plan)
    return sf, g, plans


@pytest.mark.parametrize("kind", ["normal32", "single"])
def test_compute_pvalues_flow_randomization(sf_small, kind):
    sf, g, plans = sf_small
    sf.random_seed = int(g["seed"])
    sf.load_attributes(attribute_file=g["attr_" + kind].copy())
    np.random.seed(4242)
    sf.compute_pvalues(how="randomization", num_permutations=int(g["num_permutations"]), verbose=False)
    assert len(plans) == 1 and plans[0].closed                        # one plan serves summary, null and tail
    assert plans[0].calls == ["begin", ("stream", int(g["num_permutations"]), 1, 0)]
    assert sf.enrichment_type == "randomization" and sf.num_permutations == int(g["num_permutations"])   # sticky
    assert np.array_equal(sf.ns, g["ns_%s_sum" % kind], equal_nan=True)
    assert np.array_equal(sf.pvalues_neg, g["rand_pneg_" + kind], equal_nan=True)
    assert np.array_equal(sf.pvalues_pos, g["rand_ppos_" + kind], equal_nan=True)
    assert np.array_equal(sf.nes, g["rand_nes_" + kind], equal_nan=True)
    assert np.array_equal(sf.nes_binary, g["rand_nesbin_" + kind])
    assert np.array_equal(sf.attributes["num_neighborhoods_enriched"].values, g["rand_enriched_" + kind])
    # NumPy's global generator ends where the reference's run_permutations leaves it
    orc.perm_gather_rows(g["attr_" + kind], int(g["num_permutations"]), int(g["seed"]))
    expected_state = np.random.get_state()[1].copy()
    sf.compute_pvalues(verbose=False)
    assert np.array_equal(np.random.get_state()[1], expected_state)
    assert sf._plan is None and sf._tail is None


def test_compute_pvalues_flow_auto_and_background(sf_small):
    sf, g, plans = sf_small
    sf.load_attributes(attribute_file=g["attr_binary"].copy())
    sf.compute_pvalues(verbose=False)                                   # 'auto' + binary data -> hypergeometric
    assert plans[-1].calls == ["hypergeom"]
    assert np.array_equal(sf.pvalues_pos, g["hyper_p"], equal_nan=True)
    assert np.array_equal(sf.nes, g["hyper_nes"], equal_nan=True)
    assert np.array_equal(sf.nes_binary, g["hyper_nesbin"])
    assert np.array_equal(sf.attributes["num_neighborhoods_enriched"].values, g["hyper_enriched"])
    sf.compute_pvalues(background="network", verbose=False)            # NaN -> 0 in place first (safe.py:449-451)
    assert sf.background == "network" and not np.isnan(sf.node2attribute).any()
    assert np.array_equal(sf.nes, g["hyper_bgnet_nes"], equal_nan=True)
    with pytest.raises(ValueError):
        sf.compute_pvalues(background="nowhere")
    assert sf.background == "attribute_file"                           # validate_config restored the default
    # a branch method called on its own makes (and closes) its own plan
    before = len(plans)
    sf.compute_pvalues_by_hypergeom(verbose=False)
    assert len(plans) == before + 1 and plans[-1].closed and sf._plan is None


def test_non_integer_seed_takes_numpys_seeding_path(sf_small):
    sf, g, plans = sf_small
    seed = [3, 1, 4]                                                    # array seeds go through np.random.seed itself
    sf.random_seed = seed
    sf.load_attributes(attribute_file=g["attr_single"].copy())
    sf.compute_pvalues(how="randomization", num_permutations=12, verbose=False)
    assert plans[-1].calls == ["begin"]                                 # python row stream, not the native one
    dense = plans[-1].dense
    cneg, cpos = orc.run_permutations(dense, g["attr_single"], "sum", 12, seed)
    assert np.array_equal(sf.pvalues_neg, cneg / 12) and np.array_equal(sf.pvalues_pos, cpos / 12)


def test_define_domains_flow(monkeypatch):
    g = load_golden("domains_small.npz")
    monkeypatch.setattr(safe_mod, "get_context", lambda device=-1: None)
    monkeypatch.setattr(_lib, "jaccard", lambda ctx, member, cols: orc.jaccard_condensed(member, cols))
    m = g["nes"].shape[1]
    sf = SAFE(verbose=False)
    sf.nes, sf.nes_binary = g["nes"], g["nes_binary"]
    sf.attributes = pd.DataFrame({"id": np.arange(m), "name": [str(j) for j in range(m)], "top": g["top"]})
    sf.define_domains()
    ids = list(g["domain_ids"])
    assert np.array_equal(sf.attributes["domain"].values, g["domain"])
    assert list(sf.node2domain.columns) == ids + ["primary_domain", "primary_nes"]
    assert np.array_equal(sf.node2domain[ids].values, g["node2domain"])
    assert np.array_equal(sf.node2domain["primary_domain"].values, g["primary_domain"])
    assert np.array_equal(sf.node2domain["primary_nes"].values, g["primary_nes"])
    # any other metric is SciPy's, as upstream
    sf.attribute_distance_metric = "hamming"
    sf.define_domains()
    ref_domain = orc.define_domains(g["nes"], g["nes_binary"], g["top"], sf.attribute_distance_threshold, "hamming")[0]
    assert np.array_equal(sf.attributes["domain"].values, ref_domain)


def test_define_domains_primary_nes_skips_nan(monkeypatch):
    """safe.py:687-701 takes the per-domain maximum with pandas' groupby(...).max(), which skips NaN (z-score scores
    and invalid hypergeometric cells are NaN); an all-NaN group stays NaN.  Checked against pandas itself."""
    g = load_golden("domains_small.npz")
    monkeypatch.setattr(safe_mod, "get_context", lambda device=-1: None)
    monkeypatch.setattr(_lib, "jaccard", lambda ctx, member, cols: orc.jaccard_condensed(member, cols))
    nes = g["nes"].copy()
    rng = np.random.default_rng(3)
    nes[rng.uniform(size=nes.shape) < 0.3] = np.nan
    nes[5, :] = np.nan                                                  # a node without any finite NES
    m = nes.shape[1]
    sf = SAFE(verbose=False)
    sf.nes, sf.nes_binary = nes, g["nes_binary"]
    sf.attributes = pd.DataFrame({"id": np.arange(m), "name": [str(j) for j in range(m)], "top": g["top"]})
    sf.define_domains()
    domain = sf.attributes["domain"].values
    ref = pd.DataFrame(nes.T, index=pd.Index(domain, name="domain")).groupby(level="domain").max().T
    primary = sf.node2domain["primary_domain"].values
    want = np.array([ref.loc[i, d] for i, d in enumerate(primary)])
    assert np.array_equal(sf.node2domain["primary_nes"].values, want, equal_nan=True)
    assert np.isnan(sf.node2domain["primary_nes"].values[5])
    assert np.isfinite(want).sum() > 0.9 * len(want)
    _, _, _, _, oracle_nes = orc.define_domains(nes, g["nes_binary"], g["top"], sf.attribute_distance_threshold)
    assert np.array_equal(oracle_nes, want, equal_nan=True)


def test_define_top_attributes_flow(monkeypatch):
    g = load_golden("top_small.npz")
    n, m = g["nes_binary"].shape
    net = net_from_golden(g)

    def fake_components(ctx, indptr, indices, member, cand, min_size, want_labels=False):
        from scipy.sparse import csr_matrix
        from scipy.sparse.csgraph import connected_components
        adj = csr_matrix((np.ones(len(indices), dtype=np.int8), indices, indptr), shape=(n, n))
        labels = -np.ones((len(cand), n), dtype=np.int32)
        ncc, nlarge = np.zeros(len(cand), dtype=np.int32), np.zeros(len(cand), dtype=np.int32)
        for k, j in enumerate(cand):
            nodes = np.nonzero(member[:, j])[0]
            c, lab = connected_components(adj[nodes][:, nodes], directed=False)
            first = np.array([nodes[lab == q].min() for q in range(c)])
            labels[k, nodes] = first[lab]
            ncc[k] = c
            nlarge[k] = int(np.sum(np.bincount(lab) >= min_size))
        return ncc, nlarge, labels

    monkeypatch.setattr(safe_mod, "get_context", lambda device=-1: None)
    monkeypatch.setattr(_lib, "components", fake_components)
    sf = SAFE(verbose=False)
    sf.load_network(graph=syn.to_networkx(net))
    sf.nes_binary = g["nes_binary"]
    sf.attributes = pd.DataFrame({"id": np.arange(m), "name": [str(j) for j in range(m)]})
    sf.attributes["num_neighborhoods_enriched"] = g["nes_binary"].sum(axis=0)
    sf.define_top_attributes()
    assert np.array_equal(sf.attributes["top"].values.astype(bool), g["top"])
    assert np.array_equal(sf.attributes["num_connected_components"].values, g["num_cc"])
    assert np.array_equal(sf.attributes["num_large_connected_components"].values, g["num_large_cc"])
    for j in np.nonzero(g["num_cc"])[0]:
        sizes = np.atleast_1d(np.asarray(sf.attributes.at[j, "size_connected_components"]))
        assert np.array_equal(sizes, g["cc_sizes"][j][:len(sizes)]) and not g["cc_sizes"][j][len(sizes):].any()
