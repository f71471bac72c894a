"""Generate tests/golden/*.npz by running the UNMODIFIED reference on seeded synthetic inputs.

    python oracle/make_golden.py            (this container only: needs /root/reference)

Each file holds the inputs and what the reference's own code returned for them:
  SAFE.define_neighborhoods (safepy/safe.py:369-430) for the three metrics,
  safe_extras.run_permutations / compute_neighborhood_score (safepy/safe_extras.py:6-70),
  SAFE.compute_pvalues by randomization and by the hypergeometric test (safepy/safe.py:432-608).
The oracle (oracle/safe_oracle.py) and the CUDA path are both tested against these files.
"""
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from ref_import import import_reference  # noqa: E402
from safepy_b200 import synthetic as syn  # noqa: E402
from safepy_b200._lib import pack_dense  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def reference_safe(ref, net, attrs=None, **settings):
    sf = ref.SAFE(verbose=False)
    sf.graph = syn.to_networkx(net)
    for k, v in settings.items():
        setattr(sf, k, v)
    if attrs is not None:
        sf.node2attribute = attrs.copy()
        sf.attributes = pd.DataFrame({"id": np.arange(attrs.shape[1]),
                                      "name": [str(j) for j in range(attrs.shape[1])]})
    return sf


def stage1_case(ref, n, edges, seed, r_layout, r_hops, r_euclid):
    net = syn.make_network(n, edges, seed)
    out = dict(x=net["x"], y=net["y"], edges=net["edges"], length=net["length"],
               r_layout=r_layout, r_hops=r_hops, r_euclid=r_euclid)
    sf = reference_safe(ref, net)
    sf.define_neighborhoods(node_distance_metric="shortpath_weighted_layout", neighborhood_radius=r_layout)
    out["nb_layout"] = pack_dense(sf.neighborhoods)
    sf.define_neighborhoods(node_distance_metric="shortpath", neighborhood_radius=r_hops)
    out["nb_hops"] = pack_dense(sf.neighborhoods)
    sf.define_neighborhoods(node_distance_metric="euclidean", neighborhood_radius=r_euclid)
    out["nb_euclid"] = pack_dense(sf.neighborhoods)
    # 'shortpath' with an explicit integer-ish 'weight' attribute on the edges (networkx default weight key)
    g = syn.to_networkx(net, with_length=False)
    rng = np.random.default_rng(seed + 99)
    wts = rng.integers(1, 4, len(net["edges"])).astype(float)
    for (u, v), w in zip(net["edges"], wts):
        g[int(u)][int(v)]["weight"] = float(w)
    sf.graph = g
    sf.define_neighborhoods(node_distance_metric="shortpath", neighborhood_radius=3)
    out["edge_weight"] = wts
    out["nb_weighted_hops"] = pack_dense(sf.neighborhoods)
    return net, out


def main():
    ref = import_reference()
    os.makedirs(OUT, exist_ok=True)

    # ---------------------------------------------------------------- stage 1 at two sizes
    net, small = stage1_case(ref, 400, 2800, 11, r_layout=0.15, r_hops=2, r_euclid=0.10)
    _, mid = stage1_case(ref, 1500, 10500, 12, r_layout=0.10, r_hops=2, r_euclid=0.06)
    np.savez_compressed(os.path.join(OUT, "stage1_small.npz"), **small)
    np.savez_compressed(os.path.join(OUT, "stage1_mid.npz"), **mid)

    # ---------------------------------------------------------------- stage 2 on the small network
    n = net["n"]
    sf = reference_safe(ref, net)
    sf.define_neighborhoods(node_distance_metric="shortpath_weighted_layout", neighborhood_radius=0.15)
    nb = sf.neighborhoods
    out = dict(x=net["x"], y=net["y"], edges=net["edges"], length=net["length"], radius=0.15,
               neighborhoods=pack_dense(nb), seed=7, num_permutations=60)
    kinds = {
        "normal32": syn.make_attributes(n, 6, 21, "normal32", nan_row_frac=0.1, nan_cell_frac=0.03),
        "dyadic": syn.make_attributes(n, 5, 22, "dyadic", nan_row_frac=0.1, nan_cell_frac=0.03),
        "binary": syn.make_attributes(n, 30, 23, "binary", nan_row_frac=0.08, nan_cell_frac=0.0),
        "normal64": np.random.default_rng(24).standard_normal((n, 4)),
        "single": syn.make_attributes(n, 1, 25, "normal32", nan_row_frac=0.33, nan_cell_frac=0.0),
    }
    kinds["binary"][:, 3] = np.where(np.isnan(kinds["binary"][:, 3]), np.nan, 0.0)  # an attribute nobody has
    for name, attrs in kinds.items():
        out["attr_" + name] = attrs
        for stype in ("sum", "z-score"):
            tag = "%s_%s" % (name, "sum" if stype == "sum" else "z")
            out["ns_" + tag] = ref.compute_neighborhood_score(nb, attrs, stype)
            cneg, cpos = ref.run_permutations((nb, attrs, stype, 60, 7), verbose=False)
            out["cneg_" + tag] = cneg
            out["cpos_" + tag] = cpos
        # the full method, randomization branch
        sfr = reference_safe(ref, net, attrs, random_seed=7)
        sfr.neighborhoods = nb
        sfr.compute_pvalues(how="randomization", num_permutations=60, verbose=False)
        out["rand_pneg_" + name] = sfr.pvalues_neg
        out["rand_ppos_" + name] = sfr.pvalues_pos
        out["rand_nes_" + name] = sfr.nes
        out["rand_nesbin_" + name] = sfr.nes_binary
        out["rand_enriched_" + name] = sfr.attributes["num_neighborhoods_enriched"].values
    # hypergeometric branch (auto-selected for binary data, safe.py:461-466)
    sfh = reference_safe(ref, net, kinds["binary"])
    sfh.neighborhoods = nb
    sfh.compute_pvalues(verbose=False)
    out["hyper_p"] = sfh.pvalues_pos
    out["hyper_nes"] = sfh.nes
    out["hyper_nesbin"] = sfh.nes_binary
    out["hyper_enriched"] = sfh.attributes["num_neighborhoods_enriched"].values
    # background='network' turns NaN into 0 before the test (safe.py:449-451)
    sfb = reference_safe(ref, net, kinds["binary"], background="network")
    sfb.neighborhoods = nb
    sfb.compute_pvalues(verbose=False)
    out["hyper_bgnet_p"] = sfb.pvalues_pos
    out["hyper_bgnet_nes"] = sfb.nes
    np.savez_compressed(os.path.join(OUT, "stage2_small.npz"), **out)

    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
