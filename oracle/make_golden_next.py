"""Golden vectors for the rows SURVEY.md section 8(f) marks "next", from the UNMODIFIED reference.

    python oracle/make_golden_next.py       (this container only: needs /root/reference)

  graph_small.npz : safe_io.calculate_edge_lengths (safepy/safe_io.py:311-333) on a weighted 300-node graph
                    (weights incl. a few exact zeros and one self loop), plus the CSR the Dijkstra cost rule sees
  top_small.npz   : SAFE.define_top_attributes (safepy/safe.py:610-661) on seeded nes_binary columns
  domains_small.npz : SAFE.define_domains (safepy/safe.py:661-716) on the same kind of columns.  pandas 3 (installed
                    here; the reference pins 2.2.3) no longer accepts DataFrame.groupby(axis=1), which the reference
                    calls twice, so for the duration of that call DataFrame.groupby is wrapped to answer axis=1 as
                    the documented equivalent df.T.groupby(...).agg().T -- the reference code itself is unmodified.
"""
import os
import sys

import networkx as nx
import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from ref_import import import_reference  # noqa: E402
from safepy_b200 import synthetic as syn  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


class groupby_axis1_shim:
    """pandas >= 3 dropped DataFrame.groupby(axis=1); answer it with the transpose for the reference's two calls."""

    def __enter__(self):
        self.orig = orig = pd.DataFrame.groupby

        class _T:
            def __init__(self, gb):
                self.gb = gb

            def sum(self):
                return self.gb.sum().T

            def max(self):
                return self.gb.max().T

        def groupby(frame, *args, axis=0, **kwargs):
            if axis in (1, "columns"):
                return _T(orig(frame.T, *args, **kwargs))
            return orig(frame, *args, **kwargs)

        pd.DataFrame.groupby = groupby
        return self

    def __exit__(self, *exc):
        pd.DataFrame.groupby = self.orig
        return False


def main():
    ref = import_reference()
    from safepy import safe_io
    rng = np.random.default_rng(99)

    # ---- edge lengths
    net = syn.make_network(300, 1500, 41)
    eu, ev = net["edges"][:, 0], net["edges"][:, 1]
    w = np.round(rng.uniform(0.2, 3.0, len(eu)), 3)
    w[rng.choice(len(eu), 5, replace=False)] = 0.0      # zero weight: the reference leaves 'length' unset
    g = nx.Graph()
    for i in range(net["n"]):
        g.add_node(i, x=float(net["x"][i]), y=float(net["y"][i]))
    g.add_weighted_edges_from((int(a), int(b), float(c)) for a, b, c in zip(eu, ev, w))
    g.add_edge(7, 7, weight=1.5)                          # self loop
    safe_io.calculate_edge_lengths(g, verbose=False)
    eu2 = np.append(eu, 7)
    ev2 = np.append(ev, 7)
    w2 = np.append(w, 1.5)
    length = np.array([g[int(a)][int(b)].get("length", np.nan) for a, b in zip(eu2, ev2)])
    np.savez_compressed(os.path.join(OUT, "graph_small.npz"), x=net["x"], y=net["y"], eu=eu2.astype(np.int32),
                        ev=ev2.astype(np.int32), weight=w2, length=length)

    # ---- top attributes
    net = syn.make_network(400, 2400, 43)
    n, m = net["n"], 24
    sf = ref.SAFE(verbose=False)
    sf.graph = syn.to_networkx(net)
    sf.graph_euclidean = None if not hasattr(sf, "graph_euclidean") else sf.graph_euclidean
    nb = np.zeros((n, m))
    xs, ys = net["x"], net["y"]
    for j in range(m):
        k = rng.integers(1, 4)                            # 1-3 spatial blobs per attribute
        for _ in range(k):
            c = rng.integers(0, n)
            d = np.hypot(xs - xs[c], ys - ys[c])
            nb[d < rng.uniform(0.03, 0.15), j] = 1
        if j % 5 == 0:
            nb[rng.uniform(size=n) < 0.02, j] = 1         # scattered singletons
    nb[:, 3] = 0
    nb[:4, 3] = 1                                         # below the minimum size
    sf.nes_binary = nb
    sf.attributes = pd.DataFrame({"id": np.arange(m), "name": [str(j) for j in range(m)]})
    sf.attributes["num_neighborhoods_enriched"] = np.sum(nb, axis=0)
    sf.define_top_attributes()
    sizes = np.zeros((m, n), dtype=np.int64)
    for j in range(m):
        s = sf.attributes.at[j, "size_connected_components"]
        if s is not None:
            s = np.atleast_1d(np.asarray(s))
            sizes[j, :len(s)] = s
    np.savez_compressed(os.path.join(OUT, "top_small.npz"), x=net["x"], y=net["y"], edges=net["edges"],
                        length=net["length"], nes_binary=nb,
                        top=sf.attributes["top"].values.astype(bool),
                        num_cc=sf.attributes["num_connected_components"].values.astype(np.int64),
                        num_large_cc=sf.attributes["num_large_connected_components"].values.astype(np.int64),
                        cc_sizes=sizes, min_size=np.int64(sf.attribute_enrichment_min_size))
    # ---- domains
    from scipy.cluster.hierarchy import linkage
    from scipy.spatial.distance import pdist
    net = syn.make_network(500, 3000, 47)
    n, m = net["n"], 40
    xs, ys = net["x"], net["y"]
    nb = np.zeros((n, m))
    centres = rng.integers(0, n, 7)                       # attributes share 7 regions -> real clusters
    for j in range(m):
        c = centres[j % 7]
        d = np.hypot(xs - xs[c], ys - ys[c])
        nb[d < rng.uniform(0.05, 0.12), j] = 1
        nb[rng.uniform(size=n) < 0.01, j] = 1
    nb[:, 11] = 0                                         # an attribute without any enriched neighborhood
    nes = np.where(nb > 0, rng.uniform(1.4, 6.0, (n, m)), rng.uniform(-1.0, 1.2, (n, m)))
    top = np.ones(m, dtype=bool)
    top[[5, 11, 17]] = False
    sfd = ref.SAFE(verbose=False)
    sfd.nes, sfd.nes_binary = nes, nb
    sfd.attributes = pd.DataFrame({"id": np.arange(m), "name": [str(j) for j in range(m)], "top": top})
    with groupby_axis1_shim():
        sfd.define_domains()
    ids = np.array([c for c in sfd.node2domain.columns if c not in ("primary_domain", "primary_nes")], dtype=np.int64)
    mt = nb[:, top].T
    np.savez_compressed(os.path.join(OUT, "domains_small.npz"), nes=nes, nes_binary=nb, top=top,
                        threshold=np.float64(sfd.attribute_distance_threshold),
                        jaccard=pdist(mt, metric="jaccard"), linkage=linkage(mt, method="average", metric="jaccard"),
                        domain=sfd.attributes["domain"].values.astype(np.int64), domain_ids=ids,
                        node2domain=sfd.node2domain[list(ids)].values.astype(np.float64),
                        primary_domain=sfd.node2domain["primary_domain"].values.astype(np.int64),
                        primary_nes=sfd.node2domain["primary_nes"].values.astype(np.float64))
    print("wrote graph_small.npz, top_small.npz, domains_small.npz")


if __name__ == "__main__":
    main()
