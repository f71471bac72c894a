"""Golden vectors for the rows SURVEY.md section 8(f) marks "next", from the UNMODIFIED reference.

    python oracle/make_golden_next.py       (this container only: needs /root/reference)

  graph_small.npz : safe_io.calculate_edge_lengths (safepy/safe_io.py:311-333) on a weighted 300-node graph
                    (weights incl. a few exact zeros and one self loop), plus the CSR the Dijkstra cost rule sees
  top_small.npz   : SAFE.define_top_attributes (safepy/safe.py:610-661) on seeded nes_binary columns
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
    print("wrote graph_small.npz, top_small.npz")


if __name__ == "__main__":
    main()
