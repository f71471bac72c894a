"""Seeded synthetic networks and attribute matrices of the shapes BASELINE.json names.

Coordinates come from a Gaussian mixture in the unit disk; edges are k-nearest-neighbour links with a heavy-tailed
k plus a few random long edges; `length` is sqrt(dx*dx + dy*dy) evaluated unfused, i.e. the value the reference's
calculate_edge_lengths (safe_io.py:311-333) produces for unit adjacency weights.  Nodes are emitted in Morton
(Z-curve) order of their coordinates so that spatially compact neighborhoods occupy few tiles of the packed matrix.
"""
import numpy as np
from scipy.spatial import cKDTree


def morton_order(x, y, bits=16):
    """Indices that sort points along a Z-curve of the bounding box."""
    def spread(v):
        v = v.astype(np.uint64)
        v = (v | (v << 16)) & np.uint64(0x0000FFFF0000FFFF)
        v = (v | (v << 8)) & np.uint64(0x00FF00FF00FF00FF)
        v = (v | (v << 4)) & np.uint64(0x0F0F0F0F0F0F0F0F)
        v = (v | (v << 2)) & np.uint64(0x3333333333333333)
        v = (v | (v << 1)) & np.uint64(0x5555555555555555)
        return v

    def quant(v):
        lo, hi = float(np.min(v)), float(np.max(v))
        span = (hi - lo) or 1.0
        return np.minimum(((v - lo) / span * (1 << bits)).astype(np.int64), (1 << bits) - 1)

    code = spread(quant(np.asarray(x))) | (spread(quant(np.asarray(y))) << np.uint64(1))
    return np.argsort(code, kind="stable")


def make_points(n, seed, clusters=None, sigma=0.04):
    rng = np.random.default_rng(seed)
    clusters = clusters or max(4, n // 200)
    r = np.sqrt(rng.uniform(0, 1, clusters))
    th = rng.uniform(0, 2 * np.pi, clusters)
    cx, cy = r * np.cos(th), r * np.sin(th)
    which = rng.integers(0, clusters, n)
    x = cx[which] + rng.normal(0, sigma, n)
    y = cy[which] + rng.normal(0, sigma, n)
    order = morton_order(x, y)
    return np.ascontiguousarray(x[order]), np.ascontiguousarray(y[order])


def make_network(n, target_edges, seed, sigma=0.04, long_edge_frac=0.02):
    """Returns dict(x, y, edges[E,2] (u < v, unique), length[E], indptr, indices, csr_length)."""
    x, y = make_points(n, seed, sigma=sigma)
    rng = np.random.default_rng(seed + 1)
    # heavy-tailed out-degree: lognormal, scaled so that the union of kNN links hits the edge target
    mean_k = max(1.0, target_edges * (1 - long_edge_frac) / n * 1.25)
    k = np.clip(np.rint(rng.lognormal(np.log(mean_k) - 0.5, 1.0, n)), 1, min(n - 1, 256)).astype(int)
    tree = cKDTree(np.column_stack([x, y]))
    kmax = int(k.max())
    _, nbr = tree.query(np.column_stack([x, y]), k=kmax + 1)
    src = np.repeat(np.arange(n), k)
    mask = np.arange(1, kmax + 1)[None, :] <= k[:, None]
    pick = nbr[:, 1:][mask]
    n_long = int(target_edges * long_edge_frac)
    ls = rng.integers(0, n, n_long)
    lt = rng.integers(0, n, n_long)
    u = np.concatenate([src, ls])
    v = np.concatenate([pick, lt])
    keep = u != v
    u, v = u[keep], v[keep]
    lo, hi = np.minimum(u, v), np.maximum(u, v)
    key = np.unique(lo.astype(np.int64) * n + hi)
    if len(key) > target_edges:
        key = np.sort(rng.choice(key, target_edges, replace=False))
    eu, ev = (key // n).astype(np.int64), (key % n).astype(np.int64)
    dx, dy = x[eu] - x[ev], y[eu] - y[ev]
    length = np.sqrt(dx * dx + dy * dy)
    indptr, indices, csr_len = edges_to_csr(n, eu, ev, length)
    return dict(n=n, x=x, y=y, edges=np.column_stack([eu, ev]), length=length, indptr=indptr, indices=indices,
                csr_length=csr_len)


def edges_to_csr(n, eu, ev, w=None):
    """Symmetric CSR (both directions stored, columns ascending inside a row) of an undirected edge list."""
    eu = np.asarray(eu, dtype=np.int64)
    ev = np.asarray(ev, dtype=np.int64)
    loop = eu == ev
    src = np.concatenate([eu, ev[~loop]])
    dst = np.concatenate([ev, eu[~loop]])
    order = np.lexsort((dst, src))
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(indptr, src + 1, 1)
    indptr = np.cumsum(indptr)
    ww = None
    if w is not None:
        w = np.asarray(w, dtype=np.float64)
        ww = np.concatenate([w, w[~loop]])[order]
    return indptr, dst[order].astype(np.int32), ww


def to_networkx(net, with_length=True):
    """nx.Graph with node attrs key/x/y/label and edge attr 'length' -- what the reference loaders produce."""
    import networkx as nx
    g = nx.Graph()
    for i in range(net["n"]):
        g.add_node(i, key=i, x=float(net["x"][i]), y=float(net["y"][i]), label="n%d" % i, label_orf="n%d" % i)
    if with_length:
        g.add_edges_from((int(u), int(v), {"length": float(w)}) for (u, v), w in zip(net["edges"], net["length"]))
    else:
        g.add_edges_from((int(u), int(v)) for u, v in net["edges"])
    return g


def make_attributes(n, m, seed, kind="normal32", nan_row_frac=0.05, nan_cell_frac=0.01):
    """Attribute matrices:
    normal32 : N(0,1) rounded to float32 (what the reference's .txt loader yields, safe_io.py:361)
    dyadic   : round(N(0,1) * 1024) / 1024 stored as float32 (exact ties are exact in any summation order)
    binary   : GO-style 0/1 annotations, per-attribute density log-uniform in 0.1 % .. 5 %
    NaN rows / cells mark "no data"."""
    rng = np.random.default_rng(seed)
    if kind == "binary":
        dens = np.exp(rng.uniform(np.log(1e-3), np.log(5e-2), m))
        b = (rng.uniform(0, 1, (n, m)) < dens[None, :]).astype(np.float32)
    else:
        b = rng.standard_normal((n, m)).astype(np.float32)
        if kind == "dyadic":
            b = (np.round(b.astype(np.float64) * 1024) / 1024).astype(np.float32)
    if nan_row_frac > 0:
        rows = rng.uniform(0, 1, n) < nan_row_frac
        b[rows, :] = np.nan
    if nan_cell_frac > 0 and kind != "binary":
        cells = rng.uniform(0, 1, (n, m)) < nan_cell_frac
        b[cells] = np.nan
    return b


CONFIGS = {
    # name: (n, edges, m, attribute kind, metric, radius, num_permutations)
    "C1": dict(n=3971, edges=28202, m=1, kind="normal32", metric="shortpath_weighted_layout", radius=0.10,
               perms=1000, nan_row_frac=0.33, nan_cell_frac=0.0, seed=20241),
    "C2": dict(n=6000, edges=45000, m=4373, kind="binary", metric="shortpath_weighted_layout", radius=0.10,
               perms=0, nan_row_frac=0.046, nan_cell_frac=0.0, seed=20242),
    "C3": dict(n=20000, edges=150000, m=2000, kind="normal32", metric="shortpath_weighted_layout", radius=0.10,
               perms=1000, nan_row_frac=0.05, nan_cell_frac=0.01, seed=20243),
    "C4": dict(n=100000, edges=0, m=500, kind="normal32", metric="euclidean", radius=0.06, perms=1000,
               nan_row_frac=0.05, nan_cell_frac=0.01, seed=20244),
    "C5": dict(n=100000, edges=750000, m=5000, kind="normal32", metric="shortpath_weighted_layout", radius=0.05,
               perms=1000, nan_row_frac=0.05, nan_cell_frac=0.01, seed=20245),
}


def relabel_nodes(net, attrs, seed):
    """Random renumbering of the nodes (the generators emit them along a Morton curve; real inputs come in file order).
    Returns (net, attrs) with coordinates, edges, CSR and attribute rows permuted consistently."""
    n = net["n"]
    new_of_old = np.random.default_rng(seed).permutation(n)
    old_of_new = np.argsort(new_of_old)
    x, y = net["x"][old_of_new], net["y"][old_of_new]
    edges, length = net["edges"], net["length"]
    if len(edges):
        eu, ev = new_of_old[edges[:, 0]], new_of_old[edges[:, 1]]
        lo, hi = np.minimum(eu, ev), np.maximum(eu, ev)
        order = np.lexsort((hi, lo))
        eu, ev, length = lo[order], hi[order], length[order]
        indptr, indices, csr_len = edges_to_csr(n, eu, ev, length)
        edges = np.column_stack([eu, ev])
    else:
        indptr, indices, csr_len = net["indptr"], net["indices"], net["csr_length"]
    out = dict(n=n, x=np.ascontiguousarray(x), y=np.ascontiguousarray(y), edges=edges, length=length, indptr=indptr,
               indices=indices, csr_length=csr_len)
    return out, np.ascontiguousarray(attrs[old_of_new])


def make_config(name, scale=1.0, shuffle=False):
    """Inputs of a named configuration; `scale` < 1 shrinks n, edges and m proportionally (tests); `shuffle`
    renumbers the nodes randomly so that the input order carries no spatial locality."""
    c = dict(CONFIGS[name])
    n = max(64, int(c["n"] * scale))
    m = max(1, int(c["m"] * scale)) if c["m"] > 1 else 1
    if c["edges"]:
        net = make_network(n, int(c["edges"] * scale), c["seed"])
    else:
        x, y = make_points(n, c["seed"])
        net = dict(n=n, x=x, y=y, edges=np.zeros((0, 2), dtype=np.int64), length=np.zeros(0),
                   indptr=np.zeros(n + 1, dtype=np.int64), indices=np.zeros(0, dtype=np.int32),
                   csr_length=np.zeros(0))
    attrs = make_attributes(n, m, c["seed"] + 7, c["kind"], c["nan_row_frac"], c["nan_cell_frac"])
    if shuffle:
        net, attrs = relabel_nodes(net, attrs, c["seed"] + 13)
    c.update(n=n, m=m, net=net, attributes=attrs)
    return c
