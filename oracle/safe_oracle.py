"""CPU oracle for SAFE's neighborhood + enrichment path.  TEST INFRASTRUCTURE ONLY.

This module restates, in NumPy/SciPy/networkx, what the reference (baryshnikova-lab/safepy) computes on the path
`SAFE.define_neighborhoods` -> `SAFE.compute_pvalues`.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference arm may import it; the product (safepy_b200/) never does.

Parity status: PINNED.  The reference is pure Python, so it cannot be compiled into oracle/_ref; instead
oracle/make_golden.py imports the unmodified reference from /root/reference (plotting/FDR imports stubbed, see
oracle/ref_import.py), runs it on seeded synthetic inputs and commits inputs + outputs under tests/golden/.
tests/test_oracle_golden.py checks every function below against those files.  The reference's own tests
(tests/test_neighborhoods.py, tests/test_enrichments.py) need the external `safe-data` repository, which is not
available offline, so their known-answer values cannot be replayed here.

The arithmetic of the path lives in third-party packages the reference pins in extras/requirements.txt:
networkx==3.4.2 (Dijkstra), scipy==1.15.2 (pdist, hypergeom.sf -> Boost.Math), numpy==2.2.3 (np.dot, legacy
MT19937 np.random).  The same packages (newer builds) are installed in this image and are called directly where
the reference calls them.

All citations are file:line into the reference checkout.
"""
import numpy as np
import networkx as nx
from scipy.sparse import csr_matrix
from scipy.sparse.csgraph import dijkstra as _cs_dijkstra
from scipy.spatial.distance import pdist, squareform
from scipy.stats import hypergeom


# --------------------------------------------------------------------------------------------- stage 1
def neighborhood_radius(x, radius, metric):
    """safe.py:390-391, 404-405, 409: nr = radius * (max x - min x) for the layout metrics, raw radius for
    'shortpath'."""
    if metric == "shortpath":
        return radius
    x = list(x)
    return radius * (np.max(x) - np.min(x))


def neighborhoods_euclidean(x, y, nr):
    """safe.py:393-399: squareform(pdist(coords)) < nr, int matrix, diagonal = (0 < nr)."""
    coords = np.concatenate([np.reshape(np.asarray(x, dtype=float), (-1, 1)),
                             np.reshape(np.asarray(y, dtype=float), (-1, 1))], axis=1)
    d = squareform(pdist(coords, "euclidean"))
    out = np.zeros(d.shape, dtype=int)
    out[d < nr] = 1
    return out


def neighborhoods_euclidean_rows(x, y, nr, rows):
    """Row-sampled variant for sizes where the N x N matrix does not fit: same arithmetic as pdist
    (sqrt(dx*dx + dy*dy), no fused multiply-add), evaluated for the given source rows only."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    out = np.zeros((len(rows), x.shape[0]), dtype=np.uint8)
    for k, r in enumerate(rows):
        dx = x[r] - x
        dy = y[r] - y
        d = np.sqrt(dx * dx + dy * dy)
        out[k] = d < nr
    return out


def neighborhoods_shortpath_nx(graph, nr, weight):
    """safe.py:406-415 verbatim: networkx all-pairs Dijkstra with cutoff, one store per reached pair.
    weight='length' for shortpath_weighted_layout, 'weight' (networkx default) for shortpath."""
    n = graph.number_of_nodes()
    out = np.zeros([n, n], dtype=int)
    all_shortest_paths = dict(nx.all_pairs_dijkstra_path_length(graph, weight=weight, cutoff=nr))
    for s in all_shortest_paths:
        for t in all_shortest_paths[s].keys():
            out[(s, t)] = 1
    return out


def graph_to_csr(graph, weight):
    """Symmetric CSR of an nx.Graph whose nodes are 0..N-1, edge cost = data.get(weight, 1) exactly like
    networkx's _weight_function (networkx/algorithms/shortest_paths/weighted.py)."""
    n = graph.number_of_nodes()
    src, dst, w = [], [], []
    for u, v, data in graph.edges(data=True):
        c = data.get(weight, 1)
        src.append(u); dst.append(v); w.append(c)
        if u != v:
            src.append(v); dst.append(u); w.append(c)
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    w = np.asarray(w, dtype=np.float64)
    order = np.lexsort((dst, src))
    src, dst, w = src[order], dst[order], w[order]
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(indptr, src + 1, 1)
    indptr = np.cumsum(indptr)
    return indptr, dst.astype(np.int32), w


def neighborhoods_shortpath_csr(indptr, indices, length, nr, rows=None):
    """Faster restatement for big graphs: scipy.sparse.csgraph.dijkstra(limit=nr) gives the same fp64
    left-to-right path sums as networkx (checked bit-for-bit against neighborhoods_shortpath_nx in
    tests/test_oracle_golden.py); membership is dist <= nr (networkx: `if vu_dist > cutoff: continue`)."""
    n = len(indptr) - 1
    data = np.ones(len(indices)) if length is None else np.asarray(length, dtype=np.float64)
    if np.any(data == 0.0):
        raise ValueError("zero-length edges: use neighborhoods_shortpath_nx (csgraph may drop explicit zeros)")
    g = csr_matrix((data, indices, indptr), shape=(n, n))
    rows = np.arange(n) if rows is None else np.asarray(rows)
    out = np.zeros((len(rows), n), dtype=np.uint8)
    step = max(1, (64 << 20) // max(n, 1))
    for b in range(0, len(rows), step):
        idx = rows[b:b + step]
        d = _cs_dijkstra(g, directed=True, indices=idx, limit=nr)
        out[b:b + step] = d <= nr
        out[np.arange(b, b + len(idx)), idx] = 1  # the source is always a member (distance 0)
    return out


# --------------------------------------------------------------------------------------------- stage 2
def compute_neighborhood_score(neighborhood2node, node2attribute, neighborhood_score_type):
    """safe_extras.py:6-33, statement for statement."""
    with np.errstate(invalid="ignore", divide="ignore"):
        A = neighborhood2node
        B = np.where(~np.isnan(node2attribute), node2attribute, 0)
        NA = A
        NB = np.where(~np.isnan(node2attribute), 1, 0)
        AB = np.dot(A, B)
        neighborhood_score = AB
        if neighborhood_score_type == "z-score":
            N = np.dot(NA, NB)
            M = np.divide(AB, N)
            EXX = np.divide(np.dot(A, np.power(B, 2)), N)
            EEX = np.power(M, 2)
            std = np.sqrt(EXX - EEX)
            neighborhood_score = np.divide(M, std)
            neighborhood_score[std == 0] = np.nan
            neighborhood_score[N < 3] = np.nan
    return neighborhood_score


def run_permutations(neighborhood2node, node2attribute, neighborhood_score_type, num_permutations, random_seed):
    """safe_extras.py:36-70 without the progress bar: cumulative in-place row shuffles of the rows that hold
    data, legacy global NumPy RNG."""
    np.random.seed(random_seed)
    s0 = compute_neighborhood_score(neighborhood2node, node2attribute, neighborhood_score_type)
    n2a = np.copy(node2attribute)
    indx_vals = np.nonzero(np.sum(~np.isnan(n2a), axis=1))[0]
    counts_neg = np.zeros(s0.shape)
    counts_pos = np.zeros(s0.shape)
    for _ in np.arange(num_permutations):
        n2a[indx_vals, :] = n2a[np.random.permutation(indx_vals), :]
        sp = compute_neighborhood_score(neighborhood2node, n2a, neighborhood_score_type)
        with np.errstate(invalid="ignore", divide="ignore"):
            counts_neg = np.add(counts_neg, sp <= s0)
            counts_pos = np.add(counts_pos, sp >= s0)
    return counts_neg, counts_pos


def perm_gather_rows(node2attribute, num_permutations, random_seed):
    """The permutation stream of safe_extras.py:46-58 as explicit gather indices:
    rows[p, t] = original row of node2attribute that sits at node t after permutation p (cumulative)."""
    np.random.seed(random_seed)
    n = node2attribute.shape[0]
    indx_vals = np.nonzero(np.sum(~np.isnan(node2attribute), axis=1))[0]
    cur = np.arange(n)
    rows = np.empty((num_permutations, n), dtype=np.int32)
    for p in range(num_permutations):
        cur[indx_vals] = cur[np.random.permutation(indx_vals)]
        rows[p] = cur
    return rows


def perm_counts_from_rows(neighborhood2node, node2attribute, neighborhood_score_type, rows):
    """Counts for explicit gather indices (same comparisons as run_permutations, safe_extras.py:60-66)."""
    s0 = compute_neighborhood_score(neighborhood2node, node2attribute, neighborhood_score_type)
    counts_neg = np.zeros(s0.shape, dtype=np.int64)
    counts_pos = np.zeros(s0.shape, dtype=np.int64)
    for p in range(rows.shape[0]):
        sp = compute_neighborhood_score(neighborhood2node, node2attribute[rows[p], :], neighborhood_score_type)
        with np.errstate(invalid="ignore"):
            counts_neg += sp <= s0
            counts_pos += sp >= s0
    return counts_neg, counts_pos


def randomization_nes(ns, counts_neg, counts_pos, num_permutations, attribute_sign):
    """safe.py:528-554 without FDR: NaN mask, counts / P, -log10 with the 1/P floor, sign combination."""
    counts_neg = np.array(counts_neg, dtype=np.float64)
    counts_pos = np.array(counts_pos, dtype=np.float64)
    idx = np.isnan(ns)
    counts_neg[idx] = np.nan
    counts_pos[idx] = np.nan
    pvalues_neg = counts_neg / num_permutations
    pvalues_pos = counts_pos / num_permutations
    nes_pos = -np.log10(np.where(pvalues_pos == 0, 1 / num_permutations, pvalues_pos))
    nes_neg = -np.log10(np.where(pvalues_neg == 0, 1 / num_permutations, pvalues_neg))
    if attribute_sign == "highest":
        nes = nes_pos
    elif attribute_sign == "lowest":
        nes = nes_neg
    else:
        nes = nes_pos - nes_neg
    return pvalues_neg, pvalues_pos, nes


def hypergeom_pvalues(neighborhoods, node2attribute):
    """safe.py:573-608 without FDR: returns (pvalues_pos, nes)."""
    n_nodes, n_attr = node2attribute.shape
    nodes_not_nan = np.any(~np.isnan(node2attribute), axis=1)
    n = np.sum(nodes_not_nan)
    N = np.zeros([n_nodes, n_attr]) + n
    N_in_group = np.tile(np.nansum(node2attribute, axis=0), (n_nodes, 1))
    neighborhood_size = np.dot(neighborhoods, nodes_not_nan.astype(int))[:, np.newaxis]
    N_in_neighborhood = np.tile(neighborhood_size, (1, n_attr))
    N_in_neighborhood_in_group = np.dot(neighborhoods, np.where(~np.isnan(node2attribute), node2attribute, 0))
    pvalues_pos = hypergeom.sf(N_in_neighborhood_in_group - 1, N, N_in_group, N_in_neighborhood)
    with np.errstate(divide="ignore"):
        nes = -np.log10(pvalues_pos)
    return pvalues_pos, nes


def hypergeom_pvalues_block(neighborhood_rows, node2attribute, cols):
    """The cells [rows, cols] of hypergeom_pvalues for a block of source rows (neighborhood_rows = those rows of the
    neighborhood matrix, [R, N]) and attribute columns `cols`: the same statements (safe.py:573-608), with the
    node-level quantities (which nodes carry data, n) still taken from the WHOLE attribute matrix as upstream.
    Lets the full-size configurations be checked on a sample -- scipy's hypergeom.sf runs at ~6e4 cells/s."""
    nodes_not_nan = np.any(~np.isnan(node2attribute), axis=1)
    n = np.sum(nodes_not_nan)
    sub = node2attribute[:, cols]
    r, c = neighborhood_rows.shape[0], sub.shape[1]
    N = np.zeros([r, c]) + n
    N_in_group = np.tile(np.nansum(sub, axis=0), (r, 1))
    neighborhood_size = np.dot(neighborhood_rows, nodes_not_nan.astype(int))[:, np.newaxis]
    N_in_neighborhood = np.tile(neighborhood_size, (1, c))
    N_in_neighborhood_in_group = np.dot(neighborhood_rows, np.where(~np.isnan(sub), sub, 0))
    pvalues_pos = hypergeom.sf(N_in_neighborhood_in_group - 1, N, N_in_group, N_in_neighborhood)
    with np.errstate(divide="ignore"):
        nes = -np.log10(pvalues_pos)
    return pvalues_pos, nes


def nes_binary(nes, enrichment_threshold):
    """safe.py:468-470."""
    idx = ~np.isnan(nes)
    out = np.zeros(nes.shape)
    out[idx] = np.abs(nes[idx]) > -np.log10(enrichment_threshold)
    return out


# --------------------------------------------------------------------------------------------- sparse twin
def score_sum_csr(neighborhoods_u8, node2attribute, rows=None):
    """'sum' score with fp64 accumulation in ascending-neighbor order (the order the CUDA kernels use); equal to
    np.dot bit-for-bit whenever the partial sums are exact in fp64 (binary / integer / dyadic / float32 data of
    moderate range).  Used where the dense int64 matrix of the reference would not fit."""
    b = np.where(np.isnan(node2attribute), 0, node2attribute).astype(np.float64)
    if rows is not None:
        b = b[rows]
    a = csr_matrix(neighborhoods_u8.astype(np.float64))
    out = np.zeros((a.shape[0], b.shape[1]))
    for i in range(a.shape[0]):
        acc = np.zeros(b.shape[1])
        for t in a.indices[a.indptr[i]:a.indptr[i + 1]]:
            acc += b[t]
        out[i] = acc
    return out


# --------------------------------------------------------------------------------------------- next rows (8f)
def edge_lengths(x, y, eu, ev, weight=None):
    """safe_io.py:318-331 per edge: squareform(pdist(coords))[u, v] * adjacency[u, v]; zero adjacency entries become
    NaN there and get no 'length' attribute (returned as NaN here).  pdist's sqrt(dx*dx + dy*dy) is unfused, and so
    is NumPy's elementwise arithmetic."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    dx = x[eu] - x[ev]
    dy = y[eu] - y[ev]
    d = np.sqrt(dx * dx + dy * dy)
    if weight is None:
        return d
    w = np.asarray(weight, dtype=np.float64)
    out = d * w
    out[w == 0] = np.nan
    return out


def top_attributes(indptr, indices, nes_binary, min_size):
    """safe.py:626-658 with attribute_unimodality_metric='connectivity': (top, num_connected_components,
    num_large_connected_components, list of component-size arrays sorted descending).  Attributes below the minimum
    number of enriched neighborhoods are not examined (their counters stay 0, safe.py:635-638)."""
    from scipy.sparse.csgraph import connected_components
    nes_binary = np.asarray(nes_binary)
    n, m = nes_binary.shape
    g = csr_matrix((np.ones(len(indices), dtype=np.int8), indices, indptr), shape=(n, n))
    num_enriched = np.sum(nes_binary, axis=0)
    top = num_enriched >= min_size
    num_cc = np.zeros(m, dtype=np.int64)
    num_large = np.zeros(m, dtype=np.int64)
    sizes = [None] * m
    for j in np.nonzero(top)[0]:
        nodes = np.nonzero(nes_binary[:, j] > 0)[0]
        sub = g[nodes][:, nodes]
        k, lab = connected_components(sub, directed=False)
        s = np.sort(np.bincount(lab, minlength=k))[::-1]
        num_cc[j] = k
        num_large[j] = int(np.sum(s >= min_size))
        sizes[j] = s
    top = top & ~(num_cc > 1)
    return top, num_cc, num_large, sizes


# --------------------------------------------------------------------------------------------- FDR
def fdrcorrection(pvals):
    """statsmodels.stats.multitest.fdrcorrection(pvals, alpha=0.05, method='indep', is_sorted=False)[1], the call
    safe.py:538,541,604 makes through np.apply_along_axis.  PARITY UNPINNED: statsmodels==0.14.4 (pinned in the
    reference's extras/requirements.txt:6) is not installed in this image and cannot be (no network), so this is a
    restatement of its published algorithm, not a replay:
        sortind = argsort(p); ps = p[sortind]; ecdf = arange(1, n + 1) / float(n)
        raw = ps / ecdf; adj = minimum.accumulate(raw[::-1])[::-1]; adj[adj > 1] = 1; out[sortind] = adj
    NaNs sort last and np.minimum propagates them, so one NaN turns the whole vector into NaN.
    tests/test_oracle_golden.py cross-checks it against scipy.stats.false_discovery_control(method='bh'), an
    independent implementation in an installed library (equal to the last bit or one ulp: SciPy multiplies by m / k)."""
    pvals = np.asarray(pvals, dtype=np.float64)
    nobs = len(pvals)
    sortind = np.argsort(pvals)
    ps = np.take(pvals, sortind)
    ecdf = np.arange(1, nobs + 1) / float(nobs)
    with np.errstate(invalid="ignore"):
        raw = ps / ecdf
        adj = np.minimum.accumulate(raw[::-1])[::-1]
        adj[adj > 1] = 1
    out = np.empty_like(adj)
    out[sortind] = adj
    return out


def fdr_rows(pvalues):
    """np.apply_along_axis(fdrcorrection, 1, pvalues)[:, 1, :] of safe.py:538-542 / 604-605."""
    pvalues = np.asarray(pvalues, dtype=np.float64)
    return np.stack([fdrcorrection(r) for r in pvalues]) if len(pvalues) else pvalues.copy()


def randomization_tail(ns, counts_neg, counts_pos, num_permutations, attribute_sign, multiple_testing,
                       enrichment_threshold):
    """safe.py:528-554 followed by safe.py:466-472: (pvalues_neg, pvalues_pos, nes, nes_binary, num_enriched)."""
    counts_neg = np.array(counts_neg, dtype=np.float64)
    counts_pos = np.array(counts_pos, dtype=np.float64)
    idx = np.isnan(ns)
    counts_neg[idx] = np.nan
    counts_pos[idx] = np.nan
    pvalues_neg = counts_neg / num_permutations
    pvalues_pos = counts_pos / num_permutations
    if multiple_testing:
        pvalues_neg = fdr_rows(pvalues_neg)
        pvalues_pos = fdr_rows(pvalues_pos)
    nes_pos = -np.log10(np.where(pvalues_pos == 0, 1 / num_permutations, pvalues_pos))
    nes_neg = -np.log10(np.where(pvalues_neg == 0, 1 / num_permutations, pvalues_neg))
    nes = {"highest": nes_pos, "lowest": nes_neg}.get(attribute_sign, nes_pos - nes_neg)
    nb = nes_binary(nes, enrichment_threshold)
    return pvalues_neg, pvalues_pos, nes, nb, np.sum(nb, axis=0)


# --------------------------------------------------------------------------------------------- domains (8f rank 4)
def jaccard_condensed(nes_binary_matrix, columns):
    """The distances linkage(m, metric='jaccard') evaluates for m = nes_binary[:, top].T (safe.py:672-673):
    scipy.spatial.distance.pdist on the 0/1 float rows."""
    m = np.asarray(nes_binary_matrix, dtype=np.float64)[:, columns].T
    return pdist(m, metric="jaccard")


def define_domains(nes, nes_binary_matrix, top, attribute_distance_threshold, attribute_distance_metric="jaccard"):
    """safe.py:672-708: (domain per attribute, node2domain counts [n, n_domains + 1] with column d = domain id d,
    primary_domain, primary_nes).  The reference's two `groupby(level='domain', axis=1)` calls (sum / max over the
    attributes of a domain; pandas >= 3 no longer accepts axis=1) are written out with NumPy."""
    from scipy.cluster.hierarchy import fcluster, linkage
    nes = np.asarray(nes)
    nb = np.asarray(nes_binary_matrix)
    top = np.asarray(top, dtype=bool)
    Z = linkage(nb[:, top].T, method="average", metric=attribute_distance_metric)
    max_d = np.max(Z[:, 2] * attribute_distance_threshold)
    domains = fcluster(Z, max_d, criterion="distance")
    domain = np.zeros(nb.shape[1], dtype=np.int64)
    domain[top] = domains
    ids = np.unique(domain)                                    # groupby sorts its keys
    counts = np.stack([nb[:, domain == d].sum(axis=1) for d in ids], axis=1)
    import warnings
    with warnings.catch_warnings():                            # groupby(...).max() skips NaN; all-NaN stays NaN
        warnings.simplefilter("ignore", RuntimeWarning)
        maxnes = np.stack([np.nanmax(nes[:, domain == d], axis=1) for d in ids], axis=1)
    real = ids >= 1                                            # .loc[:, 1:]
    t_max = counts[:, real].max(axis=1)
    primary = ids[real][np.argmax(counts[:, real], axis=1)]    # idxmax: first maximum
    primary[t_max == 0] = 0
    col_of = {d: k for k, d in enumerate(ids)}
    has0 = 0 in col_of
    primary_nes = np.array([maxnes[i, col_of[d]] if (d != 0 or has0) else np.nan for i, d in enumerate(primary)])
    return domain, ids, counts, primary, primary_nes
