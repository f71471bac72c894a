"""SAFE class surface for the neighborhood + enrichment path, backed by libsafe_b200 (sm_100a CUDA).

Mirrors the reference's method names, keyword arguments, sticky-override behaviour, validation errors, log lines
and output attributes for
    SAFE.define_neighborhoods                 safepy/safe.py:369-430
    SAFE.compute_pvalues                      safepy/safe.py:432-472
    SAFE.compute_pvalues_by_randomization     safepy/safe.py:474-554
    SAFE.compute_pvalues_by_hypergeom         safepy/safe.py:556-608
Everything else of the reference class (file loaders, layouts, plotting, domains) is outside this package: use
`accelerate(safepy.safe.SAFE)` to graft these four methods onto the reference class, or the standalone `SAFE`
below, which takes graphs / arrays / DataFrames directly.

There is no CPU fallback: the methods raise `SafeB200Error` when the CUDA library or a B200 is missing.
"""
import contextlib
import logging
import threading
import time
import warnings

import numpy as np

from . import _lib
from .neighborhood_matrix import PackedNeighborhoods, as_packed
from .ordering import kd_order
from .distributed import DeviceArray, active_group, broadcast_row_shards, row_shard
from .permutations import iter_perm_rows, native_seed, perm_stream

DEFAULTS = {
    # safepy/safe_default.ini:1-24 and safepy/safe.py:57-107
    "background": "attribute_file",
    "node_distance_metric": "shortpath_weighted_layout",
    "neighborhood_radius_type": "diameter",
    "neighborhood_radius": 0.1,
    "attribute_sign": "both",
    "num_permutations": 1000,
    "multiple_testing": False,
    "neighborhood_score_type": "sum",
    "enrichment_type": "auto",
    "enrichment_threshold": 0.05,
    "enrichment_max_log10": 16,
    "attribute_enrichment_min_size": 10,
    "attribute_unimodality_metric": "connectivity",
    "attribute_distance_metric": "jaccard",
    "attribute_distance_threshold": 0.75,
    "random_seed": None,
}

_CONTEXTS = {}
_CONTEXTS_LOCK = threading.Lock()


def get_context(device=-1):
    """Process-wide libsafe_b200 context per CUDA device (created on first use).  device < 0 means the calling
    thread's CURRENT device, resolved to its ordinal on every call (so a later torch.cuda.set_device is honoured).
    A context serves one caller thread at a time (include/safe_b200.h); the SAFE methods hold `ctx.lock` while they
    run, so SAFE objects on different Python threads take turns instead of sharing scratch buffers."""
    if device is None or device < 0:
        device = _lib.current_device()
        if device < 0:
            device = -1                     # no usable device: let sb_ctx_create produce the error message
    with _CONTEXTS_LOCK:
        ctx = _CONTEXTS.get(device)
        if ctx is None or ctx.h is None:
            ctx = _lib.Context(device)
            _CONTEXTS[ctx.device] = ctx
    return ctx


def _locked(method):
    """Run a SAFE method under its device context's lock (re-entrant: compute_pvalues calls the branch methods)."""
    import functools

    @functools.wraps(method)
    def wrapper(self, *args, **kwargs):
        try:
            lock = getattr(get_context(self.device), "lock", None)
        except _lib.SafeB200Error:
            lock = None     # no device: the method validates its arguments first, then fails where it needs the GPU
        if lock is None:
            return method(self, *args, **kwargs)
        with lock:
            return method(self, *args, **kwargs)
    return wrapper


# ------------------------------------------------------------------------------------------------ graph -> arrays
def _node_coordinates(graph):
    nodes = list(graph)
    n = len(nodes)
    if nodes != list(range(n)):
        # the reference indexes the matrix with node ids (safe.py:412-415); every loader yields 0..N-1
        raise ValueError("graph nodes must be the integers 0..N-1 in order (as produced by the SAFE loaders)")
    x = np.fromiter((d for _, d in graph.nodes.data("x")), dtype=np.float64, count=n)
    y = np.fromiter((d for _, d in graph.nodes.data("y")), dtype=np.float64, count=n)
    return x, y


def graph_csr(graph, weight):
    """Symmetric CSR of the graph with the cost rule of networkx's Dijkstra: data.get(weight, 1)
    (what safe.py:406-410 hands to all_pairs_dijkstra_path_length).

    Rebuilt from the graph object on EVERY call, like the reference, which re-reads the edge data every time
    define_neighborhoods runs (safe.py:406-407): safe_io.apply_network_layout / calculate_edge_lengths (and users)
    edit 'length' / 'weight' in place, and networkx keeps no modification counter that a cache could be checked
    against.  The walk goes straight over the adjacency dicts with C-level iterators (the rows of the adjacency ARE
    the CSR rows, so nothing is sorted): ~0.1 s for 145k edges."""
    import operator
    from itertools import chain
    n = graph.number_of_nodes()
    adj = graph._adj if hasattr(graph, "_adj") else dict(graph.adjacency())
    deg = np.fromiter(map(len, adj.values()), dtype=np.int64, count=n)
    nnz = int(deg.sum())
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(deg, out=indptr[1:])
    indices = np.fromiter(chain.from_iterable(adj.values()), dtype=np.int32, count=nnz)
    if weight is None:  # structure only
        return indptr, indices, np.ones(nnz, dtype=np.float64)
    data = list(chain.from_iterable(map(dict.values, adj.values())))
    try:  # every edge carries the attribute (the common case): plain item lookups
        cost = np.fromiter(map(operator.itemgetter(weight), data), dtype=np.float64, count=nnz)
    except KeyError:
        cost = np.fromiter(map(operator.methodcaller("get", weight, 1), data), dtype=np.float64, count=nnz)
    return indptr, indices, cost


class SafeB200Mixin:
    """The four hot-path methods; the host class supplies graph / node2attribute / attributes / settings."""

    device = -1
    multi_gpu = True  # follow torch.distributed when the caller has initialised it (one process per GPU)
    # Opt-in promise by the caller: the graph built by load_network(edges=..., x=..., y=...) has not been edited
    # since.  define_neighborhoods then builds its CSR on the device from the arrays load_network kept (sb_graph_csr)
    # instead of re-reading every edge of the graph object (0.4 s for 720k edges).  Off by default: the reference
    # re-reads the edge data on every call, and in-place edits of 'length' cannot be detected from outside.
    assume_graph_unchanged = False
    results_rank = None  # with several ranks: None = every rank receives the [N, M] result arrays, r = only rank r
    # Which [N, M] fp64 results of the randomization test come back to the host (the others are set to None).  All
    # five by default, as upstream; at N = 100 000 x M = 5000 they are 4 GB each and their way into pageable memory
    # is the largest part of the call, so a caller who needs, say, only the NES can ask for ("nes", "nes_binary").
    host_outputs = ("ns", "pvalues_neg", "pvalues_pos", "nes", "nes_binary")
    _plan = None  # enrichment plan of the compute_pvalues call in progress
    _ahead = None  # (stream, num_permutations, world, rank): RNG read-ahead started before the plan was uploaded
    _tail = None  # (nes_binary, num_neighborhoods_enriched) handed from the enrichment branch to compute_pvalues

    # ---------------------------------------------------------------------------------- stage 1
    @_locked
    def define_neighborhoods(self, **kwargs):
        # sticky keyword overrides, safe.py:374-381
        for k in ("node_distance_metric", "neighborhood_radius_type", "neighborhood_radius"):
            if k in kwargs:
                setattr(self, k, kwargs[k])
        self.validate_config()

        ctx = get_context(self.device)
        x, y = _node_coordinates(self.graph)
        n = x.shape[0]
        dev = _lib.Neighborhoods(ctx, n)
        # one process per GPU (torch.distributed initialised by the caller): every rank searches its block of
        # sources and the packed rows are exchanged once
        dist = active_group(self.multi_gpu)
        r0, r1 = row_shard(n, dist.get_world_size(), dist.get_rank()) if dist else (0, n)
        metric = self.node_distance_metric
        if metric == "euclidean":
            nr = self.neighborhood_radius * (np.max(x) - np.min(x))          # safe.py:390-391 (x extent only)
            dev.euclid(x, y, nr, r0, r1)
        else:
            if metric == "shortpath_weighted_layout":
                nr = self.neighborhood_radius * (np.max(x) - np.min(x))      # safe.py:404-405
                kept = getattr(self, "_loaded_edge_arrays", None)
                if self.assume_graph_unchanged and kept is not None and kept[0] is self.graph:
                    indptr, indices, cost = _lib.build_csr(ctx, n, kept[1][:, 0], kept[1][:, 1], kept[2])
                else:
                    indptr, indices, cost = graph_csr(self.graph, "length")
            else:
                nr = self.neighborhood_radius                                # safe.py:409
                indptr, indices, cost = graph_csr(self.graph, "weight")
            dev.shortpath(indptr, indices, cost, nr, r0, r1)
            # safe.py:417 stores the dict-of-dicts of distances; nothing in safepy reads it
            self.node_distances = None
        if dist:
            import torch
            ctx.synchronize()
            words = torch.as_tensor(DeviceArray(dev.words_dev, n * dev.ld), device=torch.device("cuda", ctx.device))
            broadcast_row_shards(dist, words.view(n, dev.ld), n)
            torch.cuda.synchronize(ctx.device)

        packed = PackedNeighborhoods(None, n, device=dev)    # host copy of the words is fetched on first use
        # locality hint for stage 2 (nodes sorted spatially); results do not depend on it
        packed.node_order = kd_order(x, y)
        num_neighbors = packed.row_sums()
        if self.verbose:
            logging.info("Node distance metric: %s" % self.node_distance_metric)
            logging.info("Neighborhood definition: %.2f x %s" % (self.neighborhood_radius,
                                                                 self.neighborhood_radius_type))
            logging.info("Number of nodes per neighborhood (mean +/- std): %.2f +/- %.2f"
                         % (np.mean(num_neighbors), np.std(num_neighbors)))
        self.neighborhoods = packed

    # ---------------------------------------------------------------------------------- stage 2
    @_locked
    def compute_pvalues(self, **kwargs):
        if "how" in kwargs:
            self.enrichment_type = kwargs["how"]
        for k in ("neighborhood_score_type", "multiple_testing", "background"):
            if k in kwargs:
                setattr(self, k, kwargs[k])
        self.validate_config()

        if self.background == "network":
            logging.info("Setting all null attribute values to 0. Using the network as background for enrichment.")
            self.node2attribute[np.isnan(self.node2attribute)] = 0

        # The attribute matrix goes to the device once; the look at its values (NaN share per attribute, anything
        # other than 0 / 1 / NaN -- safe.py:453-458) happens there, and the same plan serves the chosen test.
        self._tail = None
        if self.enrichment_type not in ("hypergeometric", "auto"):
            # randomization whatever the attribute values are: the permutation draws do not need the device, so
            # they start now and overlap with the upload of the attribute matrix (2 GB at N = 100 000 x M = 5000)
            self._start_read_ahead(kwargs)
        try:
            self._compute_pvalues_on_plan(kwargs)
        finally:
            if self._ahead is not None:        # an error before the null consumed it
                self._ahead[0].close()
                self._ahead = None

        if self._tail is not None:
            # nes_binary and the per-attribute sums came out of the same kernel pass as the NES (safe.py:466-472)
            self.nes_binary, enriched = self._tail
            self._tail = None
        else:
            idx = ~np.isnan(self.nes)
            self.nes_binary = np.zeros(self.nes.shape)
            self.nes_binary[idx] = np.abs(self.nes[idx]) > -np.log10(self.enrichment_threshold)
            enriched = np.sum(self.nes_binary, axis=0)
        self.attributes["num_neighborhoods_enriched"] = enriched

    def _randomization_permutations(self, kwargs):
        """num_permutations as compute_pvalues_by_randomization will use it (kwarg + safe.py:503-504 rounding)."""
        if "num_permutations" in kwargs:
            self.num_permutations = kwargs["num_permutations"]
        num_processes = kwargs.get("processes", 1)
        self.validate_config()
        if num_processes > 1:
            # safe.py:503-504 rounds the permutation count up to a multiple of the worker count
            per = int(np.ceil(self.num_permutations / num_processes))
            self.num_permutations = per * num_processes

    def _start_read_ahead(self, kwargs):
        self._ahead = None
        self._randomization_permutations(kwargs)
        dist = active_group(self.multi_gpu)
        if not native_seed(self.random_seed) or (dist and self.random_seed is None):
            return                              # NumPy's own seeding path / the error is raised by the branch method
        world, rank = (dist.get_world_size(), dist.get_rank()) if dist else (1, 0)
        stream = perm_stream(self.node2attribute, self.random_seed)
        stream.prefetch(self.num_permutations, world, rank)
        self._ahead = (stream, self.num_permutations, world, rank)

    def _take_stream(self, world, rank):
        """The read-ahead stream of this call when it matches, else a fresh one."""
        ahead, self._ahead = self._ahead, None
        if ahead is not None:
            if ahead[1:] == (self.num_permutations, world, rank):
                return ahead[0]
            ahead[0].close()
        return perm_stream(self.node2attribute, self.random_seed)

    def _compute_pvalues_on_plan(self, kwargs):
        with self._plan_scope() as plan:
            nans, num_other_values = plan.attr_summary()
            if np.any(nans / plan.n > 0.5):
                logging.warning("WARNING: more than 50% of nodes in the network are set to NaN and will be ignored "
                                "for calculating enrichment.\n'Consider setting sf.background = ''network''.'")
            if self.enrichment_type == "hypergeometric" or (self.enrichment_type == "auto" and num_other_values == 0):
                self.compute_pvalues_by_hypergeom(**kwargs)
            else:
                self.compute_pvalues_by_randomization(**kwargs)

    @contextlib.contextmanager
    def _plan_scope(self):
        """The enrichment plan of the enclosing compute_pvalues call, or a fresh one when a branch method is called
        on its own (both are public upstream)."""
        if self._plan is not None:
            yield self._plan
            return
        plan = self._enrichment_plan()
        self._plan = plan
        try:
            yield plan
        finally:
            self._plan = None
            plan.close()

    def _enrichment_plan(self):
        ctx = get_context(self.device)
        packed = as_packed(self.neighborhoods)
        b = np.asarray(self.node2attribute)
        if b.shape[0] != packed.n:
            raise ValueError("node2attribute has %d rows but the network has %d nodes" % (b.shape[0], packed.n))
        plan = _lib.Enrichment(packed.on_device(ctx), b)
        order = getattr(packed, "node_order", None)
        if order is not None:
            plan.set_node_order(order)
        return plan

    @_locked
    def compute_pvalues_by_randomization(self, **kwargs):
        if kwargs:
            logging.warning("Current settings (possibly overwriting global ones):")
            for k in kwargs:
                logging.warning("\t%s=%s" % (k, str(kwargs[k])))
        logging.info("Using randomization to calculate enrichment...")
        # (the reference sleeps 1 s here to keep its progress bar tidy, safe.py:484; not reproduced)

        self._randomization_permutations(kwargs)

        # The host replays the reference's RNG stream piece by piece (background thread) while the device counts
        # the previous piece; counts -> p-values -> (FDR) -> NES -> nes_binary happen on the device in one tail
        # pass (safe.py:526-554, 466-472), so the count arrays never visit the host.
        t0 = time.perf_counter()
        with self._plan_scope() as plan:
            t1 = time.perf_counter()
            plan.null_begin(self.neighborhood_score_type, getattr(self, "engine", "auto"))
            dist = active_group(self.multi_gpu)
            if dist and not native_seed(self.random_seed):
                raise ValueError("permutation shards over several GPUs need random_seed to be an int in [0, 2**32); "
                                 "got %r" % (self.random_seed,))
            if dist and self.random_seed is None:
                raise ValueError("permutation shards over several GPUs need a random_seed: with None every rank "
                                 "would draw its own stream")
            if dist:
                # permutations sharded over the ranks; ONE sum all-reduce of the device counts (safe.py:518-519 sums
                # its worker results the same way) -- the packed word per cell (pos << 16 | neg) while P < 65536,
                # half the bytes of the two arrays
                import torch
                device = torch.device("cuda", plan.ctx.device)
                stream = self._take_stream(dist.get_world_size(), dist.get_rank())
                plan.null_add_stream(stream, self.num_permutations, dist.get_world_size(), dist.get_rank())
                stream.sync_numpy()
                stream.close()
                plan.ctx.synchronize()
                if self.num_permutations < 65536:
                    pk, _ = plan.null_packed_dev()
                    counts = torch.as_tensor(DeviceArray(pk, plan.n * plan.m), device=device)
                else:
                    neg, _ = plan.null_counts_dev()
                    plan.ctx.synchronize()
                    counts = torch.as_tensor(DeviceArray(neg, 2 * plan.n * plan.m), device=device)
                dist.all_reduce(counts)
                # the observed scores the tail needs (NaN mask, self.ns): every rank computes its block of node rows,
                # the blocks are exchanged in place (the exact fp64 kernel is the expensive part of the tail)
                r0, r1 = row_shard(plan.n, dist.get_world_size(), dist.get_rank())
                ns_ptr = plan.observed_rows_dev(self.neighborhood_score_type, r0, r1)
                plan.ctx.synchronize()
                ns_t = torch.as_tensor(DeviceArray(ns_ptr, plan.n * plan.m, "<f8"), device=device).view(plan.n, plan.m)
                broadcast_row_shards(dist, ns_t, plan.n)
                torch.cuda.synchronize(plan.ctx.device)
                plan.observed_set_ready(self.neighborhood_score_type)
                plan.null_set_perms(self.num_permutations)
            elif native_seed(self.random_seed):
                # one C call: a producer thread replays the RNG stream piece by piece into pinned memory while the
                # device counts the previous piece
                stream = self._take_stream(1, 0)
                plan.null_add_stream(stream, self.num_permutations)
                stream.sync_numpy()          # the global generator ends where upstream's would
                stream.close()
            else:
                with contextlib.closing(iter_perm_rows(self.node2attribute, self.num_permutations,
                                                       self.random_seed)) as perm_rows:
                    for rows in perm_rows:
                        plan.null_add(rows)
            t2 = time.perf_counter()
            if self.multiple_testing:
                logging.info("Running FDR-adjustment of p-values...")
            known = ("ns", "pvalues_neg", "pvalues_pos", "nes", "nes_binary")
            unknown = [k for k in self.host_outputs if k not in known]
            if unknown:
                raise ValueError("host_outputs: unknown result %r (choose from %s)" % (unknown[0], ", ".join(known)))
            want = tuple(k for k in known if k in self.host_outputs)
            if dist and self.results_rank is not None and dist.get_rank() != self.results_rank:
                # this rank keeps only the per-attribute sums, which the results rank sends: no tail here at all
                import torch
                enriched = torch.empty(plan.m, dtype=torch.float64, device=torch.device("cuda", plan.ctx.device))
                dist.broadcast(enriched, src=self.results_rank)
                out = {"num_neighborhoods_enriched": enriched.cpu().numpy()}
            else:
                out = plan.null_finalize(self.num_permutations, self.attribute_sign, self.enrichment_threshold,
                                         self.multiple_testing, want=want)
                if dist and self.results_rank is not None:
                    import torch
                    enriched = torch.from_numpy(out["num_neighborhoods_enriched"]).to(
                        torch.device("cuda", plan.ctx.device))
                    dist.broadcast(enriched, src=self.results_rank)
            self.last_enrichment_stats = plan.stats()
        # host wall clock of the three phases (upload + CSR view, streamed null, fused tail + result copies)
        self.last_enrichment_seconds = {"plan": t1 - t0, "null": t2 - t1, "tail": time.perf_counter() - t2}
        self.ns = out.get("ns")
        self.pvalues_neg = out.get("pvalues_neg")
        self.pvalues_pos = out.get("pvalues_pos")
        self.nes = out.get("nes")
        self._tail = (out.get("nes_binary"), out["num_neighborhoods_enriched"])

    @_locked
    def compute_pvalues_by_hypergeom(self, **kwargs):
        if kwargs:
            if "verbose" in kwargs:
                self.verbose = kwargs["verbose"]
            if self.verbose:
                logging.warning("Overwriting global settings:")
                for k in kwargs:
                    logging.warning("\t%s=%s" % (k, str(kwargs[k])))
        self.validate_config()
        if self.verbose:
            logging.info("Using the hypergeometric test to calculate enrichment...")
        if self.multiple_testing and self.verbose:
            logging.info("Running FDR-adjustment of p-values...")
        with self._plan_scope() as plan:
            out = plan.hypergeom_finalize(self.enrichment_threshold, self.multiple_testing)
        self.pvalues_pos = out["pvalues_pos"]
        self.nes = out["nes"]
        self._tail = (out["nes_binary"], out["num_neighborhoods_enriched"])

    # ---------------------------------------------------------------------------------- next call of the workflow
    @_locked
    def define_top_attributes(self, **kwargs):
        """safe.py:610-661.  The connectivity test (connected components of the subgraph induced by the enriched
        nodes, one attribute after the other through networkx upstream) runs batched over attributes on the GPU."""
        for k in ("attribute_unimodality_metric", "attribute_enrichment_min_size"):
            if k in kwargs:
                setattr(self, k, kwargs[k])
        self.validate_config()
        logging.info("Criteria for top attributes:")
        logging.info("- minimum number of enriched neighborhoods: %d" % self.attribute_enrichment_min_size)
        logging.info("- region-specific distribution of enriched neighborhoods as defined by: %s"
                     % self.attribute_unimodality_metric)
        self.attributes["top"] = False
        self.attributes.loc[
            self.attributes["num_neighborhoods_enriched"] >= self.attribute_enrichment_min_size, "top"] = True
        if self.attribute_unimodality_metric == "connectivity":
            self.attributes["num_connected_components"] = 0
            self.attributes["size_connected_components"] = None
            self.attributes["size_connected_components"] = self.attributes["size_connected_components"].astype(object)
            self.attributes["num_large_connected_components"] = 0
            cand = self.attributes.index.values[self.attributes["top"].values.astype(bool)]
            if len(cand):
                # safe.py:644-645: an edgeless network falls back to the Euclidean-distance graph
                graph = self.graph_euclidean if getattr(self, "graph_euclidean", None) else self.graph
                indptr, indices, _ = graph_csr(graph, None)
                ncc, nlarge, labels = _lib.components(get_context(self.device), indptr, indices, self.nes_binary > 0,
                                                      cand, self.attribute_enrichment_min_size, want_labels=True)
                for k, attribute in enumerate(cand):
                    lab = labels[k]
                    sizes = np.sort(np.bincount(lab[lab >= 0]))[::-1]
                    sizes = sizes[sizes > 0]
                    self.attributes.loc[attribute, "num_connected_components"] = int(ncc[k])
                    self.attributes.at[attribute, "size_connected_components"] = sizes
                    self.attributes.loc[attribute, "num_large_connected_components"] = int(nlarge[k])
            self.attributes.loc[self.attributes["num_connected_components"] > 1, "top"] = False
        if self.verbose:
            logging.info("Number of top attributes: %d" % np.sum(self.attributes["top"]))


    @_locked
    def define_domains(self, **kwargs):
        """safe.py:661-716.  The pairwise Jaccard distances between the nes_binary columns of the top attributes
        (the metric evaluation inside the reference's linkage(m, 'average', metric='jaccard')) are computed on the
        GPU from bit-packed columns; the O(k^2) average-linkage clustering of the k top attributes itself is SciPy's,
        as upstream.  Any other attribute_distance_metric goes to SciPy unchanged."""
        import pandas as pd
        from scipy.cluster.hierarchy import fcluster, linkage
        if "attribute_distance_threshold" in kwargs:
            self.attribute_distance_threshold = kwargs["attribute_distance_threshold"]
        self.validate_config()

        top = self.attributes["top"].values.astype(bool)
        if self.attribute_distance_metric == "jaccard":
            cols = np.flatnonzero(top)
            if len(cols) < 2:   # what scipy's linkage raises for a single observation
                raise ValueError("The number of observations cannot be determined on an empty distance matrix.")
            dist = _lib.jaccard(get_context(self.device), self.nes_binary, cols)
            Z = linkage(dist, method="average")
        else:
            Z = linkage(self.nes_binary[:, top].T, method="average", metric=self.attribute_distance_metric)
        max_d = np.max(Z[:, 2] * self.attribute_distance_threshold)
        domains = fcluster(Z, max_d, criterion="distance")

        self.attributes["domain"] = 0
        self.attributes.loc[self.attributes["top"], "domain"] = domains

        # node2nes_binary.groupby(level='domain', axis=1).sum() / node2nes.groupby(...).max(), safe.py:680-701
        domain = self.attributes["domain"].values
        ids = np.unique(domain)
        counts = np.stack([self.nes_binary[:, domain == d].sum(axis=1) for d in ids], axis=1)
        # pandas' groupby(...).max() skips NaN (z-score / invalid hypergeometric cells); an all-NaN group stays NaN
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            maxnes = np.stack([np.nanmax(self.nes[:, domain == d], axis=1) for d in ids], axis=1)
        self.node2domain = pd.DataFrame(counts, columns=pd.Index(ids, name="domain"))
        real = ids >= 1
        t_max = counts[:, real].max(axis=1)
        t_idxmax = ids[real][np.argmax(counts[:, real], axis=1)]
        t_idxmax[t_max == 0] = 0
        self.node2domain["primary_domain"] = t_idxmax
        col_of = {d: k for k, d in enumerate(ids)}
        self.node2domain["primary_nes"] = [maxnes[i, col_of[d]] for i, d in enumerate(t_idxmax)]

        if self.verbose:
            num_domains = len(np.unique(domains))
            per_domain = self.attributes.loc[self.attributes["domain"] > 0].groupby("domain")["id"].count()
            logging.info("Number of domains: %d (containing %d-%d attributes)"
                         % (num_domains, per_domain.min(), per_domain.max()))


class SAFE(SafeB200Mixin):
    """Standalone host class: same settings, validation and method names as safepy.safe.SAFE (safe.py:37-235) for
    the neighborhood/enrichment path, fed from in-memory objects instead of the reference's file loaders."""

    def __init__(self, path_to_ini_file="", path_to_safe_data=None, verbose=True, device=-1):
        self.verbose = verbose
        self.device = device
        self.path_to_safe_data = path_to_safe_data
        self.graph = None
        self.graph_euclidean = None
        self.node_key_attribute = "label_orf"
        self.attributes = None
        self.nodes = None
        self.node2attribute = None
        for k, v in DEFAULTS.items():
            setattr(self, k, v)
        self.neighborhoods = None
        self.node_distances = None
        self.ns = None
        self.pvalues_neg = None
        self.pvalues_pos = None
        self.nes = None
        self.nes_threshold = None
        self.nes_binary = None
        self.domains = None
        self.node2domain = None
        if path_to_ini_file:
            self.read_config(path_to_ini_file)
        self.validate_config()

    def read_config(self, path_to_ini_file):
        """The 'Input files' / 'Analysis parameters' keys safe.py:147-184 reads (others are ignored upstream too)."""
        import configparser
        cfg = configparser.ConfigParser(allow_no_value=True, comment_prefixes=("#", ";", "{"),
                                        inline_comment_prefixes="#")
        cfg.read(path_to_ini_file)

        def get(section, key, cast=str):
            for sec in (section, "DEFAULT"):
                if cfg.has_option(sec, key) and cfg.get(sec, key) not in (None, ""):
                    return cast(cfg.get(sec, key).strip())
            return None

        for attr, section, key, cast in (
                ("attribute_sign", "Input files", "annotationsign", str),
                ("background", "Analysis parameters", "background", str),
                ("node_distance_metric", "Analysis parameters", "nodeDistanceType", str),
                ("neighborhood_radius_type", "Analysis parameters", "neighborhoodRadiusType", str),
                ("neighborhood_radius", "Analysis parameters", "neighborhoodRadius", float),
                ("attribute_unimodality_metric", "Analysis parameters", "unimodalityType", str),
                ("attribute_distance_metric", "Analysis parameters", "groupDistanceType", str),
                ("attribute_distance_threshold", "Analysis parameters", "groupDistanceThreshold", float)):
            val = get(section, key, cast)
            if val is not None:
                setattr(self, attr, val)
        try:
            self.random_seed = int(get("Analysis parameters", "randomSeed"))
        except (ValueError, TypeError):
            self.random_seed = None

    def validate_config(self):
        """Same checks, messages and restore-the-default behaviour as safe.py:190-235."""
        def option(attr, valid, label=None):
            val = getattr(self, attr)
            if val not in valid:
                setattr(self, attr, DEFAULTS[attr])
                raise ValueError("%s is not a valid setting for %s. Valid options are: %s"
                                 % (val, label or attr, ", ".join(valid)))

        option("background", ["attribute_file", "network"])
        option("node_distance_metric", ["euclidean", "shortpath", "shortpath_weighted_layout"])
        option("attribute_sign", ["highest", "lowest", "both"])
        if not isinstance(self.num_permutations, (int, np.integer)) or self.num_permutations < 10:
            self.num_permutations = DEFAULTS["num_permutations"]
            raise ValueError("num_permutations must be an integer equal or greater than 10.")
        if not isinstance(self.enrichment_threshold, float) or not (0 < self.enrichment_threshold < 1):
            self.enrichment_threshold = DEFAULTS["enrichment_threshold"]
            raise ValueError("enrichment_threshold must be in the (0,1) range.")
        if not isinstance(self.enrichment_max_log10, (int, float)):
            self.enrichment_max_log10 = DEFAULTS["enrichment_max_log10"]
            raise ValueError("enrichment_max_log10 must be a number.")
        if not isinstance(self.attribute_enrichment_min_size, int) or self.attribute_enrichment_min_size < 2:
            self.attribute_enrichment_min_size = DEFAULTS["attribute_enrichment_min_size"]
            raise ValueError("attribute_enrichment_min_size must be an integer equal or greater than 2.")
        if not isinstance(self.attribute_distance_threshold, float) or not (0 < self.attribute_distance_threshold < 1):
            self.attribute_distance_threshold = DEFAULTS["attribute_distance_threshold"]
            raise ValueError("attribute_distance_threshold must be a float number in the (0,1) range.")

    # -- in-memory loaders (the reference's file formats are out of scope here)
    def load_network(self, graph=None, edges=None, x=None, y=None, length=None, weight=None, **kwargs):
        """Either an nx.Graph in the reference's conventions (nodes 0..N-1 with 'x', 'y'; edges with 'length'), or
        arrays: edges [E, 2], coordinates x, y and either edge lengths or adjacency weights (default 1): lengths are
        then computed on the device as safe_io.calculate_edge_lengths does (layout distance x weight,
        safe_io.py:311-333; an edge with weight 0 gets no 'length', like upstream).  The graph object is the single
        source of truth afterwards: define_neighborhoods re-reads its edge data on every call."""
        import networkx as nx
        if "node_key_attribute" in kwargs:
            self.node_key_attribute = kwargs["node_key_attribute"]
        self.validate_config()
        self._loaded_edge_arrays = None
        if graph is None:
            x = np.asarray(x, dtype=np.float64)
            y = np.asarray(y, dtype=np.float64)
            edges = np.zeros((0, 2), dtype=np.int64) if edges is None else np.asarray(edges, dtype=np.int64)
            if length is None and len(edges):
                length = _lib.edge_lengths(get_context(self.device), x, y, edges[:, 0], edges[:, 1], weight)
            graph = nx.Graph()
            graph.add_nodes_from((i, {"key": i, "x": float(x[i]), "y": float(y[i]), "label": str(i),
                                      self.node_key_attribute: str(i)}) for i in range(x.shape[0]))
            if len(edges):
                length = np.asarray(length, dtype=np.float64)
                graph.add_edges_from((int(u), int(v), {} if w != w else {"length": float(w)})
                                     for (u, v), w in zip(edges, length))
                lo, hi = np.minimum(edges[:, 0], edges[:, 1]), np.maximum(edges[:, 0], edges[:, 1])
                if len(np.unique(lo * x.shape[0] + hi)) == len(edges):
                    # no repeated edge (nx.Graph would keep only the last).  Used only under assume_graph_unchanged;
                    # an edge without 'length' costs Dijkstra's default 1 (safe.py:406-407)
                    self._loaded_edge_arrays = (graph, edges.copy(), np.where(np.isnan(length), 1.0, length))
        self.graph = graph

    def load_attributes(self, attribute_file=None, **kwargs):
        """DataFrame (index = node labels, as read_attributes accepts, safe_io.py:375-379) or a plain [N, M] array
        already aligned to node order."""
        import pandas as pd
        self.validate_config()
        if isinstance(attribute_file, pd.DataFrame):
            frame = attribute_file.apply(pd.to_numeric, errors="coerce")
            names = [str(c) for c in frame.columns]
            import networkx as nx
            labels = list(nx.get_node_attributes(self.graph, self.node_key_attribute).values())
            if not frame.index.is_unique:
                frame = frame.groupby(frame.index).mean()
            if labels:
                frame = frame.reindex(index=labels, fill_value=np.nan)
            values = frame.values
        else:
            values = np.asarray(attribute_file)
            if values.ndim == 1:
                values = values[:, None]
            names = [str(j) for j in range(values.shape[1])]
        self.attributes = pd.DataFrame({"id": np.arange(len(names)), "name": names})
        self.node2attribute = values


def accelerate(reference_class):
    """Subclass of the reference's SAFE whose neighborhood/enrichment methods run on the B200:

        from safepy import safe
        from safepy_b200 import accelerate
        SAFE = accelerate(safe.SAFE)
        sf = SAFE(); sf.load_network(); sf.define_neighborhoods(); sf.load_attributes(); sf.compute_pvalues()
    """
    return type("SAFE", (SafeB200Mixin, reference_class), {"__doc__": reference_class.__doc__})
