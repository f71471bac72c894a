"""Device time and roofline bookkeeping of the non-GEMM kernels at the BASELINE.json shapes.

    python tools/kernel_bench.py [--only sssp,euclid,hypergeom,components,tail] [--configs C1,C3,C5] [--small]

Algorithmic bytes follow SURVEY.md section 8(d):
  k_sssp      sum over sources s and settled nodes t of (8 + deg(t) * 12) + N*N/8   (computed exactly from the result)
  k_euclid    16 N + N*N/8 bytes, 5 N*N fp64 flops
  k_hypergeom 20 bytes per (node, attribute) element (X in, p and NES out)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from safepy_b200 import _lib, get_context, synthetic as syn  # noqa: E402

PEAK = 6536.4
try:
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
        PEAK = json.load(f)["hbm_gbs"]
except OSError:
    pass


def timed(ctx, cls, fn, reps=3):
    fn()
    ctx.profile(True)
    ctx.kernel_ms(cls)
    for _ in range(reps):
        fn()
    ms, cnt = ctx.kernel_ms(cls)
    ctx.profile(False)
    return ms / max(cnt, 1)


def main():
    only = set(sys.argv[sys.argv.index("--only") + 1].replace("+", ",").split(",")) if "--only" in sys.argv else None
    sssp_cfgs = tuple(sys.argv[sys.argv.index("--configs") + 1].split(",")) if "--configs" in sys.argv else None
    small = "--small" in sys.argv
    ctx = get_context()
    out = []
    if only is None or "sssp" in only:
        for name in sssp_cfgs or (("C1",) if small else ("C1", "C3", "C5")):
            cfg = syn.make_config(name, shuffle=True)
            net, n = cfg["net"], cfg["n"]
            nr = cfg["radius"] * (net["x"].max() - net["x"].min())
            nb = _lib.Neighborhoods(ctx, n)
            ms = timed(ctx, "sssp", lambda: nb.shortpath(net["indptr"], net["indices"], net["csr_length"], nr))
            deg = np.diff(net["indptr"]).astype(np.float64)
            # settled (s, t) pairs: column sums of the matrix weight each node by how often it was settled
            sums = nb.rowsums().astype(np.float64)  # symmetric up to last-ulp effects: row sums ~ column sums
            pairs = float(sums.sum())
            alg = float(np.dot(sums, 8 + deg * 12)) + n * n / 8.0
            out.append(dict(kernel="k_sssp", workload=name, n=n, edges=int(len(net["edges"])), ms=ms,
                            settled_pairs=pairs, pairs_per_s=pairs / (ms * 1e-3), algorithmic_bytes=alg,
                            achieved_gbs=alg / (ms * 1e-3) / 1e9, peak_gbs=PEAK, frac=alg / (ms * 1e-3) / 1e9 / PEAK))
            nb.close()
    if only is None or "euclid" in only:
        cfg = syn.make_config("C4", 0.1 if small else 1.0, shuffle=True)
        net, n = cfg["net"], cfg["n"]
        nr = cfg["radius"] * (net["x"].max() - net["x"].min())
        nb = _lib.Neighborhoods(ctx, n)
        ms = timed(ctx, "euclid", lambda: nb.euclid(net["x"], net["y"], nr))
        alg = 16.0 * n + n * n / 8.0
        out.append(dict(kernel="k_euclid", workload="C4", n=n, ms=ms, pairs_per_s=float(n) * n / (ms * 1e-3),
                        algorithmic_bytes=alg, achieved_gbs=alg / (ms * 1e-3) / 1e9, peak_gbs=PEAK,
                        frac=alg / (ms * 1e-3) / 1e9 / PEAK, fp64_gflops=5.0 * n * n / (ms * 1e-3) / 1e9))
        nb.close()
    if only is None or "hypergeom" in only:
        cfg = syn.make_config("C2", 0.2 if small else 1.0, shuffle=True)
        net, n, m = cfg["net"], cfg["n"], cfg["m"]
        nr = cfg["radius"] * (net["x"].max() - net["x"].min())
        nb = _lib.Neighborhoods(ctx, n).shortpath(net["indptr"], net["indices"], net["csr_length"], nr)
        plan = _lib.Enrichment(nb, cfg["attributes"])
        t0 = time.perf_counter()
        plan.hypergeom()
        wall_first = time.perf_counter() - t0
        ms = timed(ctx, "hypergeom", lambda: plan.hypergeom())
        t0 = time.perf_counter()
        plan.hypergeom()
        wall = time.perf_counter() - t0
        alg = 20.0 * n * m
        out.append(dict(kernel="k_hypergeom", workload="C2", n=n, m=m, ms=ms, elements_per_s=n * m / (ms * 1e-3),
                        algorithmic_bytes=alg, achieved_gbs=alg / (ms * 1e-3) / 1e9, peak_gbs=PEAK,
                        frac=alg / (ms * 1e-3) / 1e9 / PEAK, call_wall_s=wall, first_call_wall_s=wall_first))
        plan.close()
        nb.close()
    if only is None or "components" in only:
        # define_top_attributes' connectivity test at C3 size: 2000 attributes, each enriched in 1-3 spatial blobs
        cfg = syn.make_config("C3", 0.1 if small else 1.0, shuffle=True)
        net, n, m = cfg["net"], cfg["n"], cfg["m"]
        rng = np.random.default_rng(3)
        nb = np.zeros((n, m), dtype=np.uint8)
        for j in range(m):
            for _ in range(rng.integers(1, 4)):
                c = rng.integers(0, n)
                d = np.hypot(net["x"] - net["x"][c], net["y"] - net["y"][c])
                nb[d < rng.uniform(0.03, 0.12), j] = 1
        cand = np.nonzero(nb.sum(axis=0) >= 10)[0]
        _lib.components(ctx, net["indptr"], net["indices"], nb, cand[:8], 10)
        t0 = time.perf_counter()
        ncc, nlarge, _ = _lib.components(ctx, net["indptr"], net["indices"], nb, cand, 10)
        wall = time.perf_counter() - t0
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import safe_oracle as orc
        t0 = time.perf_counter()
        k = min(len(cand), 100)
        sub = np.zeros_like(nb)
        sub[:, cand[:k]] = nb[:, cand[:k]]
        _, occ, _, _ = orc.top_attributes(net["indptr"], net["indices"], sub, 10)
        cpu = (time.perf_counter() - t0) / k
        assert np.array_equal(occ[cand[:k]], ncc[:k])
        out.append(dict(kernel="k_components", workload="C3 nodes x %d candidate attributes" % len(cand), n=n,
                        call_wall_s=wall, attributes_per_s=len(cand) / wall,
                        cpu_scipy_s_per_attribute=cpu, mean_components=float(ncc.mean())))
    if only is None or "tail" in only:
        # what follows the counts inside compute_pvalues at C3 size (20k x 2000): fused tail, row-wise FDR, and the
        # Jaccard distances define_domains needs between 1000 top attributes
        from safepy_b200.permutations import make_perm_rows
        cfg = syn.make_config("C3", 0.1 if small else 1.0, shuffle=True)
        net, n, m = cfg["net"], cfg["n"], cfg["m"]
        nr = cfg["radius"] * (net["x"].max() - net["x"].min())
        nb = _lib.Neighborhoods(ctx, n).shortpath(net["indptr"], net["indices"], net["csr_length"], nr)
        plan = _lib.Enrichment(nb, cfg["attributes"])
        plan.null_begin("sum", "auto")
        plan.null_add(make_perm_rows(cfg["attributes"], 16, 7))
        for fdr in (False, True):
            plan.null_finalize(16, multiple_testing=fdr)
            ctx.profile(True)
            ctx.kernel_ms("tail"), ctx.kernel_ms("fdr")
            t0 = time.perf_counter()
            plan.null_finalize(16, multiple_testing=fdr)
            wall = time.perf_counter() - t0
            ms_tail, _ = ctx.kernel_ms("tail")
            ms_fdr, _ = ctx.kernel_ms("fdr")
            ctx.profile(False)
            cells = float(n) * m
            # tail: 2 x 4 B counts + 8 B observed score in, 4 x 8 B out per cell
            alg = 48.0 * cells
            rec = dict(kernel="k_null_tail", workload="C3 finalize, multiple_testing=%s" % fdr, n=n, m=m, ms=ms_tail,
                       algorithmic_bytes=alg, achieved_gbs=alg / (ms_tail * 1e-3) / 1e9, peak_gbs=PEAK,
                       frac=alg / (ms_tail * 1e-3) / 1e9 / PEAK, call_wall_s=wall,
                       d2h_bytes=40.0 * cells, d2h_gbs_incl_kernels=40.0 * cells / wall / 1e9)
            if fdr:
                # per matrix: sort (8 B key + 4 B index in and out) + adjustment pass (12 B in, 8 B out); two matrices
                rec.update(fdr_ms=ms_fdr, fdr_algorithmic_bytes=2 * 44.0 * cells,
                           fdr_gbs=2 * 44.0 * cells / (ms_fdr * 1e-3) / 1e9)
            out.append(rec)
        plan.close()
        nb.close()
        rng = np.random.default_rng(5)
        nbin = np.zeros((n, m), dtype=np.uint8)
        for j in range(m):
            c = rng.integers(0, n)
            d = np.hypot(net["x"] - net["x"][c], net["y"] - net["y"][c])
            nbin[d < rng.uniform(0.03, 0.12), j] = 1
        cols = np.arange(0, m, 2)
        _lib.jaccard(ctx, nbin, cols[:4])
        ctx.profile(True)
        ctx.kernel_ms("jaccard")
        t0 = time.perf_counter()
        dist = _lib.jaccard(ctx, nbin, cols)
        wall = time.perf_counter() - t0
        ms, _ = ctx.kernel_ms("jaccard")
        ctx.profile(False)
        from scipy.spatial.distance import pdist
        t0 = time.perf_counter()
        k = min(len(cols), 200)
        ref = pdist(nbin[:, cols[:k]].T.astype(np.float64), metric="jaccard")
        cpu = time.perf_counter() - t0
        assert np.array_equal(_lib.jaccard(ctx, nbin, cols[:k]), ref)
        pairs = len(dist)
        out.append(dict(kernel="k_jaccard", workload="C3 nodes, %d top attributes" % len(cols), n=n, pairs=pairs, ms=ms,
                        pairs_per_s=pairs / (ms * 1e-3), l2_bytes=pairs * 2.0 * ((n + 31) // 32) * 4,
                        call_wall_s=wall, cpu_scipy_pairs_per_s=k * (k - 1) / 2 / cpu))
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
