#!/usr/bin/env python
"""Benchmark of SAFE's enrichment path on B200 (BASELINE.json metric: enrichment node-attr-perm scores/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C3] [--scale S]

Workload (default C3, BASELINE.json configs[2] -- the configuration the scores/s metric and the 1/2/4/8-GPU sharding are
quoted on): synthetic 20k-node / 150k-edge network, 2000 float32 attributes, shortpath_weighted_layout r=0.10,
1000 permutations.  One step = the whole permutation null: all P permutations scored against all N x M
(node, attribute) cells (N*M*P scores), permutations sharded over the ranks and combined by ONE all-reduce of the
count arrays (strong scaling).  `value` times it with neighborhoods, attributes and permutation indices resident in
HBM; `e2e` times the host-buffer C-ABI call (H2D of packed neighborhoods + attributes + indices, D2H of the counts,
inside the region).  `--workload C2` (configs[1], binary attributes) benchmarks the hypergeometric path instead:
one step = X = A @ B, p = hypergeom.sf(X - 1, ...), NES = -log10 p for all N x M cells (metric: elements/s).

Every run carries its own correctness evidence (`parity`): sampled (or all) stage-1 rows against the oracle's Dijkstra
/ pdist rows, and the counts of a short permutation pass -- through the same sharding and all-reduce as the timed
steps -- against the oracle's fp64 np.dot counts on sampled attribute columns (SURVEY 8d "parity gates run with
every benchmark").

--impl reference times the reference's CPU algorithm for the same path (oracle/safe_oracle.py restating
safepy/safe_extras.py:36-70) on the host cores, two ways: (A) as shipped -- np.dot with OpenBLAS on all cores,
(B) the reference's CLI pattern (safe.py:1335-1355) -- a multiprocessing pool over attribute chunks, one BLAS thread
per worker.  The line's value is the faster of the two.
"""
import os
import sys

if "--impl" in sys.argv and "reference" in sys.argv[sys.argv.index("--impl") + 1:][:1]:
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm must not inherit that (r1 SCALE records ran the
    # reference on one BLAS thread).  Set before NumPy loads OpenBLAS; the thread count actually used is re-checked
    # with threadpoolctl and printed.
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = str(os.cpu_count() or 1)

import argparse
import json
import subprocess
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (smoke runs only)")
    ap.add_argument("--perms", type=int, default=None)
    ap.add_argument("--cpu-sample-perms", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-safe-api", action="store_true", help="skip the SAFE-class wall-clock measurement")
    ap.add_argument("--engine", default="auto")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-run parity gate (profiling runs only)")
    ap.add_argument("--no-variant-b", action="store_true", help="reference arm: skip the multiprocessing variant")
    ap.add_argument("--no-ceiling", action="store_true",
                    help="skip the same-box tensor-rate probe reported as roofline.tensor_ceiling_same_box")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock + throttle reasons sampled every 100 ms while the timed region runs.

    NVML is read in-process (pynvml): an `nvidia-smi -lms` child polling the same fields was measured to stall CUDA
    memory-management calls of the benchmarked process for tens of milliseconds per query.  nvidia-smi remains the
    fallback when pynvml is unavailable."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx = device_index
        self.proc = None
        self.lines = []
        self.samples = []
        self.stop_flag = threading.Event()
        self.recording = threading.Event()
        self.call_ms = (0.0, 0.0)
        self.nvml = None
        self.how = None

    def start(self):
        """begin recording (prepare() must have run: NVML is initialised and polled before the timed region, so
        that its first-use costs do not land inside it)"""
        self.recording.set()

    def prepare(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = self.idx
            if visible:
                try:
                    phys = int(visible.split(",")[self.idx])
                except (ValueError, IndexError):
                    phys = self.idx
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            self.how = "nvml"
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:  # noqa: BLE001
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "500"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.how = "nvidia-smi"
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll_nvml(self):
        nv = self.nvml
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        try:
            smax = nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            smax = None
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop_flag.is_set():
            try:
                t0 = time.perf_counter()
                sm = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                t1 = time.perf_counter()
                mask = get_reasons(self.handle)
                t2 = time.perf_counter()
                self.call_ms = (max(self.call_ms[0], 1e3 * (t1 - t0)), max(self.call_ms[1], 1e3 * (t2 - t1)))
                if self.recording.is_set():
                    self.samples.append((float(sm), smax, {k for k, bit in names.items() if mask & bit}))
            except Exception:  # noqa: BLE001
                pass
            self.stop_flag.wait(0.1)

    def _pump(self):
        for line in self.proc.stdout:
            if self.recording.is_set():
                self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            sm = [x[0] for x in self.samples]
            smax = [x[1] for x in self.samples if x[1]]
            reasons = set().union(*[x[2] for x in self.samples]) if self.samples else set()
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(smax)) if smax else None,
                    "reasons": sorted(reasons), "samples": len(sm), "source": "nvml",
                    "max_query_ms": [round(self.call_ms[0], 2), round(self.call_ms[1], 2)]}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------------ workload
def build_workload(args):
    from safepy_b200 import synthetic as syn
    # nodes are renumbered randomly: the input order carries no locality, the library gets its hint from the layout
    cfg = syn.make_config(args.workload, args.scale, shuffle=True)
    if args.perms:
        cfg["perms"] = args.perms
    net = cfg["net"]
    cfg["nr"] = cfg["radius"] * (np.max(net["x"]) - np.min(net["x"]))
    return cfg


def workload_name(cfg, args):
    kind = {"normal32": "float32 N(0,1)", "binary": "binary GO-style", "dyadic": "dyadic-grid"}.get(cfg["kind"],
                                                                                                  cfg["kind"])
    return "%s%s: N=%d E=%d M=%d P=%d %s r=%.2f %s attributes" % (
        args.workload, "" if args.scale == 1.0 else "(scale %.3g)" % args.scale, cfg["n"], len(cfg["net"]["edges"]),
        cfg["m"], cfg["perms"], cfg["metric"], cfg["radius"], kind)


def syn_to_networkx(net):
    from safepy_b200 import synthetic as syn
    return syn.to_networkx(net)


def oracle_rows(cfg, rows):
    """Rows of the reference's neighborhood matrix (uint8 [len(rows), N]) from the oracle: scipy Dijkstra with the
    cutoff for the shortest-path metrics, pdist's arithmetic for 'euclidean'."""
    import safe_oracle as orc
    net = cfg["net"]
    if cfg["metric"] == "euclidean":
        return orc.neighborhoods_euclidean_rows(net["x"], net["y"], cfg["nr"], rows)
    length = net["csr_length"] if cfg["metric"] == "shortpath_weighted_layout" else None
    return orc.neighborhoods_shortpath_csr(net["indptr"], net["indices"], length, cfg["nr"], rows=rows)


# ------------------------------------------------------------------------------------------------ CPU reference arm
def blas_threads(want=None):
    """Set (if asked) and report the thread count OpenBLAS really uses."""
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        if want:
            threadpool_limits(limits=int(want))
        counts = [d.get("num_threads") for d in threadpool_info() if d.get("user_api") == "blas"]
        return int(max(counts)) if counts else None
    except Exception:  # noqa: BLE001
        return None


def cpu_reference_problem(cfg):
    """The dense int64 neighborhood matrix exactly as the reference holds it (safe.py:387), via the oracle, and the
    attribute matrix.  N = 100k cannot be allocated by the reference at all (80 GB): there a spatially contiguous
    20k-node sub-problem is timed instead and the rate is labelled 'extrapolated' (BASELINE.md section 3)."""
    n = cfg["n"]
    attrs = cfg["attributes"]
    if n * n * 8 <= (8 << 30):
        return oracle_rows(cfg, np.arange(n)).astype(np.int64), attrs, None
    from safepy_b200.ordering import kd_order
    sub = np.sort(kd_order(cfg["net"]["x"], cfg["net"]["y"])[:20000])
    nb = oracle_rows(cfg, sub)[:, sub].astype(np.int64)
    return nb, np.ascontiguousarray(attrs[sub]), "extrapolated from a %d-node sub-problem (the reference cannot " \
        "allocate its %d x %d int64 matrix)" % (len(sub), n, n)


def cpu_reference_steps(nb, attrs, steps, warmup):
    """Variant A, the reference's permutation loop body (safe_extras.py:56-66) as shipped: one permutation per step,
    np.dot on the dense int64 matrix with every BLAS thread."""
    import safe_oracle as orc
    np.random.seed(7)
    s0 = orc.compute_neighborhood_score(nb, attrs, "sum")
    n2a = np.copy(attrs)
    indx_vals = np.nonzero(np.sum(~np.isnan(n2a), axis=1))[0]
    counts_neg = np.zeros(s0.shape)
    counts_pos = np.zeros(s0.shape)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        n2a[indx_vals, :] = n2a[np.random.permutation(indx_vals), :]
        sp = orc.compute_neighborhood_score(nb, n2a, "sum")
        counts_neg = np.add(counts_neg, sp <= s0)
        counts_pos = np.add(counts_pos, sp >= s0)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return times


_VB = {}


def _variant_b_worker(task):
    cols, perms, seed, threads = task
    import safe_oracle as orc
    blas_threads(threads)
    t0 = time.perf_counter()
    orc.run_permutations(_VB["nb"], np.ascontiguousarray(_VB["attrs"][:, cols]), "sum", perms, seed)
    return time.perf_counter() - t0


def cpu_variant_b(nb, attrs, perms):
    """Variant B, "multiprocessing over all host cores": the reference's own processes= option is broken
    (safe.py:506-507 builds a 4-tuple, safe_extras.py:43 unpacks 5), so this follows its CLI pattern
    (safe.py:1335-1355): a pool over attribute chunks, one BLAS thread per worker, each worker running
    run_permutations (observed score + `perms` permutations) on its chunk.  M = 1: permutation chunks with distinct
    seeds.  Workers are capped by memory: every np.dot casts the int64 matrix to float64 (8 N^2 bytes per worker)."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    n, m = attrs.shape
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:  # noqa: BLE001
        avail = 64 << 30
    per_worker = n * n * 8 + 3 * n * max(1, m // cores) * 8 + (64 << 20)
    workers = int(max(1, min(cores, (0.6 * avail) // per_worker)))
    threads = max(1, cores // workers)
    if m >= workers:
        tasks = [(chunk, perms, 7, threads) for chunk in np.array_split(np.arange(m), workers)]
        total_perm_cols = float(m) * (perms + 1)     # the observed-score pass of every worker counts as one more
    else:
        tasks = [(np.arange(m), perms, 7 + k, threads) for k in range(workers)]
        total_perm_cols = float(m) * (perms + 1) * workers
    _VB["nb"], _VB["attrs"] = nb, attrs
    ctx = mp.get_context("fork")     # the dense matrix is shared copy-on-write, as in the reference's fork pool
    t0 = time.perf_counter()
    with ctx.Pool(workers) as pool:
        worker_s = pool.map(_variant_b_worker, tasks)
    wall = time.perf_counter() - t0
    _VB.clear()
    return {"value": n * total_perm_cols / wall, "unit": "scores/s", "workers": workers,
            "blas_threads_per_worker": threads, "wall_s": wall, "slowest_worker_s": float(max(worker_s)),
            "permutations_per_worker": perms}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    cfg = build_workload(args)
    if cfg["perms"] <= 0:
        return run_reference_hypergeom(args, cfg)
    cores = os.cpu_count() or 1
    threads = blas_threads(cores)
    nb, attrs, extrap = cpu_reference_problem(cfg)
    n_ref, m = attrs.shape
    times = cpu_reference_steps(nb, attrs, args.steps, args.warmup)
    t = float(np.sum(times))
    value_a = float(n_ref) * m * args.steps / t
    variants = {"A_openblas_threads": {"value": value_a, "unit": "scores/s", "blas_threads": threads,
                                       "seconds_per_permutation": t / args.steps}}
    best, best_name, best_cores = value_a, "A (np.dot, OpenBLAS on %s threads)" % threads, threads or cores
    if not args.no_variant_b:
        try:
            vb = cpu_variant_b(nb, attrs, max(1, min(args.steps, 2)))
            variants["B_multiprocessing_pool"] = vb
            if vb["value"] > best:
                best, best_name = vb["value"], "B (pool of %d workers x %d BLAS thread(s))" % (
                    vb["workers"], vb["blas_threads_per_worker"])
                best_cores = vb["workers"] * vb["blas_threads_per_worker"]
        except Exception as exc:  # noqa: BLE001
            variants["B_multiprocessing_pool"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
    ms = 1e3 * float(n_ref) * m / best
    sample = "%d timed permutation(s) of %d per variant; value = the faster variant, %s%s" % (
        args.steps, cfg["perms"], best_name, "; " + extrap if extrap else "")
    out = {
        "impl": "reference", "metric": "enrichment node-attr-perm scores/s", "value": best, "unit": "scores/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(cfg, args), "step": "one permutation of the reference loop "
                   "(row shuffle + np.dot of the dense int64 neighborhood matrix + 2 compares)"},
        "cpu_baseline": {"value": best, "unit": "scores/s", "cores": best_cores, "kind": "port",
                         "host_cores": cores, "blas_threads_measured": threads, "sample": sample,
                         "variants": variants,
                         "ran": "oracle/safe_oracle.py (statement-for-statement port of safe_extras.py:6-70; the "
                                "unmodified reference checkout does not travel to the GPU box)"},
        "e2e": {"value": best, "unit": "scores/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(json.dumps(out))


def _hg_worker(task):
    import safe_oracle as orc
    rows, cols = task
    t0 = time.perf_counter()
    orc.hypergeom_pvalues_block(_VB["nb"][rows], _VB["attrs"], cols)
    return time.perf_counter() - t0


def run_reference_hypergeom(args, cfg):
    """CPU arm of the hypergeometric workload: scipy.stats.hypergeom.sf as called at safe.py:596 (serial Boost ufunc)
    on a sampled block of cells per step -- single process as shipped, and a fork pool over column chunks."""
    import multiprocessing as mp
    import safe_oracle as orc
    cores = os.cpu_count() or 1
    n, m = cfg["n"], cfg["m"]
    attrs = cfg["attributes"]
    nb = oracle_rows(cfg, np.arange(n)).astype(np.int64)
    rng = np.random.default_rng(11)
    block_cols = min(m, 48)
    times = []
    for it in range(args.warmup + args.steps):
        cols = np.sort(rng.choice(m, block_cols, replace=False))
        t0 = time.perf_counter()
        orc.hypergeom_pvalues_block(nb, attrs, cols)
        if it >= args.warmup:
            times.append(time.perf_counter() - t0)
    t = float(np.sum(times))
    value_a = float(n) * block_cols * args.steps / t
    variants = {"A_single_process": {"value": value_a, "unit": "elements/s", "block": [n, block_cols]}}
    best, best_name, best_cores = value_a, "A (single process, as shipped)", 1
    if not args.no_variant_b:
        try:
            _VB["nb"], _VB["attrs"] = nb, attrs
            chunks = [np.sort(rng.choice(m, min(m, 24), replace=False)) for _ in range(cores)]
            t0 = time.perf_counter()
            with mp.get_context("fork").Pool(cores) as pool:
                pool.map(_hg_worker, [(np.arange(n), c) for c in chunks])
            wall = time.perf_counter() - t0
            _VB.clear()
            vb = float(n) * sum(len(c) for c in chunks) / wall
            variants["B_multiprocessing_pool"] = {"value": vb, "unit": "elements/s", "workers": cores, "wall_s": wall}
            if vb > best:
                best, best_name, best_cores = vb, "B (pool of %d workers over attribute chunks)" % cores, cores
        except Exception as exc:  # noqa: BLE001
            variants["B_multiprocessing_pool"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
    out = {
        "impl": "reference", "metric": "hypergeometric enrichment elements/s", "value": best, "unit": "elements/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(n) * m / best,
        "higher_is_better": True, "scaling": "replicas", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(cfg, args),
                   "step": "scipy.stats.hypergeom.sf + -log10 on a sampled [N, %d] block of cells (safe.py:573-608)"
                           % block_cols},
        "cpu_baseline": {"value": best, "unit": "elements/s", "cores": best_cores, "kind": "port",
                         "host_cores": cores, "variants": variants,
                         "sample": "%d block(s) of %d x %d cells; value = the faster variant, %s"
                                   % (args.steps, n, block_cols, best_name)},
        "e2e": {"value": best, "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(json.dumps(out))


def cpu_baseline_subprocess(args, steps):
    """cpu_baseline of the default run: the reference arm in a fresh CPU-only process (clean BLAS thread settings,
    and its fork pool never shares a process with a CUDA context)."""
    env = {k: v for k, v in os.environ.items()
           if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "OMP_NUM_THREADS",
                        "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "TORCHELASTIC_RUN_ID", "GROUP_RANK")}
    env["CUDA_VISIBLE_DEVICES"] = ""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", args.workload,
           "--scale", repr(args.scale), "--steps", str(steps), "--warmup", "1"]
    if args.perms:
        cmd += ["--perms", str(args.perms)]
    try:
        res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
        for ln in reversed(res.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)["cpu_baseline"]
        return {"value": None, "error": "reference arm printed no JSON line (exit %d): %s"
                % (res.returncode, res.stderr.strip()[-300:])}
    except Exception as exc:  # noqa: BLE001
        return {"value": None, "error": "%s: %s" % (type(exc).__name__, exc)}


# ------------------------------------------------------------------------------------------------ parity gate
def sample_rows_for_parity(n):
    """All rows up to 20k nodes (full matrix <= 20k), else 1000 sampled source rows."""
    if n <= 20000:
        return np.arange(n)
    return np.sort(np.random.default_rng(5).choice(n, 1000, replace=False))


def stage1_mismatches(nb, ref_rows, rows):
    """Bits of the library's packed matrix that differ from the oracle's rows."""
    from safepy_b200._lib import unpack_packed
    n = nb.n
    bad = 0
    if len(rows) == n:
        got = unpack_packed(nb.packed(), n)
        return int(np.count_nonzero(got != ref_rows))
    for k, r in enumerate(rows):
        bad += int(np.count_nonzero(nb.dense(int(r), int(r) + 1)[0] != ref_rows[k]))
    return bad


def counts_mismatches(ref_rows, rows, attrs, cols, perm_rows, cneg, cpos):
    """Cells of the device counts (already summed over ranks; uint32 [N, M]) that differ from the oracle's on the
    sampled node rows x attribute columns.  The oracle's np.dot runs on the float64 copy of the 0/1 rows (the values
    the reference's int64 -> float64 cast produces)."""
    import safe_oracle as orc
    a = ref_rows.astype(np.float64)
    sub = np.ascontiguousarray(attrs[:, cols])
    oneg, opos = orc.perm_counts_from_rows(a, sub, "sum", perm_rows)
    gneg = cneg[np.ix_(rows, cols)].astype(np.int64)
    gpos = cpos[np.ix_(rows, cols)].astype(np.int64)
    return int(np.count_nonzero(gneg != oneg) + np.count_nonzero(gpos != opos)), int(2 * oneg.size)


# ------------------------------------------------------------------------------------------------ ours
def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except OSError:
        return {}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from safepy_b200 import _lib
    from safepy_b200.permutations import make_perm_rows, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun (one rank per GPU)" % args.gpus)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))

    cfg = build_workload(args)
    if cfg["perms"] <= 0:
        return run_ours_hypergeom(args, cfg, world, rank, local_rank, dev)
    n, m, P = cfg["n"], cfg["m"], cfg["perms"]
    net, attrs = cfg["net"], cfg["attributes"]

    ctx = _lib.Context(local_rank, stream=torch.cuda.current_stream().cuda_stream)

    # ---- stage 1 (timed separately; reported as define_neighborhoods seconds).  With several ranks the source rows
    # are sharded (independent searches, no data-path collective inside) and the packed rows all-gathered once,
    # because the permutation-sharded stage 2 wants the whole matrix on every rank.
    from safepy_b200.distributed import row_shard
    ld = _lib.neigh_ld(n)
    r0, r1 = row_shard(n, world, rank)
    rows_pad = -(-n // world)                      # rows per rank in the gather buffer (last shard zero-padded)
    words_t = torch.zeros((world * rows_pad, ld), dtype=torch.int32, device=dev)

    def stage1():
        nbh = _lib.Neighborhoods(ctx, n, words_dev=words_t.data_ptr())
        if cfg["metric"] == "euclidean":
            nbh.euclid(net["x"], net["y"], cfg["nr"], r0, r1)
        else:
            nbh.shortpath(net["indptr"], net["indices"], net["csr_length"], cfg["nr"], r0, r1)
        if world > 1:
            mine = words_t[rank * rows_pad:(rank + 1) * rows_pad]
            if r0 != rank * rows_pad:                 # cannot happen with row_shard's equal-size shards
                raise RuntimeError("row shard does not line up with the gather buffer")
            dist.all_gather_into_tensor(words_t, mine.clone())
        return nbh

    stage1().close()
    words_t.zero_()
    ctx.profile(True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    nb = stage1()
    torch.cuda.synchronize()
    t_stage1 = time.perf_counter() - t0
    k1_ms, _ = ctx.kernel_ms("euclid" if cfg["metric"] == "euclidean" else "sssp")
    ctx.profile(False)
    rowsums = nb.rowsums()

    # ---- permutation indices: replay of the reference's legacy RNG stream (host, sequential by nature)
    t0 = time.perf_counter()
    rows_all = make_perm_rows(attrs, P, 7)
    t_rng = time.perf_counter() - t0
    # locality hint, computed from the layout exactly as SAFE.define_neighborhoods does (safepy_b200/safe.py)
    from safepy_b200.ordering import kd_order
    t0 = time.perf_counter()
    node_order = kd_order(net["x"], net["y"])
    t_order = time.perf_counter() - t0
    lo, hi = shard_bounds(P, world, rank)
    rows_host = torch.from_numpy(rows_all[lo:hi]).pin_memory()
    attrs_host = torch.from_numpy(attrs).pin_memory()
    packed_host = torch.from_numpy(nb.packed().view(np.int32)).pin_memory()

    rows_dev = rows_host.to(dev)
    attrs_dev = attrs_host.to(dev)
    packed_counts = P < 65536          # one uint32 word per cell (pos << 16 | neg) crosses NVLink instead of two
    counts = torch.zeros((2, n, m), dtype=torch.int32, device=dev)
    cpk = torch.zeros((n, m), dtype=torch.int32, device=dev) if packed_counts else None
    counts_host = torch.empty((2, n, m), dtype=torch.int32).pin_memory()
    torch.cuda.synchronize()

    def null_pass(plan, perm_ptr, nperm):
        """this rank's permutations -> counts (summed over ranks when world > 1)"""
        if packed_counts:
            cpk.zero_()
            if nperm:
                plan.perm_counts_packed_dev(perm_ptr, nperm, cpk.data_ptr(), "sum", args.engine)
            if world > 1:
                dist.all_reduce(cpk)
            plan.unpack_counts_dev(cpk.data_ptr(), counts[0].data_ptr(), counts[1].data_ptr())
        else:
            counts.zero_()
            if nperm:
                plan.perm_counts_dev(perm_ptr, nperm, counts[0].data_ptr(), counts[1].data_ptr(), "sum", args.engine)
            if world > 1:
                dist.all_reduce(counts)

    def step_resident():
        plan = _lib.Enrichment(nb, b_dev=attrs_dev.data_ptr(), dtype=np.float32, shape=(n, m))
        plan.set_node_order(node_order)
        null_pass(plan, rows_dev.data_ptr(), hi - lo)
        st = plan.stats()
        plan.close()
        return st

    def step_e2e():
        if world == 1:
            # the host-buffer C-ABI entry points: sb_neigh_upload_packed + sb_enrich_create + sb_enrich_perm_counts
            nbh = _lib.Neighborhoods(ctx, n).upload_packed(packed_host.numpy().view(np.uint32))
            plan = _lib.Enrichment(nbh, attrs_host.numpy())
            plan.set_node_order(node_order)
            out = counts_host.numpy().view(np.uint32)
            plan.perm_counts(rows_host.numpy(), "sum", args.engine, out=(out[0], out[1]))
            plan.close()
            nbh.close()
        else:
            pk = packed_host.to(dev, non_blocking=True)
            b = attrs_host.to(dev, non_blocking=True)
            r = rows_host.to(dev, non_blocking=True)
            nbh = _lib.Neighborhoods(ctx, n, words_dev=pk.data_ptr())
            plan = _lib.Enrichment(nbh, b_dev=b.data_ptr(), dtype=np.float32, shape=(n, m))
            plan.set_node_order(node_order)
            null_pass(plan, r.data_ptr(), hi - lo)
            if rank == 0:       # the result is wanted once (SAFE.results_rank), not on every rank
                counts_host.copy_(counts, non_blocking=True)
            torch.cuda.synchronize()
            plan.close()
            nbh.close()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        last = None
        walls = []
        for _ in range(steps):
            ts = time.perf_counter()
            last = fn()
            walls.append(1e3 * (time.perf_counter() - ts))
        timed.last_walls = walls
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms, wall * 1e3], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
        return ms, wall, last

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.prepare()
    # warm-up with the per-kernel event brackets already on (their first use allocates the driver's event pool, which
    # would otherwise land in the first timed step), then drop what the warm-up recorded
    ctx.profile(True)
    for _ in range(args.warmup):
        step_resident()
    for k in _lib.KERNEL_CLASSES:
        ctx.kernel_ms(k)
    launches0 = ctx.launch_count
    if rank == 0:
        sampler.start()
    ms_total, wall_total, stats = timed(step_resident, args.steps, 0)
    launches = ctx.launch_count - launches0
    step_walls = list(getattr(timed, "last_walls", []))
    kern = {k: ctx.kernel_ms(k) for k in ("gemm", "gather", "fixup", "prep", "score")}
    ctx.profile(False)
    clocks = sampler.stop() if rank == 0 else None
    torch.cuda.synchronize()
    final_counts = counts.cpu().numpy().view(np.uint32) if rank == 0 else None

    e2e_steps = max(1, min(args.steps, 3))
    e2e_ms, e2e_wall, _ = timed(step_e2e, e2e_steps, 1)
    agree = bool(np.array_equal(final_counts, counts_host.numpy().view(np.uint32))) if rank == 0 else True

    # ---- parity gate: every rank takes part in the short permutation pass (same sharding, same
    # all-reduce as the timed steps); rank 0 compares with the oracle
    parity = None
    if not args.no_parity:
        t0 = time.perf_counter()
        p_par = max(3, world)
        plo, phi = shard_bounds(p_par, world, rank)
        par_rows_dev = torch.from_numpy(np.ascontiguousarray(rows_all[plo:phi])).to(dev) if phi > plo else None
        plan = _lib.Enrichment(nb, b_dev=attrs_dev.data_ptr(), dtype=np.float32, shape=(n, m))
        plan.set_node_order(node_order)
        null_pass(plan, par_rows_dev.data_ptr() if par_rows_dev is not None else 0, phi - plo)
        plan.close()
        torch.cuda.synchronize()
        if rank == 0:
            par_counts = counts.cpu().numpy().view(np.uint32)
            rows_s = sample_rows_for_parity(n)
            ref_rows = oracle_rows(cfg, rows_s)
            bad1 = stage1_mismatches(nb, ref_rows, rows_s)
            cols = np.sort(np.random.default_rng(9).choice(m, min(m, 16), replace=False))
            bad2, cells = counts_mismatches(ref_rows, rows_s, attrs, cols, rows_all[:p_par], par_counts[0],
                                            par_counts[1])
            # size-independent identities on the counts of the TIMED run (all P permutations, all cells)
            tot = final_counts[0].astype(np.int64) + final_counts[1].astype(np.int64)
            bad3 = int(np.count_nonzero(tot < P) + np.count_nonzero(final_counts[0] > P) +
                       np.count_nonzero(final_counts[1] > P))
            parity = {"stage1_rows": int(len(rows_s)), "stage1_bits": int(len(rows_s)) * n,
                      "count_cells": cells, "count_columns": int(len(cols)), "count_permutations": p_par,
                      "timed_run_identity_cells": int(tot.size),
                      "mismatches": bad1 + bad2 + bad3,
                      "mismatches_by_gate": {"stage1": bad1, "counts": bad2, "timed_run_identities": bad3},
                      "resident_and_e2e_counts_equal": agree,
                      "oracle": "oracle/safe_oracle.py: scipy Dijkstra / pdist rows; fp64 np.dot counts for the same "
                                "host-generated permutation indices, after the all-reduce",
                      "seconds": time.perf_counter() - t0}
            if parity["mismatches"] or not agree:
                sys.stderr.write("PARITY GATE FAILED: %r\n" % (parity,))

    if rank == 0:
        peaks = load_peaks()
        peak_tf = peaks.get("bf16_tflops_sustained", 1408.3)
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PF sustained"
        scores_per_step = float(n) * m * P
        sec_per_step = ms_total / 1e3 / args.steps
        value = scores_per_step / sec_per_step
        gemm_ms, gemm_launches = kern["gemm"]
        # algorithmic FLOPs of the score GEMM: 2 * (cells of non-empty A tiles) * M per permutation,
        # digit passes and padding are implementation factors and are NOT counted
        tiles = stats["a_tiles"]
        tile_rows = stats.get("tile_rows", 256)
        flops_total = 2.0 * tiles * tile_rows * 64 * m * (hi - lo) * args.steps
        achieved_tf = flops_total / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else None
        int8_ops = 2.0 * stats["ktile_iters"] * tile_rows * 64 * 64 * stats["digits"] * args.steps
        nnz = float(rowsums.sum())
        # DRAM traffic of the dominant kernel per launch, from the committed ncu --set full capture of this same
        # command (bytes per permutation x permutations per launch); null when no capture of this round exists
        traffic, traffic_src = None, None
        for name in ("r2_gemm_traffic.json",):
            try:
                with open(os.path.join(ROOT, "profiles", name)) as f:
                    tr = json.load(f)
                if args.workload == "C3" and args.scale == 1.0 and gemm_launches:
                    traffic = tr["dram_bytes_per_permutation"] * (hi - lo) * args.steps / gemm_launches
                    traffic_src = "profiles/" + name
            except (OSError, KeyError, ValueError):
                pass
        out = {
            "metric": "enrichment node-attr-perm scores/s", "value": value, "unit": "scores/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int8 digits -> int32/int64",
            "data": "synthetic",
            "config": {
                "workload": workload_name(cfg, args),
                "step": "whole permutation null (operand prep + slab-ordered row gather + tcgen05 digit GEMM with fused "
                        "compare + fp64 fix-up%s)" % (" + NCCL all-reduce of the packed counts" if world > 1 else ""),
                "parallelism": "stage 2: permutations sharded %d-way + one all-reduce of the counts; stage 1: source "
                               "rows sharded %d-way + one all-gather of the packed rows" % (world, world),
                "l2": "per step the kernel streams %.0f GB of gathered operand rows through the SMs (digit planes "
                      "%.0f MB + count / observed-score arrays %.0f MB); inputs are re-uploaded / re-planned every "
                      "step, no result is cached between steps"
                      % (stats["ktile_iters"] * 64.0 * 64 * stats["digits"] * 1e-9,
                         n * ((m + 63) // 64 * 64) * stats["digits"] / 1e6, n * m * 12 / 1e6),
                "node_order": "input nodes randomly renumbered; k-d tree order of the layout passed as a hint "
                              "(sb_enrich_set_node_order)",
                "mean_neighborhood": float(rowsums.mean()), "nonempty_a_tiles": tiles,
                "a_tile_shape": [tile_rows, 64], "a_tile_fill": nnz / max(1.0, tiles * tile_rows * 64.0),
                "dense_a_tiles": stats["a_tiles_dense"], "digits": stats["digits"],
                "fixup_fraction": stats["fixups"] / max(1, stats["fixups"] + stats["decided"]),
            },
            "e2e": {
                "value": scores_per_step / (e2e_ms / 1e3 / e2e_steps), "unit": "scores/s",
                "h2d_bytes_per_step": int(packed_host.numel() * 4 + attrs_host.numel() * 4 + rows_host.numel() * 4),
                "d2h_bytes_per_step": int(counts_host.numel() * 4),
                "ms_per_step": e2e_ms / e2e_steps,
                "path": "sb_neigh_upload_packed + sb_enrich_create + sb_enrich_perm_counts (host buffers, pinned)"
                        if world == 1 else "pinned H2D on every rank + sb_enrich_perm_counts_packed_dev + NCCL "
                                           "all-reduce of the packed counts + D2H on rank 0",
                "resident_and_e2e_counts_equal": agree,
            },
            "parity": parity,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {
                "kernel": "k_gemm<%d> (tcgen05.mma.kind::i8, %d launches)" % (stats["digits"], gemm_launches),
                "bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_tf if achieved_tf else None, "traffic": traffic,
                "traffic_note": "DRAM read+write bytes per k_gemm launch from the ncu --set full capture summarised in "
                                "%s (bytes per permutation x permutations per launch)" % traffic_src
                                if traffic_src else "no ncu capture of this round's kernel committed yet",
                "peak_source": peak_src,
                "step_level_frac": (flops_total / (ms_total / 1e3) / 1e12) / peak_tf if ms_total else None,
                "nnz_frac": (2.0 * nnz * m * (hi - lo) * args.steps / (gemm_ms / 1e3) / 1e12) / peak_tf
                            if gemm_ms > 0 else None,
                "executed_int8_tops": int8_ops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else None,
                "gemm_share_of_step": gemm_ms / ms_total if ms_total else None,
                "kernel_ms_per_step": {k: v[0] / args.steps for k, v in kern.items()},
                "host_wall_ms_of_each_step": step_walls,
            },
            "stages": {"define_neighborhoods_s": t_stage1, "define_neighborhoods_kernel_ms": k1_ms,
                       "perm_index_replay_host_s": t_rng, "node_order_hint_host_s": t_order,
                       "compute_pvalues_null_s": sec_per_step},
        }
    # ---- what the tensor pipe of THIS chip sustains on operands with this null's statistics (random digits, masks
    # filled like the non-empty A tiles), L2-hot, through the same k_gemm pipeline with a store-nothing epilogue: the
    # chip runs at its power cap and the draw depends on the data, so the executed rate is compared with this ceiling
    # measured in the same process rather than with a nominal peak (speed-only probe, sb_selftest_mma_rate)
    if rank == 0 and not args.no_ceiling:
        try:
            fill = out["config"]["a_tile_fill"]
            ncols, kt, slots, sms = 64 * stats["digits"], 32, 32768, 148
            os.environ["SB_RATE_RANDOM"] = str(max(1, min(100, int(round(100 * fill)))))
            ceil = {}
            for tag, dbg in (("pipeline_l2_hot", 0), ("no_bulk_copies", 2)):
                ms = _lib.selftest_mma_rate(ctx, ncols, kt, slots, sms, dbg)
                ceil[tag + "_int8_tops"] = sms * slots * kt * 2.0 * 128 * ncols * 64 / ms / 1e9
                ceil[tag + "_ms"] = ms
            del os.environ["SB_RATE_RANDOM"]
            ex = out["roofline"]["executed_int8_tops"]
            ceil["executed_frac_of_pipeline_l2_hot"] = ex / ceil["pipeline_l2_hot_int8_tops"] if ex else None
            ceil["executed_frac_of_no_bulk_copies"] = ex / ceil["no_bulk_copies_int8_tops"] if ex else None
            ceil["note"] = ("k_gemm pipeline on %d L2-resident k-tiles, random int8 digits, masks %s %% filled, %d "
                            "accumulations per CTA pair; second figure without the bulk copies (MMAs + A expansion "
                            "only)" % (kt, os.environ.get("SB_RATE_RANDOM", str(int(round(100 * fill)))), slots))
            out["roofline"]["tensor_ceiling_same_box"] = ceil
        except Exception as exc:  # noqa: BLE001
            out["roofline"]["tensor_ceiling_same_box"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
    # ---- neighborhood_score_type='z-score' on the same inputs (value / square / non-NaN planes in one accumulation per
    # permutation): a short resident pass of this rank's first permutations, reported beside the 'sum' null
    zscore = None
    try:
        zp = int(min(hi - lo, 96))
        if zp > 0:
            plan = _lib.Enrichment(nb, b_dev=attrs_dev.data_ptr(), dtype=np.float32, shape=(n, m))
            plan.set_node_order(node_order)
            counts.zero_()
            plan.perm_counts_dev(rows_dev.data_ptr(), zp, counts[0].data_ptr(), counts[1].data_ptr(), "z-score", "auto")
            torch.cuda.synchronize()
            counts.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.perm_counts_dev(rows_dev.data_ptr(), zp, counts[0].data_ptr(), counts[1].data_ptr(), "z-score", "auto")
            e1.record()
            torch.cuda.synchronize()
            zst = plan.stats()
            plan.close()
            zms = e0.elapsed_time(e1)
            zscore = {"permutations": zp, "ms_per_permutation": zms / zp, "scores_per_s": float(n) * m * zp / (zms / 1e3),
                      "fixup_fraction": zst["fixups"] / max(1, zst["fixups"] + zst["decided"]),
                      "note": "resident z-score null of this rank's first %d permutations (digit GEMM on value / "
                              "square / non-NaN planes, comparison in the epilogue + exact fix-ups)" % zp}
    except Exception as exc:  # noqa: BLE001
        zscore = {"error": "%s: %s" % (type(exc).__name__, exc)}
    if rank == 0:
        out["stages"]["zscore_null"] = zscore
    if not args.no_safe_api:
        # BASELINE.json's second metric, through the SAFE class itself (host call to host return: graph -> CSR,
        # layout order, RNG replay, H2D / D2H, NES arithmetic all included).  With several ranks the class shards
        # over the caller's torch.distributed group and only rank 0 receives the [N, M] result arrays.
        api = None
        try:
            from safepy_b200 import SAFE
            sf = SAFE(verbose=False, device=local_rank)
            sf.graph = syn_to_networkx(net)
            sf.node_distance_metric = cfg["metric"]
            sf.neighborhood_radius = cfg["radius"]
            sf.random_seed = 7
            sf.results_rank = 0
            sf.load_attributes(attribute_file=attrs)
            for rep in range(2):        # the second pass is the warm one
                barrier()
                t0 = time.perf_counter()
                sf.define_neighborhoods()
                t_dn = time.perf_counter() - t0
                t0 = time.perf_counter()
                sf.compute_pvalues(num_permutations=P)
                t_cp = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([t_dn, t_cp], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                t_dn, t_cp = float(t[0]), float(t[1])
            api = {"define_neighborhoods_s": t_dn, "compute_pvalues_s": t_cp, "total_s": t_dn + t_cp,
                   "metric": "define_neighborhoods+compute_pvalues sec (BASELINE.json's second metric)",
                   "compute_pvalues_phases_s": getattr(sf, "last_enrichment_seconds", None),
                   "note": "safepy_b200.SAFE.define_neighborhoods() + compute_pvalues(num_permutations=%d) on the "
                           "same workload, host wall clock of the second call, max over %d rank(s) (graph object "
                           "already built; the edge data is re-read from it on every call)" % (P, world)}
        except Exception as exc:  # noqa: BLE001  (the API timing must never take the benchmark line down)
            api = {"error": "%s: %s" % (type(exc).__name__, exc)}
        if rank == 0:
            out["stages"]["safe_api"] = api
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline_subprocess(args, args.cpu_sample_perms)
        emit(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and parity is not None and (parity["mismatches"] or not agree):
        sys.exit(3)


def run_ours_hypergeom(args, cfg, world, rank, local_rank, dev):
    """--workload C2: the hypergeometric path (safe.py:556-608) as its own workload.  One step = observed integer
    scores X = A @ nan0(B), group / neighborhood sizes, p = hypergeom.sf(X - 1, ...) and NES = -log10 p for every
    (node, attribute) cell.  The path does not shard (one 26M-cell problem, milliseconds): single GPU only."""
    import torch
    from safepy_b200 import _lib
    if world > 1:
        raise SystemExit("the hypergeometric workload is a single-GPU benchmark (replicas only)")
    import safe_oracle as orc
    n, m = cfg["n"], cfg["m"]
    net, attrs = cfg["net"], cfg["attributes"]
    ctx = _lib.Context(local_rank, stream=torch.cuda.current_stream().cuda_stream)
    ctx.profile(True)
    t0 = time.perf_counter()
    nb = _lib.Neighborhoods(ctx, n)
    if cfg["metric"] == "euclidean":
        nb.euclid(net["x"], net["y"], cfg["nr"])
    else:
        nb.shortpath(net["indptr"], net["indices"], net["csr_length"], cfg["nr"])
    torch.cuda.synchronize()
    t_stage1 = time.perf_counter() - t0
    rowsums = nb.rowsums()
    attrs_host = torch.from_numpy(attrs).pin_memory()
    packed_host = torch.from_numpy(nb.packed().view(np.int32)).pin_memory()
    attrs_dev = attrs_host.to(dev)
    pv = torch.empty((n, m), dtype=torch.float64, device=dev)
    nes = torch.empty((n, m), dtype=torch.float64, device=dev)
    pv_host = np.empty((n, m), dtype=np.float64)
    nes_host = np.empty((n, m), dtype=np.float64)

    def step_resident():
        plan = _lib.Enrichment(nb, b_dev=attrs_dev.data_ptr(), dtype=np.float32, shape=(n, m))
        plan.hypergeom_dev(pv.data_ptr(), nes.data_ptr())
        plan.close()

    def step_e2e():
        nbh = _lib.Neighborhoods(ctx, n).upload_packed(packed_host.numpy().view(np.uint32))
        plan = _lib.Enrichment(nbh, attrs_host.numpy())
        plan.lib.sb_enrich_hypergeom(plan.h, pv_host.ctypes.data, nes_host.ctypes.data)
        plan.close()
        nbh.close()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    sampler = ClockSampler(local_rank)
    sampler.prepare()
    for _ in range(args.warmup):
        step_resident()
    for k in _lib.KERNEL_CLASSES:
        ctx.kernel_ms(k)
    launches0 = ctx.launch_count
    sampler.start()
    steps = max(args.steps, 20)      # a step is a few milliseconds
    ms_total = timed(step_resident, steps, 0)
    launches = ctx.launch_count - launches0
    kern = {k: ctx.kernel_ms(k) for k in ("hypergeom", "score", "gemm", "prep")}
    ctx.profile(False)
    clocks = sampler.stop()
    e2e_steps = 3
    e2e_ms = timed(step_e2e, e2e_steps, 1)

    # ---- parity gate: >= 1e5 sampled cells against scipy.stats.hypergeom.sf (safe.py:596) + stage-1 rows
    parity = None
    if not args.no_parity:
        t0 = time.perf_counter()
        rows_s = sample_rows_for_parity(n)
        ref_rows = oracle_rows(cfg, rows_s)
        bad1 = stage1_mismatches(nb, ref_rows, rows_s)
        rng = np.random.default_rng(9)
        cell_rows = np.sort(rng.choice(len(rows_s), min(len(rows_s), 4096), replace=False))
        cols = np.sort(rng.choice(m, min(m, 32), replace=False))
        pref, nref = orc.hypergeom_pvalues_block(ref_rows[cell_rows].astype(np.int64), attrs, cols)
        got_p = pv.cpu().numpy()[np.ix_(rows_s[cell_rows], cols)]
        got_n = nes.cpu().numpy()[np.ix_(rows_s[cell_rows], cols)]
        bad = np.isnan(got_n) != np.isnan(nref)
        big = np.isinf(nref) | (nref > 300)
        bad |= (np.isinf(got_n) | (got_n > 300)) != big
        ok = ~np.isnan(nref) & ~big
        hi_ = ok & (np.abs(nref) >= 1e-3)
        lo_ = ok & ~hi_
        with np.errstate(invalid="ignore"):
            bad |= hi_ & ~(np.abs(got_n - nref) <= 1e-6 * np.abs(nref))
            bad |= lo_ & ~(np.abs(got_n - nref) <= 1e-12)
            bad |= (got_p == 1.0) != (pref == 1.0)
            rel = np.abs(got_n - nref)[hi_] / np.abs(nref[hi_])
        bad2 = int(np.count_nonzero(bad))
        agree = bool(np.array_equal(pv.cpu().numpy(), pv_host, equal_nan=True) and
                     np.array_equal(nes.cpu().numpy(), nes_host, equal_nan=True))
        parity = {"stage1_rows": int(len(rows_s)), "stage1_bits": int(len(rows_s)) * n,
                  "nes_cells": int(nref.size), "nes_tolerance": "1e-6 relative where |NES| >= 1e-3, 1e-12 absolute "
                  "below; same NaN / inf / p == 1 positions", "max_relative_error": float(rel.max()) if rel.size else 0.0,
                  "mismatches": bad1 + bad2, "mismatches_by_gate": {"stage1": bad1, "nes": bad2},
                  "resident_and_e2e_equal": agree,
                  "oracle": "oracle/safe_oracle.py::hypergeom_pvalues_block (scipy.stats.hypergeom.sf)",
                  "seconds": time.perf_counter() - t0}
        if parity["mismatches"] or not agree:
            sys.stderr.write("PARITY GATE FAILED: %r\n" % (parity,))

    peaks = load_peaks()
    peak = peaks.get("hbm_gbs", 6650.0)
    cells = float(n) * m
    sec = ms_total / 1e3 / steps
    hg_ms, hg_launches = kern["hypergeom"]
    hg_s = hg_ms / 1e3 / max(1, hg_launches)
    achieved = 20.0 * cells / hg_s / 1e9 if hg_s > 0 else None
    out = {
        "metric": "hypergeometric enrichment elements/s", "value": cells / sec, "unit": "elements/s", "n_gpus": 1,
        "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec, "higher_is_better": True,
        "scaling": "replicas", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(cfg, args),
                   "step": "plan (CSR view of the packed matrix) + exact integer scores X = A @ B + group / "
                           "neighborhood sizes + fused hypergeometric survival function and -log10",
                   "l2": "outputs (2 x %.0f MB fp64) exceed the 126 MB L2" % (cells * 8 / 1e6),
                   "mean_neighborhood": float(rowsums.mean())},
        "e2e": {"value": cells / (e2e_ms / 1e3 / e2e_steps), "unit": "elements/s",
                "h2d_bytes_per_step": int(packed_host.numel() * 4 + attrs_host.numel() * 4),
                "d2h_bytes_per_step": int(2 * cells * 8), "ms_per_step": e2e_ms / e2e_steps,
                "path": "sb_neigh_upload_packed + sb_enrich_create + sb_enrich_hypergeom (host buffers)"},
        "parity": parity, "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"kernel": "k_hypergeom (%d launches)" % hg_launches, "bound": "hbm", "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak if achieved else None, "traffic": None,
                     "algorithmic_bytes_per_element": 20,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s",
                     "note": "SURVEY 8d: 4 B X in + 8 B p + 8 B NES per element; the kernel is fp64 special-function "
                             "bound in practice (lgamma-table pmf seed + ratio recurrence per element)",
                     "kernel_ms_per_step": {k: v[0] / steps for k, v in kern.items()}},
        "stages": {"define_neighborhoods_s": t_stage1, "compute_pvalues_hypergeom_s": sec},
    }
    if not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_subprocess(args, 1)
    emit(json.dumps(out))
    if parity is not None and parity["mismatches"]:
        sys.exit(3)


_REAL_STDOUT = None


def emit(line):
    """The one JSON line goes to the process's original stdout; everything else that libraries print (NCCL banners,
    warnings) is routed to stderr for the whole run."""
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (line + "\n").encode())


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
