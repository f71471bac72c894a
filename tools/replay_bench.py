"""Host-only: rate of the native replay of run_permutations' RNG stream (sb_perm_stream_*) against NumPy's legacy
generator doing the same draws, and a bit-for-bit comparison of what they produce.  No GPU involved.

    python tools/replay_bench.py [--n 20000] [--perms 300]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from safepy_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=20000)
    ap.add_argument("--perms", type=int, default=300)
    args = ap.parse_args()
    n, P = args.n, args.perms
    idx = np.nonzero(np.random.default_rng(0).random(n) > 0.05)[0]
    out = np.zeros((P, n), dtype=np.int32)          # touched once: no page faults inside the timed calls
    best = {"native_rows": 1e9, "native_skip": 1e9, "numpy_shuffle_only": 1e9, "numpy_rows": 1e9}
    for _ in range(5):
        s = _lib.PermStream(n, idx, 7)
        t0 = time.perf_counter()
        s.next(P, out)
        best["native_rows"] = min(best["native_rows"], time.perf_counter() - t0)
        t0 = time.perf_counter()
        s.skip(P)
        best["native_skip"] = min(best["native_skip"], time.perf_counter() - t0)
        np.random.seed(7)
        t0 = time.perf_counter()
        for _ in range(P):
            np.random.permutation(idx)
        best["numpy_shuffle_only"] = min(best["numpy_shuffle_only"], time.perf_counter() - t0)
        np.random.seed(7)
        cur = np.arange(n, dtype=np.int32)
        ref = np.empty_like(out)
        t0 = time.perf_counter()
        for p in range(P):
            cur[idx] = cur[np.random.permutation(idx)]
            ref[p] = cur
        best["numpy_rows"] = min(best["numpy_rows"], time.perf_counter() - t0)
    per = 1e9 / (P * len(idx))
    print(json.dumps({"n": n, "rows_with_data": int(len(idx)), "permutations": P,
                      "identical_rows": bool(np.array_equal(out, ref)),
                      "ns_per_permuted_row": {k: v * per for k, v in best.items()},
                      "speedup_rows": best["numpy_rows"] / best["native_rows"],
                      "cpu": open("/proc/cpuinfo").read().split("model name")[1].split("\n")[0].strip(": \t")
                      if os.path.exists("/proc/cpuinfo") else None,
                      "timing": "best of 5, one core"}))


if __name__ == "__main__":
    main()
