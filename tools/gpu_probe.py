"""Diagnostics for the GPU box: which smem-descriptor variant reproduces the integer product, plus quick parity
probes.  Writes human-readable findings to stdout (run under gpurun, tee into gpurun_out/)."""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from safepy_b200 import _lib, get_context  # noqa: E402


def probe_mma(ctx):
    rng = np.random.default_rng(1)
    for ncols in (64, 192):
        for ktiles in (1, 3):
            k = 64 * ktiles
            a = (rng.uniform(size=(128, k)) < 0.3).astype(np.int8)
            b = rng.integers(-128, 128, size=(k, ncols), dtype=np.int64).astype(np.int8)
            ref = a.astype(np.int32) @ b.astype(np.int32)
            for variant in range(4):
                try:
                    d = _lib.selftest_mma_i8(ctx, a, b, variant)
                    bad = int((d != ref).sum())
                    print("mma ncols=%d ktiles=%d variant=%d mismatches=%d/%d" % (ncols, ktiles, variant, bad, d.size),
                          flush=True)
                    if bad and variant == 0:
                        rows = np.unique(np.argwhere(d != ref)[:, 0])
                        cols = np.unique(np.argwhere(d != ref)[:, 1])
                        print("   bad rows", rows[:16], "bad cols", cols[:16], "sample d", d[0, :6], "ref", ref[0, :6])
                except Exception as e:  # noqa: BLE001
                    print("mma ncols=%d ktiles=%d variant=%d ERROR %s" % (ncols, ktiles, variant, e), flush=True)
                    return
    # one-hot localisation for the production variant
    k = 64
    for r, kk, c in ((0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (8, 0, 0), (0, 8, 0), (0, 16, 0), (0, 32, 0), (0, 0, 16), (0, 0, 64)):
        a = np.zeros((128, k), dtype=np.int8); b = np.zeros((k, 192), dtype=np.int8)
        a[r, kk] = 1; b[kk, c] = 5
        d = _lib.selftest_mma_i8(ctx, a, b, 0)
        print("one-hot a[%d,%d] b[%d,%d] -> nonzero at" % (r, kk, kk, c), np.argwhere(d != 0)[:6].tolist(), flush=True)


def main():
    ctx = get_context()
    t = time.time()
    try:
        probe_mma(ctx)
    except Exception:  # noqa: BLE001
        traceback.print_exc()
    print("probe done in %.1fs" % (time.time() - t))


if __name__ == "__main__":
    main()
