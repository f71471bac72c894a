"""L2-hot streaming-rate experiments on the production GEMM pipeline (speed only, see sb_selftest_mma_rate).
With SB_TRACE=1 the library also prints the per-role cycle accounting of every run."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from safepy_b200 import _lib, get_context  # noqa: E402

ctx = get_context()


def run(tag, ncols, grid, dbg=0, kt=32, slots=64):
    ms = _lib.selftest_mma_rate(ctx, ncols, kt, slots, grid, dbg, None)
    it = grid * slots * kt
    ops = it * 2.0 * 128 * ncols * 64
    print("%-16s N=%3d grid=%3d ktiles=%4d slots=%4d: %.3f ms %5.0f TOPS  %.2f us/k-tile/CTA"
          % (tag, ncols, grid, kt, slots, ms, ops / ms / 1e9, ms * 1e3 / slots / kt), flush=True)


slots = int(sys.argv[1]) if len(sys.argv) > 1 else 512
grids = tuple(int(g) for g in sys.argv[2].split(",")) if len(sys.argv) > 2 else (1, 148)
widths = tuple(int(g) for g in sys.argv[3].split(",")) if len(sys.argv) > 3 else (64, 128, 192)
for grid in grids:
    for n in widths:
        for dbg, tag in ((0, "production"), (1, "copies only"), (2, "MMAs only"), (3, "barriers only"),
                         (6, "MMA issue only"), (7, "issue loop only")):
            run(tag, n, grid, dbg, 32, slots)
