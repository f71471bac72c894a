"""L2-hot streaming rate of the production GEMM pipeline (no HBM traffic): what one B200 sustains when every CTA
streams k-tiles that are already L2-resident.  Separates the L2->SM / shared-memory ceiling from HBM effects."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from safepy_b200 import _lib, get_context  # noqa: E402

ctx = get_context()
for grid in (1, 32, 148):
    for ncols in (64, 128, 192):
        ktiles, slots = 32, 64
        ms = _lib.selftest_mma_rate(ctx, ncols, ktiles, slots, grid)
        it = grid * slots * ktiles
        ops = it * 2.0 * 128 * ncols * 64
        byt = it * (8192 + 64 * ncols)
        cyc = ms * 1e-3 * 1.965e9 / (slots * ktiles)
        print("grid=%3d ncols=%3d: %.3f ms  %.0f int8 TOPS  %.2f TB/s smem fill  ~%.0f cycles/k-tile@1.965GHz (MMA floor %d)"
              % (grid, ncols, ms, ops / ms / 1e9, byt / ms / 1e9, cyc, ncols), flush=True)
