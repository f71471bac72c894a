"""GPU: the production tcgen05 kernel (bulk-copy pipeline, smem descriptors, int8 MMA, TMEM loads) on raw operands
against an integer matrix product."""
import numpy as np
import pytest

from safepy_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ncols", [64, 128, 192])
@pytest.mark.parametrize("ktiles", [1, 2, 9, 40])
def test_mma_i8_matches_integer_product(ctx, ncols, ktiles):
    rng = np.random.default_rng(ncols * 100 + ktiles)
    k = 64 * ktiles
    a = (rng.uniform(size=(128, k)) < 0.3).astype(np.int8)
    b = rng.integers(-128, 128, size=(k, ncols), dtype=np.int64).astype(np.int8)
    d = _lib.selftest_mma_i8(ctx, a, b)
    ref = a.astype(np.int32) @ b.astype(np.int32)
    assert np.array_equal(d, ref)


def test_mma_i8_structured_operands(ctx):
    """One-hot operands localise any layout error to a (row, k) or (k, column) coordinate."""
    k = 128
    for r, kk, c in ((0, 0, 0), (5, 17, 3), (127, 63, 191), (64, 64, 100), (9, 127, 64)):
        a = np.zeros((128, k), dtype=np.int8)
        b = np.zeros((k, 192), dtype=np.int8)
        a[r, kk] = 1
        b[kk, c] = -7
        d = _lib.selftest_mma_i8(ctx, a, b)
        ref = np.zeros((128, 192), dtype=np.int32)
        ref[r, c] = -7
        assert np.array_equal(d, ref), (r, kk, c, np.argwhere(d != 0)[:8].tolist())
