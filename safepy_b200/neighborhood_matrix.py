"""`SAFE.neighborhoods` without the N x N int64 matrix.

The reference stores neighborhoods as a dense int64 ndarray (safepy/safe.py:387: 8 N^2 bytes, 80 GB at 100k nodes)
and reads it back only through `np.sum(..., axis=1)`, `np.dot(neighborhoods, X)` and plain indexing.  This class
keeps the bit-packed rows the CUDA kernels produce (N^2 / 8 bytes) and answers those three uses; a dense copy is
made only on explicit request and only below a size limit.  Instances pickle (SAFE.save, safe.py:237-242): the
device handle is dropped and rebuilt on demand.
"""
import numpy as np

from . import _lib

DENSE_LIMIT_BYTES = 8 << 30


class PackedNeighborhoods:
    ndim = 2
    dtype = np.dtype(np.int64)  # what the reference's matrix reports

    def __init__(self, words, n, device=None):
        """`words` may be None when `device` holds the matrix: the host copy is then fetched on first use."""
        if words is None and device is None:
            raise ValueError("PackedNeighborhoods needs packed words or a device handle")
        self._words = None if words is None else \
            np.ascontiguousarray(words, dtype=np.uint32).reshape(n, _lib.neigh_ld(n))
        self.n = int(n)
        self._device = device  # _lib.Neighborhoods or None
        self.node_order = None  # optional locality hint for the enrichment plan (see ordering.py)

    @property
    def words(self):
        if self._words is None:
            self._words = self._device.packed()
        return self._words

    # -- ndarray-like surface
    @property
    def shape(self):
        return (self.n, self.n)

    def __len__(self):
        return self.n

    def row_sums(self):
        if self._device is not None and self._device.h is not None:
            return self._device.rowsums()
        return np.bitwise_count(self.words).sum(axis=1, dtype=np.int64)

    def sum(self, axis=None, **kwargs):
        rs = self.row_sums()
        if axis is None:
            return int(rs.sum())
        if axis in (1, -1):
            return rs
        if axis == 0:
            return self.dense(np.uint8).sum(axis=0, dtype=np.int64)
        raise ValueError("axis out of range")

    def dense(self, dtype=np.int64):
        dtype = np.dtype(dtype)
        need = self.n * self.n * dtype.itemsize
        if need > DENSE_LIMIT_BYTES:
            raise MemoryError("dense %d x %d %s neighborhood matrix needs %.1f GB; use .rows(r0, r1) or the packed "
                              "words instead" % (self.n, self.n, dtype, need / 2**30))
        return self.rows(0, self.n, dtype)

    def rows(self, r0, r1, dtype=np.uint8):
        return _lib.unpack_packed(self.words[r0:r1], self.n).astype(dtype, copy=False)

    def __array__(self, dtype=None, copy=None):
        return self.dense(dtype or np.int64)

    def __getitem__(self, key):
        if isinstance(key, tuple) and len(key) == 2 and all(isinstance(k, (int, np.integer)) for k in key):
            s, t = int(key[0]), int(key[1])
            return np.int64((self.words[s, t >> 5] >> np.uint32(t & 31)) & np.uint32(1))
        if isinstance(key, (int, np.integer)):
            return self.rows(int(key), int(key) + 1, np.int64)[0]
        return self.dense()[key]

    # -- device side
    def on_device(self, ctx):
        """The matrix as a device handle on `ctx` (uploaded once, then cached)."""
        dev = self._device
        if dev is None or dev.h is None or dev.ctx is not ctx:
            dev = _lib.Neighborhoods(ctx, self.n)
            dev.upload_packed(self.words)
            self._device = dev
        return dev

    def __getstate__(self):
        return {"words": self.words, "n": self.n, "node_order": self.node_order}

    def __setstate__(self, state):
        self._words = state["words"]
        self.n = state["n"]
        self.node_order = state.get("node_order")
        self._device = None


def as_packed(neighborhoods):
    """Accept what callers of the reference API may hand over: PackedNeighborhoods or a dense 0/1 array."""
    if isinstance(neighborhoods, PackedNeighborhoods):
        return neighborhoods
    dense = np.asarray(neighborhoods)
    if dense.ndim != 2 or dense.shape[0] != dense.shape[1]:
        raise ValueError("neighborhoods must be a square matrix")
    if np.any((dense != 0) & (dense != 1)):
        raise ValueError("neighborhoods must contain only 0 and 1")
    return PackedNeighborhoods(_lib.pack_dense(dense), dense.shape[0])
