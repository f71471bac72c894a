"""ctypes binding of libsafe_b200.so (the C ABI declared in include/safe_b200.h).

Nothing here computes: every method forwards to one C entry point and raises `SafeB200Error` with the library's
message when the call fails.  There is no CPU fallback -- if the shared library is missing or no sm_100 device is
visible, importing works (so that CPU-only tooling can introspect the ABI) but creating a `Context` raises.
"""
import ctypes as C
import os
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# SAFE_B200_LIB selects another build of the library (A/B kernel experiments on one GPU box); default: the in-tree one
LIB_PATH = os.environ.get("SAFE_B200_LIB") or os.path.join(HERE, "libsafe_b200.so")

SB_F32, SB_F64 = 0, 1
SCORE_TYPES = {"sum": 0, "z-score": 1}
ENGINES = {"auto": 0, "simt": 1, "tc": 2}
ATTRIBUTE_SIGNS = {"highest": 0, "lowest": 1, "both": 2}
KERNEL_CLASSES = {"gemm": 0, "gather": 1, "fixup": 2, "sssp": 3, "euclid": 4, "hypergeom": 5, "score": 6, "prep": 7,
                  "tail": 8, "fdr": 9, "jaccard": 10}

_vp = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int32

# name -> (restype, argtypes); mirrors include/safe_b200.h one to one (tests/test_abi.py checks both directions)
SIGNATURES = {
    "sb_abi_version": (C.c_int, []),
    "sb_last_error": (C.c_char_p, []),
    "sb_ctx_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "sb_ctx_destroy": (C.c_int, [_vp]),
    "sb_current_device": (C.c_int, []),
    "sb_ctx_device": (C.c_int, [_vp]),
    "sb_ctx_set_stream": (C.c_int, [_vp, _vp]),
    "sb_ctx_synchronize": (C.c_int, [_vp]),
    "sb_ctx_release_memory": (C.c_int, [_vp]),
    "sb_ctx_launch_count": (_i64, [_vp]),
    "sb_ctx_profile": (C.c_int, [_vp, C.c_int]),
    "sb_ctx_kernel_ms": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_double), C.POINTER(_i64)]),
    "sb_host_register": (C.c_int, [_vp, _i64]),
    "sb_host_unregister": (C.c_int, [_vp]),
    "sb_neigh_ld": (_i64, [_i64]),
    "sb_neigh_create": (C.c_int, [_vp, _i64, C.POINTER(_vp)]),
    "sb_neigh_wrap_dev": (C.c_int, [_vp, _i64, _vp, C.POINTER(_vp)]),
    "sb_neigh_destroy": (C.c_int, [_vp]),
    "sb_neigh_n": (_i64, [_vp]),
    "sb_neigh_words_dev": (_vp, [_vp]),
    "sb_neigh_shortpath": (C.c_int, [_vp, _vp, _vp, _vp, C.c_double, _i64, _i64]),
    "sb_neigh_euclid": (C.c_int, [_vp, _vp, _vp, C.c_double, _i64, _i64]),
    "sb_neigh_upload_packed": (C.c_int, [_vp, _vp, _i64, _i64]),
    "sb_neigh_download_packed": (C.c_int, [_vp, _vp, _i64, _i64]),
    "sb_neigh_rowsums": (C.c_int, [_vp, _vp]),
    "sb_neigh_unpack_rows": (C.c_int, [_vp, _i64, _i64, C.c_int, _vp]),
    "sb_enrich_create": (C.c_int, [_vp, _vp, _vp, C.c_int, _i64, _i64, C.POINTER(_vp)]),
    "sb_enrich_create_dev": (C.c_int, [_vp, _vp, _vp, C.c_int, _i64, _i64, C.POINTER(_vp)]),
    "sb_enrich_destroy": (C.c_int, [_vp]),
    "sb_enrich_set_node_order": (C.c_int, [_vp, _vp]),
    "sb_enrich_score": (C.c_int, [_vp, C.c_int, _vp]),
    "sb_enrich_score_dev": (C.c_int, [_vp, C.c_int, _vp]),
    "sb_enrich_perm_counts": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _i64, _vp, _vp]),
    "sb_enrich_perm_counts_dev": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _i64, _vp, _vp]),
    "sb_enrich_perm_counts_packed_dev": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _i64, _vp]),
    "sb_counts_unpack_dev": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "sb_enrich_stats": (C.c_int, [_vp, _vp]),
    "sb_enrich_hypergeom": (C.c_int, [_vp, _vp, _vp]),
    "sb_enrich_hypergeom_dev": (C.c_int, [_vp, _vp, _vp]),
    "sb_enrich_attr_summary": (C.c_int, [_vp, _vp, C.POINTER(_i64)]),
    "sb_enrich_null_begin": (C.c_int, [_vp, C.c_int, C.c_int]),
    "sb_enrich_null_add": (C.c_int, [_vp, _vp, _i64]),
    "sb_enrich_null_counts": (C.c_int, [_vp, C.POINTER(_i64), _vp, _vp]),
    "sb_enrich_null_counts_dev": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp)]),
    "sb_enrich_null_packed_dev": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_i64)]),
    "sb_enrich_null_set_perms": (C.c_int, [_vp, _i64]),
    "sb_enrich_observed_rows_dev": (C.c_int, [_vp, C.c_int, _i64, _i64, C.POINTER(_vp)]),
    "sb_enrich_observed_set_ready": (C.c_int, [_vp, C.c_int]),
    "sb_perm_stream_create": (C.c_int, [_i64, _vp, _i64, C.c_int, C.c_uint32, C.POINTER(_vp)]),
    "sb_perm_stream_destroy": (C.c_int, [_vp]),
    "sb_perm_stream_next": (C.c_int, [_vp, _i64, _vp]),
    "sb_perm_stream_state": (C.c_int, [_vp, _vp, C.POINTER(_i32), C.POINTER(_i64)]),
    "sb_perm_stream_prefetch": (C.c_int, [_vp, _i64, C.c_int, C.c_int]),
    "sb_enrich_null_add_stream": (C.c_int, [_vp, _vp, _i64]),
    "sb_enrich_null_add_stream_shard": (C.c_int, [_vp, _vp, _i64, C.c_int, C.c_int]),
    "sb_enrich_null_finalize": (C.c_int, [_vp, _vp, _vp, _i64, C.c_int, C.c_double, C.c_int, C.c_double,
                                          _vp, _vp, _vp, _vp, _vp, _vp]),
    "sb_enrich_hypergeom_finalize": (C.c_int, [_vp, C.c_int, C.c_double, _vp, _vp, _vp, _vp]),
    "sb_fdr_rows": (C.c_int, [_vp, _i64, _i64, _vp, _vp]),
    "sb_attr_jaccard": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _vp]),
    "sb_graph_edge_lengths": (C.c_int, [_vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "sb_graph_csr": (C.c_int, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(_i64)]),
    "sb_graph_components": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _i32, _vp, _vp, _vp]),
    "sb_selftest_mma_i8": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]),
    "sb_selftest_mma_rate": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.POINTER(C.c_double)]),
}


class SafeB200Error(RuntimeError):
    pass


_lib = None
_lock = threading.Lock()


def load_library(path=None):
    """dlopen libsafe_b200.so and attach the prototypes. Raises SafeB200Error if it has not been built."""
    global _lib
    with _lock:
        if _lib is not None and path is None:
            return _lib
        p = path or LIB_PATH
        if not os.path.exists(p):
            raise SafeB200Error(
                "%s is missing: build it with `python -m safepy_b200.build` (needs nvcc); "
                "safepy_b200 has no CPU fallback" % p)
        lib = C.CDLL(p)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.sb_abi_version() != 1:
            raise SafeB200Error("libsafe_b200.so ABI version %d, expected 1" % lib.sb_abi_version())
        if path is None:
            _lib = lib
        return lib


def _check(lib, rc):
    if rc != 0:
        raise SafeB200Error(lib.sb_last_error().decode("utf-8", "replace"))


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _as(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class Context:
    """One CUDA device (sb_ctx). `stream` may be a raw cudaStream_t integer (e.g. torch's current stream)."""

    def __init__(self, device=-1, stream=None):
        self.lib = load_library()
        h = _vp()
        _check(self.lib, self.lib.sb_ctx_create(int(device), C.byref(h)))
        self.h = h
        # one context = one caller thread at a time (include/safe_b200.h): callers that share a context between
        # threads (safepy_b200.get_context hands out one per device) serialise on this lock
        self.lock = threading.RLock()
        if stream is not None:
            self.set_stream(stream)

    @property
    def device(self):
        """CUDA device ordinal this context runs on."""
        return int(self.lib.sb_ctx_device(self.h))

    def set_stream(self, stream):
        _check(self.lib, self.lib.sb_ctx_set_stream(self.h, _vp(int(stream))))

    def synchronize(self):
        _check(self.lib, self.lib.sb_ctx_synchronize(self.h))

    def release_memory(self):
        """Return the library's cached device memory (pool + scratch) to the driver."""
        _check(self.lib, self.lib.sb_ctx_release_memory(self.h))

    @property
    def launch_count(self):
        return int(self.lib.sb_ctx_launch_count(self.h))

    def profile(self, enable=True):
        _check(self.lib, self.lib.sb_ctx_profile(self.h, int(bool(enable))))

    def kernel_ms(self, kernel_class):
        """(milliseconds, brackets) accumulated for one kernel class since the last query; see KERNEL_CLASSES."""
        ms, cnt = C.c_double(), _i64()
        _check(self.lib, self.lib.sb_ctx_kernel_ms(self.h, KERNEL_CLASSES[kernel_class], C.byref(ms), C.byref(cnt)))
        return ms.value, cnt.value

    def close(self):
        if getattr(self, "h", None):
            self.lib.sb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def current_device():
    """Ordinal of the calling thread's current CUDA device (-1 without one)."""
    return int(load_library().sb_current_device())


def neigh_ld(n):
    """Words per packed row (pure arithmetic, mirrors sb_neigh_ld without needing the library)."""
    return ((int(n) + 31) // 32 + 3) // 4 * 4


class Neighborhoods:
    """Bit-packed N x N neighborhood matrix on the device (sb_neigh)."""

    def __init__(self, ctx, n, words_dev=None):
        self.ctx = ctx
        self.lib = ctx.lib
        self.n = int(n)
        self.ld = neigh_ld(n)
        h = _vp()
        if words_dev is None:
            _check(self.lib, self.lib.sb_neigh_create(ctx.h, self.n, C.byref(h)))
        else:
            _check(self.lib, self.lib.sb_neigh_wrap_dev(ctx.h, self.n, _vp(int(words_dev)), C.byref(h)))
        self.h = h

    # -- stage 1
    def shortpath(self, indptr, indices, length, cutoff, row0=0, row1=None):
        indptr = _as(indptr, np.int64)
        indices = _as(indices, np.int32)
        if indptr.shape[0] != self.n + 1:
            raise ValueError("indptr must have n + 1 entries")
        if length is not None:
            length = _as(length, np.float64)
            if length.shape[0] != indices.shape[0]:
                raise ValueError("length and indices differ in size")
        row1 = self.n if row1 is None else row1
        _check(self.lib, self.lib.sb_neigh_shortpath(self.h, _ptr(indptr), _ptr(indices), _ptr(length),
                                                     float(cutoff), int(row0), int(row1)))
        return self

    def euclid(self, x, y, nr, row0=0, row1=None):
        x = _as(x, np.float64)
        y = _as(y, np.float64)
        if x.shape != (self.n,) or y.shape != (self.n,):
            raise ValueError("x and y must have n entries")
        row1 = self.n if row1 is None else row1
        _check(self.lib, self.lib.sb_neigh_euclid(self.h, _ptr(x), _ptr(y), float(nr), int(row0), int(row1)))
        return self

    def upload_packed(self, words, row0=0, row1=None):
        row1 = self.n if row1 is None else row1
        words = _as(words, np.uint32)
        if words.size != (row1 - row0) * self.ld:
            raise ValueError("packed block has the wrong size")
        _check(self.lib, self.lib.sb_neigh_upload_packed(self.h, _ptr(words), int(row0), int(row1)))
        return self

    def upload_dense(self, dense):
        """Pack a dense 0/1 matrix on the host (np.packbits) and upload it."""
        return self.upload_packed(pack_dense(dense))

    # -- readback
    def packed(self, row0=0, row1=None):
        row1 = self.n if row1 is None else row1
        out = np.empty((row1 - row0, self.ld), dtype=np.uint32)
        _check(self.lib, self.lib.sb_neigh_download_packed(self.h, _ptr(out), int(row0), int(row1)))
        return out

    def rowsums(self):
        out = np.empty(self.n, dtype=np.int64)
        _check(self.lib, self.lib.sb_neigh_rowsums(self.h, _ptr(out)))
        return out

    def dense(self, row0=0, row1=None, dtype=np.uint8):
        row1 = self.n if row1 is None else row1
        dtype = np.dtype(dtype)
        if dtype.itemsize not in (1, 8):
            raise ValueError("dense(): dtype must be 1 or 8 bytes wide")
        out = np.empty((row1 - row0, self.n), dtype=dtype)
        _check(self.lib, self.lib.sb_neigh_unpack_rows(self.h, int(row0), int(row1), dtype.itemsize, _ptr(out)))
        return out

    @property
    def words_dev(self):
        return int(self.lib.sb_neigh_words_dev(self.h) or 0)

    def close(self):
        if getattr(self, "h", None):
            self.lib.sb_neigh_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pack_dense(dense):
    """Dense 0/1 [n, n] -> packed uint32 [n, ld] in the library's bit order (bit t&31 of word t>>5)."""
    dense = np.asarray(dense)
    n = dense.shape[0]
    if dense.ndim != 2 or dense.shape[1] != n:
        raise ValueError("neighborhood matrix must be square")
    ld = neigh_ld(n)
    bits = np.zeros((n, ld * 32), dtype=np.uint8)
    bits[:, :n] = dense != 0
    return np.packbits(bits, axis=1, bitorder="little").view(np.uint32).reshape(n, ld)


def unpack_packed(words, n):
    """Packed uint32 [rows, ld] -> dense uint8 [rows, n] (host helper for tests)."""
    words = np.ascontiguousarray(words, dtype=np.uint32)
    bits = np.unpackbits(words.view(np.uint8), axis=1, bitorder="little")
    return bits[:, :n]


def _attr_matrix(b):
    b = np.asarray(b)
    if b.ndim != 2:
        raise ValueError("attribute matrix must be 2-D")
    if b.dtype == np.float32:
        return np.ascontiguousarray(b), SB_F32
    return np.ascontiguousarray(b, dtype=np.float64), SB_F64


class Enrichment:
    """Stage-2 plan (sb_enrich): neighborhoods x attribute matrix."""

    def __init__(self, neigh, b=None, b_dev=None, dtype=None, shape=None):
        self.neigh = neigh
        self.ctx = neigh.ctx
        self.lib = neigh.lib
        h = _vp()
        if b_dev is None:
            b, code = _attr_matrix(b)
            self.n, self.m = b.shape
            _check(self.lib, self.lib.sb_enrich_create(self.ctx.h, neigh.h, _ptr(b), code, self.n, self.m,
                                                       C.byref(h)))
        else:
            self.n, self.m = shape
            code = SB_F32 if np.dtype(dtype) == np.float32 else SB_F64
            _check(self.lib, self.lib.sb_enrich_create_dev(self.ctx.h, neigh.h, _vp(int(b_dev)), code, self.n,
                                                           self.m, C.byref(h)))
        self.h = h

    def set_node_order(self, order):
        """Locality hint for the tensor-core null: order[i] = node at internal position i (None: identity)."""
        if order is not None:
            order = _as(order, np.int32)
            if order.shape != (self.n,):
                raise ValueError("order must have n entries")
        _check(self.lib, self.lib.sb_enrich_set_node_order(self.h, _ptr(order)))
        return self

    def score(self, score_type="sum"):
        out = np.empty((self.n, self.m), dtype=np.float64)
        _check(self.lib, self.lib.sb_enrich_score(self.h, SCORE_TYPES[score_type], _ptr(out)))
        return out

    def perm_counts(self, perm_rows, score_type="sum", engine="auto", out=None):
        perm_rows = _as(perm_rows, np.int32)
        if perm_rows.ndim != 2 or perm_rows.shape[1] != self.n:
            raise ValueError("perm_rows must be [num_permutations, n]")
        if out is None:
            cneg = np.empty((self.n, self.m), dtype=np.uint32)
            cpos = np.empty((self.n, self.m), dtype=np.uint32)
        else:
            cneg, cpos = out
            for a in (cneg, cpos):
                if a.dtype != np.uint32 or a.shape != (self.n, self.m) or not a.flags.c_contiguous:
                    raise ValueError("out arrays must be C-contiguous uint32 [n, m]")
        _check(self.lib, self.lib.sb_enrich_perm_counts(self.h, SCORE_TYPES[score_type], ENGINES[engine],
                                                        _ptr(perm_rows), perm_rows.shape[0], _ptr(cneg),
                                                        _ptr(cpos)))
        return cneg, cpos

    def perm_counts_dev(self, perm_dev, num_perm, cneg_dev, cpos_dev, score_type="sum", engine="auto"):
        _check(self.lib, self.lib.sb_enrich_perm_counts_dev(self.h, SCORE_TYPES[score_type], ENGINES[engine],
                                                            _vp(int(perm_dev)), int(num_perm), _vp(int(cneg_dev)),
                                                            _vp(int(cpos_dev))))

    def perm_counts_packed_dev(self, perm_dev, num_perm, packed_dev, score_type="sum", engine="auto"):
        """Counts added to one uint32 word per cell, pos << 16 | neg (device array [n, m], zeroed by the caller)."""
        _check(self.lib, self.lib.sb_enrich_perm_counts_packed_dev(self.h, SCORE_TYPES[score_type], ENGINES[engine],
                                                                   _vp(int(perm_dev)), int(num_perm),
                                                                   _vp(int(packed_dev))))

    def unpack_counts_dev(self, packed_dev, cneg_dev, cpos_dev):
        """packed [n, m] -> counts_neg, counts_pos (stored, device arrays)."""
        _check(self.lib, self.lib.sb_counts_unpack_dev(self.ctx.h, _vp(int(packed_dev)), self.n * self.m,
                                                       _vp(int(cneg_dev)), _vp(int(cpos_dev))))

    def attr_summary(self):
        """(NaNs per attribute column, number of values that are neither 0, 1 nor NaN) -- safe.py:453-458."""
        nans = np.empty(self.m, dtype=np.int64)
        other = _i64()
        _check(self.lib, self.lib.sb_enrich_attr_summary(self.h, _ptr(nans), C.byref(other)))
        return nans, other.value

    # -- streaming null: counts stay on the device until null_finalize / null_counts
    def null_begin(self, score_type="sum", engine="auto"):
        _check(self.lib, self.lib.sb_enrich_null_begin(self.h, SCORE_TYPES[score_type], ENGINES[engine]))
        return self

    def null_add(self, perm_rows):
        perm_rows = _as(perm_rows, np.int32)
        if perm_rows.ndim != 2 or perm_rows.shape[1] != self.n:
            raise ValueError("perm_rows must be [num_permutations, n]")
        _check(self.lib, self.lib.sb_enrich_null_add(self.h, _ptr(perm_rows), perm_rows.shape[0]))
        return self

    def null_add_stream(self, stream, num_perm, world=1, rank=0):
        """Count the next `num_perm` permutations of a PermStream (replay and device work overlap inside the call).
        With world > 1 the permutations are dealt round-robin in pieces and only this rank's pieces are counted; the
        others are drawn and dropped so that every rank's stream stays aligned."""
        _check(self.lib, self.lib.sb_enrich_null_add_stream_shard(self.h, stream.h, int(num_perm), int(world),
                                                                  int(rank)))
        return self

    def null_counts(self, want_counts=True):
        """(permutations counted so far, counts_neg, counts_pos)"""
        num = _i64()
        cneg = np.empty((self.n, self.m), dtype=np.uint32) if want_counts else None
        cpos = np.empty((self.n, self.m), dtype=np.uint32) if want_counts else None
        _check(self.lib, self.lib.sb_enrich_null_counts(self.h, C.byref(num), _ptr(cneg), _ptr(cpos)))
        return num.value, cneg, cpos

    def null_counts_dev(self):
        """Device addresses (neg, pos) of the open null's count arrays; pos == neg + 4 * n * m."""
        neg, pos = _vp(), _vp()
        _check(self.lib, self.lib.sb_enrich_null_counts_dev(self.h, C.byref(neg), C.byref(pos)))
        return int(neg.value), int(pos.value)

    def null_packed_dev(self):
        """(device address of the packed count words [n, m] (pos << 16 | neg), permutations they hold)."""
        pk, held = _vp(), _i64()
        _check(self.lib, self.lib.sb_enrich_null_packed_dev(self.h, C.byref(pk), C.byref(held)))
        return int(pk.value), held.value

    def observed_rows_dev(self, score_type, row0, row1):
        """Compute observed-score rows [row0, row1) into the plan's [n, m] fp64 array; returns its device address."""
        ptr = _vp()
        _check(self.lib, self.lib.sb_enrich_observed_rows_dev(self.h, SCORE_TYPES[score_type], int(row0), int(row1),
                                                              C.byref(ptr)))
        return int(ptr.value)

    def observed_set_ready(self, score_type):
        _check(self.lib, self.lib.sb_enrich_observed_set_ready(self.h, SCORE_TYPES[score_type]))

    def null_set_perms(self, num_perm):
        _check(self.lib, self.lib.sb_enrich_null_set_perms(self.h, int(num_perm)))

    def null_finalize(self, num_permutations, attribute_sign="both", enrichment_threshold=0.05,
                      multiple_testing=False, want=("ns", "pvalues_neg", "pvalues_pos", "nes", "nes_binary")):
        """Counts -> everything compute_pvalues leaves on the SAFE object (safe.py:526-554, 466-472).  The p-value and
        NES of every possible count are computed here with NumPy, by the reference's own expressions, and looked up
        on the device."""
        num_permutations = int(num_permutations)
        counts = np.arange(num_permutations + 1, dtype=np.float64)
        ptab = counts / num_permutations                                            # safe.py:532-533
        floor = 1 / num_permutations
        nestab = -np.log10(np.where(ptab == 0, floor, ptab))                        # safe.py:546-547
        thr = float(-np.log10(enrichment_threshold))                                # safe.py:468
        out = {k: np.empty((self.n, self.m), dtype=np.float64) for k in want}
        out["num_neighborhoods_enriched"] = np.empty(self.m, dtype=np.float64)
        _check(self.lib, self.lib.sb_enrich_null_finalize(
            self.h, _ptr(ptab), _ptr(nestab), ptab.shape[0], int(bool(multiple_testing)), float(floor),
            ATTRIBUTE_SIGNS[attribute_sign], thr, _ptr(out.get("ns")), _ptr(out.get("pvalues_neg")),
            _ptr(out.get("pvalues_pos")), _ptr(out.get("nes")), _ptr(out.get("nes_binary")),
            _ptr(out["num_neighborhoods_enriched"])))
        return out

    def hypergeom_finalize(self, enrichment_threshold=0.05, multiple_testing=False,
                           want=("pvalues_pos", "nes", "nes_binary")):
        """Hypergeometric p-values, optional row-wise FDR, NES, nes_binary and the enriched-neighborhood counts
        (safe.py:573-608, 466-472)."""
        thr = float(-np.log10(enrichment_threshold))
        out = {k: np.empty((self.n, self.m), dtype=np.float64) for k in want}
        out["num_neighborhoods_enriched"] = np.empty(self.m, dtype=np.float64)
        _check(self.lib, self.lib.sb_enrich_hypergeom_finalize(
            self.h, int(bool(multiple_testing)), thr, _ptr(out.get("pvalues_pos")), _ptr(out.get("nes")),
            _ptr(out.get("nes_binary")), _ptr(out["num_neighborhoods_enriched"])))
        return out

    def stats(self):
        out = np.zeros(7, dtype=np.int64)
        _check(self.lib, self.lib.sb_enrich_stats(self.h, _ptr(out)))
        keys = ["decided", "fixups", "a_tiles", "a_tiles_dense", "digits", "ktile_iters", "overflow_batches"]
        return dict(zip(keys, (int(v) for v in out)))

    def hypergeom(self, want_pvalues=True, want_nes=True):
        pv = np.empty((self.n, self.m), dtype=np.float64) if want_pvalues else None
        nes = np.empty((self.n, self.m), dtype=np.float64) if want_nes else None
        _check(self.lib, self.lib.sb_enrich_hypergeom(self.h, _ptr(pv), _ptr(nes)))
        return pv, nes

    def hypergeom_dev(self, pvalues_dev, nes_dev):
        """Hypergeometric p-values / NES into device arrays ([n, m] fp64 each; either may be 0 = not wanted)."""
        _check(self.lib, self.lib.sb_enrich_hypergeom_dev(self.h, _vp(int(pvalues_dev)) if pvalues_dev else None,
                                                          _vp(int(nes_dev)) if nes_dev else None))

    def close(self):
        if getattr(self, "h", None):
            self.lib.sb_enrich_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PermStream:
    """run_permutations' RNG stream (np.random.seed + np.random.permutation per iteration, applied cumulatively)
    replayed by the library on the host.  seed: int in [0, 2**32) or None (OS entropy, like np.random.seed(None))."""

    def __init__(self, n, rows_with_data, seed):
        self.lib = load_library()
        self.n = int(n)
        idx = _as(rows_with_data, np.int64)
        if seed is not None and not (0 <= int(seed) < 2 ** 32):
            raise ValueError("Seed must be between 0 and 2**32 - 1")     # numpy's message
        h = _vp()
        _check(self.lib, self.lib.sb_perm_stream_create(self.n, _ptr(idx), idx.shape[0], int(seed is not None),
                                                        int(seed or 0), C.byref(h)))
        self.h = h

    def next(self, num_perm, out=None):
        """Gather rows int32 [num_perm, n] of the next permutations."""
        rows = out if out is not None else np.empty((int(num_perm), self.n), dtype=np.int32)
        _check(self.lib, self.lib.sb_perm_stream_next(self.h, int(num_perm), _ptr(rows)))
        return rows

    def skip(self, num_perm):
        _check(self.lib, self.lib.sb_perm_stream_next(self.h, int(num_perm), None))

    def prefetch(self, num_perm, world=1, rank=0):
        """Start drawing this rank's share of the next num_perm permutations in the background; the following
        Enrichment.null_add_stream(self, num_perm, world, rank) consumes them as they become ready."""
        _check(self.lib, self.lib.sb_perm_stream_prefetch(self.h, int(num_perm), int(world), int(rank)))
        return self

    def state(self):
        """(key[624] uint32, pos, permutations drawn)"""
        key = np.empty(624, dtype=np.uint32)
        pos, drawn = _i32(), _i64()
        _check(self.lib, self.lib.sb_perm_stream_state(self.h, _ptr(key), C.byref(pos), C.byref(drawn)))
        return key, pos.value, drawn.value

    def sync_numpy(self):
        """Leave NumPy's legacy global generator where the reference's calls would have left it."""
        key, pos, _ = self.state()
        np.random.set_state(("MT19937", key, pos, 0, 0.0))

    def close(self):
        if getattr(self, "h", None):
            self.lib.sb_perm_stream_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def edge_lengths(ctx, x, y, eu, ev, weight=None):
    """safe_io.calculate_edge_lengths on the device: sqrt(dx*dx + dy*dy) * weight per edge (NaN for zero weights)."""
    x = _as(x, np.float64)
    y = _as(y, np.float64)
    eu = _as(eu, np.int32)
    ev = _as(ev, np.int32)
    w = None if weight is None else _as(weight, np.float64)
    out = np.empty(eu.shape[0], dtype=np.float64)
    _check(ctx.lib, ctx.lib.sb_graph_edge_lengths(ctx.h, x.shape[0], _ptr(x), _ptr(y), eu.shape[0], _ptr(eu), _ptr(ev),
                                                  _ptr(w), _ptr(out)))
    return out


def build_csr(ctx, n, eu, ev, value=None):
    """Symmetric CSR (indptr int64, indices int32 ascending per row, values) of an undirected edge list."""
    eu = _as(eu, np.int32)
    ev = _as(ev, np.int32)
    ne = eu.shape[0]
    val = None if value is None else _as(value, np.float64)
    indptr = np.empty(n + 1, dtype=np.int64)
    indices = np.empty(max(2 * ne, 1), dtype=np.int32)
    vout = None if value is None else np.empty(max(2 * ne, 1), dtype=np.float64)
    nnz = _i64()
    _check(ctx.lib, ctx.lib.sb_graph_csr(ctx.h, int(n), ne, _ptr(eu), _ptr(ev), _ptr(val), _ptr(indptr), _ptr(indices),
                                         _ptr(vout), C.byref(nnz)))
    k = nnz.value
    return indptr, indices[:k].copy(), (None if vout is None else vout[:k].copy())


def components(ctx, indptr, indices, member, candidates, min_size, want_labels=False):
    """Connected components of the subgraphs induced by member[:, j] for j in candidates.
    Returns (num_components, num_large_components, labels or None)."""
    indptr = _as(indptr, np.int64)
    indices = _as(indices, np.int32)
    member = np.ascontiguousarray(np.asarray(member) != 0, dtype=np.uint8)
    n, m = member.shape
    cand = _as(candidates, np.int32)
    k = cand.shape[0]
    ncc = np.zeros(k, dtype=np.int32)
    nlarge = np.zeros(k, dtype=np.int32)
    labels = np.empty((k, n), dtype=np.int32) if want_labels else None
    _check(ctx.lib, ctx.lib.sb_graph_components(ctx.h, n, _ptr(indptr), _ptr(indices), _ptr(member), m, _ptr(cand), k,
                                                int(min_size), _ptr(labels), _ptr(ncc), _ptr(nlarge)))
    return ncc, nlarge, labels


def fdr_rows(ctx, pvalues):
    """Benjamini-Hochberg adjustment of every row across its entries (statsmodels fdrcorrection, method 'indep')."""
    p = _as(pvalues, np.float64)
    if p.ndim != 2:
        raise ValueError("pvalues must be 2-D")
    out = np.empty_like(p)
    if p.size:
        _check(ctx.lib, ctx.lib.sb_fdr_rows(ctx.h, p.shape[0], p.shape[1], _ptr(p), _ptr(out)))
    return out


def jaccard(ctx, member, columns):
    """Condensed Jaccard distance matrix (scipy pdist order) between the given columns of a 0/1 [n, m] matrix."""
    member = np.ascontiguousarray(np.asarray(member) != 0, dtype=np.uint8)
    n, m = member.shape
    cols = _as(columns, np.int32)
    k = cols.shape[0]
    out = np.empty(k * (k - 1) // 2, dtype=np.float64)
    _check(ctx.lib, ctx.lib.sb_attr_jaccard(ctx.h, n, _ptr(member), m, _ptr(cols), k, _ptr(out)))
    return out


def selftest_mma_i8(ctx, a, b, variant=0):
    """128 x K int8 (values 0/1) times K x ncols int8 through the production tcgen05 kernel -> int32."""
    a = _as(a, np.int8)
    b = _as(b, np.int8)
    k = a.shape[1]
    ncols = b.shape[1]
    if a.shape[0] != 128 or k % 64 or b.shape[0] != k:
        raise ValueError("selftest operands must be 128 x 64k and 64k x ncols")
    d = np.empty((128, ncols), dtype=np.int32)
    _check(ctx.lib, ctx.lib.sb_selftest_mma_i8(ctx.h, ncols, k // 64, int(variant), _ptr(a), _ptr(b), _ptr(d)))
    return d


def selftest_mma_rate(ctx, ncols=192, ktiles=32, slots=64, grid=148, dbg=0, desc=None):
    """Device ms for grid x slots accumulations of ktiles L2-resident k-tiles (128 x ncols x 64 int8 each)."""
    ms = C.c_double()
    d = None if desc is None else np.ascontiguousarray(desc, dtype=np.uint32)
    _check(ctx.lib, ctx.lib.sb_selftest_mma_rate(ctx.h, ncols, ktiles, slots, grid, dbg, _ptr(d), C.byref(ms)))
    return ms.value
