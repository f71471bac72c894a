"""CPU: the C-ABI library builds, loads and exports exactly the symbols include/safe_b200.h declares
(no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import pytest

from safepy_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "safe_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load_library()


def test_header_and_binding_agree():
    assert declared_functions() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol(lib):
    for name in declared_functions():
        assert hasattr(lib, name), name
    assert lib.sb_abi_version() == 1


def test_pure_helpers_without_gpu(lib):
    for n in (1, 31, 32, 33, 127, 128, 129, 100000):
        assert lib.sb_neigh_ld(n) == _lib.neigh_ld(n)
        assert lib.sb_neigh_ld(n) % 4 == 0 and lib.sb_neigh_ld(n) * 32 >= n


def test_fails_loudly_without_device(lib):
    """No CPU fallback: without a B200 the context cannot be created and says why."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    assert lib.sb_ctx_create(-1, ctypes.byref(h)) != 0
    assert b"no CPU fallback" in lib.sb_last_error()
    with pytest.raises(_lib.SafeB200Error):
        _lib.Context()


def test_missing_library_is_an_error(tmp_path):
    with pytest.raises(_lib.SafeB200Error):
        _lib.load_library(str(tmp_path / "nope.so"))
