"""Shared fixtures.  GPU tests are marked `gpu`; everything else runs on the CPU-only box."""
import ctypes
import os
import sys

import numpy as np
import pytest

REPO = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def built():
    import __graft_entry__ as entry
    entry.build()
    return True


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """Without a CUDA device the `gpu` tests are skipped, not failed (the product itself refuses loudly: test_abi)."""
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def devcheck(built):
    """Test-only host compile of the device functions (tests/devcheck/devcheck.cpp)."""
    from ctypes import POINTER, c_int, c_uint32, c_void_p, c_float
    from oracle import oracle
    lib = ctypes.CDLL(os.path.join(REPO, "tests", "devcheck", "_build", "libgdpt_devcheck.so"))
    lib.devcheck_path_trace.restype = c_int
    lib.devcheck_path_trace.argtypes = [POINTER(oracle.OrcScene), c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                        c_void_p, c_void_p, c_int, c_void_p, c_uint32, c_void_p]
    lib.devcheck_rng.argtypes = [c_uint32, c_uint32, c_uint32, c_void_p, c_void_p, c_void_p]
    lib.devcheck_sincos.argtypes = [c_float, c_void_p]
    lib.devcheck_progressive.argtypes = [c_void_p, c_void_p, c_int, c_int, c_uint32]
    lib.devcheck_temporal.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    return lib


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def records_equal(a, b):
    """Bitwise comparison of two trace-record arrays where `b` (the oracle) marks unreached
    segments with hit == 0xFFFFFFFF; those slots only need the same marker."""
    if not np.array_equal(a["hit"], b["hit"]):
        return False, "hit flags differ"
    live = b["hit"] != 0xFFFFFFFF
    for f in a.dtype.names:
        x, y = a[f][live], b[f][live]
        if x.dtype == np.float32:
            x, y = x.view(np.uint32), y.view(np.uint32)
        if not np.array_equal(x, y):
            return False, f"field {f}: {(x != y).sum()} of {live.sum()} records differ"
    return True, ""
