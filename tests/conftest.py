import ctypes
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
TESTS = os.path.dirname(os.path.abspath(__file__))
if TESTS not in sys.path:
    sys.path.insert(0, TESTS)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _have_cuda():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a CPU-only box skips the gpu-marked tests (some of them only start
    subprocesses and would otherwise fail with RCB_ERR_NO_DEVICE instead of skipping)."""
    if _have_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    import oracle_bind

    oracle_bind.lib()
    return oracle_bind


@pytest.fixture(scope="session")
def hostcore():
    """g++ build of the host instantiation of rcb_core.cuh (test-only harness)."""
    src = os.path.join(TESTS, "hostcore", "hostcore.cpp")
    core = os.path.join(ROOT, "range_coder_rust_b200", "csrc", "rcb_core.cuh")
    so = os.path.join(TESTS, "hostcore", "libhostcore.so")
    if not os.path.exists(so) or max(os.path.getmtime(src), os.path.getmtime(core)) > os.path.getmtime(so):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", src, "-o", so],
                       check=True, capture_output=True)
    L = ctypes.CDLL(so)
    vp, u64, u32, ci = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int
    L.hc_encode.restype = ctypes.c_int64
    L.hc_encode.argtypes = [vp, u64, ci, u32, vp, vp, u32, vp, u32, ci, vp]
    L.hc_decode.restype = ctypes.c_int64
    L.hc_decode.argtypes = [vp, u64, u64, u64, u64, ci, u32, vp, vp, u32, vp, ci, ci, u32, vp, vp]
    L.hc_check_division.restype = u64
    L.hc_check_division.argtypes = [vp, u64, u32]
    L.hc_check_cs.restype = u64
    L.hc_check_cs.argtypes = [vp, vp, vp, u64, u32, vp]
    L.hc_check_m2.restype = u64
    L.hc_check_m2.argtypes = [vp, vp, u64, u32, vp]
    return L


@pytest.fixture(scope="session")
def ctx():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import range_coder_rust_b200 as rcb

    c = rcb.Context(0)
    yield c
    c.close()
