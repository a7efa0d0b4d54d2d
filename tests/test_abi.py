"""The C-ABI shared library builds for sm_100a, loads, and exports exactly the
entry points include/rcb200.h declares (no compute calls: no GPU needed)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rcb200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rcb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from range_coder_rust_b200 import _lib

    lib = _lib.load()
    names = declared_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in rcb200.h but not exported"
    # the binding table and the header must not drift apart
    assert sorted(_lib.SIGNATURES) == names


def test_library_is_sm100a_only():
    from range_coder_rust_b200 import _lib

    _lib.load()
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_device():
    """Without a CUDA device every entry point refuses to work instead of falling back."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from range_coder_rust_b200 import _lib

    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.rcb_ctx_create(0, None, ctypes.byref(h))
    assert rc == _lib.RCB_ERR_NO_DEVICE and not h.value
    assert b"no CPU fallback" in lib.rcb_strerror(rc)
    import range_coder_rust_b200 as rcb

    with pytest.raises(rcb.RcbError):
        rcb.Context(0)


def test_product_does_not_reference_the_oracle():
    """oracle/ is test infrastructure: nothing under the package may import, link or call it."""
    pkg = os.path.join(ROOT, "range_coder_rust_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "rc_oracle" not in text and "oracle_bind" not in text and "rc_pyref" not in text, f
    from range_coder_rust_b200 import _lib

    _lib.load()
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "rc_oracle" not in out


def test_version_and_strerror():
    from range_coder_rust_b200 import _lib

    lib = _lib.load()
    assert lib.rcb_version() == 100
    assert lib.rcb_strerror(0) == b"ok"
    assert b"LowerBoundOverflow" in lib.rcb_strerror(_lib.RCB_ERR_LOWER_OVERFLOW)


def test_rust_ffi_declarations_match_header():
    """rust/range_coder_gpu/src/ffi.rs cannot be compiled here (no Rust toolchain): at least keep its
    extern block in step with include/rcb200.h -- same names, same parameter counts."""
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "rcb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    rs = open(os.path.join(root, "rust", "range_coder_gpu", "src", "ffi.rs")).read()
    decls = re.findall(r"pub fn (rcb_\w+)\s*\((.*?)\)\s*(?:->\s*[\w:* ]+)?;", rs, flags=re.S)
    assert len(decls) >= 12
    for name, params in decls:
        m = re.search(r"\b%s\s*\((.*?)\)\s*;" % name, hdr, flags=re.S)
        assert m, f"{name} not declared in rcb200.h"
        n_c = 0 if m.group(1).strip() in ("", "void") else m.group(1).count(",") + 1
        n_rs = len([p for p in params.split(",") if p.strip()])
        assert n_c == n_rs, f"{name}: header has {n_c} parameters, ffi.rs {n_rs}"
