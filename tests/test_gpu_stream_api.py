"""GPU tests of the per-symbol ("continued") API and of the C++ mirror of the
reference's Encoder / Decoder / PModel types (include/rcb200.hpp).

These read like the reference's only test, the round trip at the end of
examples/sample_impl.rs:72-128, plus byte-exact comparison with the oracle.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class StreamState(ctypes.Structure):
    _fields_ = [("lower_bound", ctypes.c_uint64), ("range", ctypes.c_uint64), ("data", ctypes.c_uint64),
                ("consumed", ctypes.c_uint64), ("status", ctypes.c_uint32), ("pad", ctypes.c_uint32)]


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_cpp_sample_impl_example():
    """examples/sample_impl.cpp == examples/sample_impl.rs on the C++ mirror; prints what the Rust
    example prints and checks the known answer + the bulk path."""
    subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], check=True, capture_output=True)
    res = subprocess.run([os.path.join(ROOT, "examples", "sample_impl")], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "index:1, c:5, cum:1" in res.stdout
    assert "output : 0x64475f8970365a2f83b2246c0" in res.stdout  # "{:x}" without zero padding, as in Rust
    assert "length : 13byte" in res.stdout
    assert "decode : 2,1,1,4,1,4,2,1,0,1,5,9,8,7,6,5," in res.stdout
    assert "test passed" in res.stdout


def test_cpp_multi_gpu_example():
    """examples/multi_gpu.cpp: gpu::MultiGpu (rcb_comm_init_all + rcb_allreduce_counts_multi behind the
    C++ mirror) on every visible GPU -- one is enough to run the whole code path."""
    subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], check=True, capture_output=True)
    res = subprocess.run([os.path.join(ROOT, "examples", "multi_gpu")], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "test passed" in res.stdout and "==" in res.stdout


def test_encoder_decoder_symbol_by_symbol(ctx, oracle):
    """Encoder::encode returns the bytes each symbol produced; the concatenation + finish() equals
    the oracle's Encoder run; Decoder::decode returns the symbols one at a time."""
    lib = ctx.lib
    rng = np.random.default_rng(21)
    c = np.array([1, 5, 2, 0, 2, 2, 1, 1, 1, 1], dtype=np.uint32)
    cum, total = oracle.calc_cum(c)
    model = ctx.model_from_tables(c, cum, total)
    used = np.flatnonzero(c)
    syms = rng.choice(used, size=400).astype(np.uint8)
    ref = oracle.encode(syms, c, cum, total)

    st = StreamState()
    lib.rcb_stream_state_init(ctypes.byref(st))
    assert st.lower_bound == 0 and st.range == 2 ** 64 - 1  # RangeCoder::new, src/range_coder.rs:13-20
    code = bytearray()
    out = np.zeros(64, dtype=np.uint8)
    n_out = ctypes.c_uint64()
    per = np.zeros(1, dtype=np.uint32)
    for s in syms[:50]:  # one call per symbol, like the loop at examples/sample_impl.rs:94-97
        one = np.array([s], dtype=np.uint8)
        rc = lib.rcb_encode_stream(ctx.h, ctypes.byref(st), _p(one), 1, 1, model.h, _p(out), out.size,
                                   ctypes.byref(n_out), _p(per), 0)
        assert rc == 0 and per[0] == n_out.value
        code += out[:n_out.value].tobytes()
    big = np.zeros(4096, dtype=np.uint8)
    pers = np.zeros(syms.size - 50, dtype=np.uint32)
    rc = lib.rcb_encode_stream(ctx.h, ctypes.byref(st), _p(syms[50:].copy()), syms.size - 50, 1, model.h, _p(big),
                               big.size, ctypes.byref(n_out), _p(pers), 1)  # rest of the symbols + finish()
    assert rc == 0 and int(pers.sum()) + 8 == n_out.value
    code += big[:n_out.value].tobytes()
    assert bytes(code) == ref

    st = StreamState()
    lib.rcb_stream_state_init(ctypes.byref(st))
    codea = np.frombuffer(ref, dtype=np.uint8).copy()
    got = []
    one = np.zeros(1, dtype=np.uint8)
    for _ in range(30):  # Decoder::new on the first call, then one decode() per call
        rc = lib.rcb_decode_stream(ctx.h, ctypes.byref(st), _p(codea), codea.size, 1, 1, model.h, _p(one))
        assert rc == 0
        got.append(int(one[0]))
    rest = np.zeros(syms.size - 30, dtype=np.uint8)
    rc = lib.rcb_decode_stream(ctx.h, ctypes.byref(st), _p(codea), codea.size, rest.size, 1, model.h, _p(rest))
    assert rc == 0
    assert got + rest.tolist() == syms.tolist()
    assert st.consumed == len(ref)  # 8 + sum(n): exactly the encoder's output
    # one more symbol than was encoded: the reference panics in pop_front (src/decoder.rs:33)
    rc = lib.rcb_decode_stream(ctx.h, ctypes.byref(st), _p(codea), codea.size, 40, 1, model.h,
                               _p(np.zeros(40, np.uint8)))
    assert rc == -9

    # error behaviour: zero-frequency symbol (the reference would never return), bad index
    st = StreamState()
    lib.rcb_stream_state_init(ctypes.byref(st))
    rc = lib.rcb_encode_stream(ctx.h, ctypes.byref(st), _p(np.array([3], np.uint8)), 1, 1, model.h, _p(out), out.size,
                               ctypes.byref(n_out), None, 0)
    assert rc == -4
    rc = lib.rcb_encode_stream(ctx.h, ctypes.byref(st), _p(np.array([10], np.uint8)), 1, 1, model.h, _p(out),
                               out.size, ctypes.byref(n_out), None, 0)
    assert rc == -7
