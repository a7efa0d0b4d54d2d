"""RCB2 container: the checker-side writer/reader on CPU, and -- on a GPU --
frames written by the library decoded by the oracle and the reverse, byte for byte."""
import ctypes
import os

import numpy as np
import pytest

import frame_ref


def _data(oracle, n, K, s=1.1, seed=0x5EED0001):
    return oracle.generate(n, K, seed, oracle.zipf_thresholds(K, s), sym_bytes=1 if K <= 256 else 2)


def test_checker_frame_round_trip(oracle):
    syms = _data(oracle, 5 * 1000 + 123, 256)
    m = oracle.model_from_symbols(syms, 256)
    frame = frame_ref.write_frame(syms, 1000, [m], 256)
    f = frame_ref.read_frame(frame)
    assert f["n"] == syms.size and f["n_chunks"] == 6 and f["mode"] == 0
    assert np.array_equal(frame_ref.decode_frame(frame), syms)
    per = [oracle.model_from_symbols(syms[j * 1000:(j + 1) * 1000], 256) for j in range(6)]
    frame2 = frame_ref.write_frame(syms, 1000, per, 256)
    assert np.array_equal(frame_ref.decode_frame(frame2), syms)


def test_frame_parse_rejects_garbage():
    # header validation is host-only code in the library: callable without a GPU
    from range_coder_rust_b200 import _lib
    from range_coder_rust_b200.api import FrameInfo

    lib = _lib.load()
    info = FrameInfo()
    bad = np.zeros(64, dtype=np.uint8)
    assert lib.rcb_frame_parse(bad.ctypes.data_as(ctypes.c_void_p), bad.size, ctypes.byref(info)) != 0
    assert lib.rcb_frame_parse(bad.ctypes.data_as(ctypes.c_void_p), 10, ctypes.byref(info)) != 0


def test_frame_parse_accepts_checker_frame(oracle):
    from range_coder_rust_b200 import _lib
    from range_coder_rust_b200.api import FrameInfo

    lib = _lib.load()
    syms = _data(oracle, 4321, 256)
    frame = frame_ref.write_frame(syms, 1024, [oracle.model_from_symbols(syms, 256)], 256)
    info = FrameInfo()
    assert lib.rcb_frame_parse(frame.ctypes.data_as(ctypes.c_void_p), frame.size, ctypes.byref(info)) == 0
    assert (info.K, info.sym_bytes, info.n_syms, info.n_chunks, info.chunk_syms) == (256, 1, 4321, 5, 1024)
    assert info.frame_bytes == frame.size
    assert lib.rcb_frame_bound(256, 5, 0, info.payload_bytes) == frame.size
    # truncated by one byte
    assert lib.rcb_frame_parse(frame.ctypes.data_as(ctypes.c_void_p), frame.size - 1, ctypes.byref(info)) != 0
    # non-monotone offsets
    broken = frame.copy()
    broken[info.offsets_off + 8:info.offsets_off + 16].view("<u8")[0] = info.payload_bytes + 1
    assert lib.rcb_frame_parse(broken.ctypes.data_as(ctypes.c_void_p), broken.size, ctypes.byref(info)) != 0


def test_frame_parse_rejects_crafted_headers(oracle):
    """Untrusted frames: sizes that wrap u64 arithmetic (n_syms * sym_bytes == 0 mod 2^64), a chunk size
    above the coder's limit, and a chunk shorter than the 8 bytes of Encoder::finish must not parse."""
    import struct

    from range_coder_rust_b200 import _lib
    from range_coder_rust_b200.api import FrameInfo

    lib = _lib.load()
    info = FrameInfo()

    def parse(fr):
        return lib.rcb_frame_parse(fr.ctypes.data_as(ctypes.c_void_p), fr.size, ctypes.byref(info))

    syms = _data(oracle, 4000, 256)
    frame = frame_ref.write_frame(syms, 1000, [oracle.model_from_symbols(syms, 256)], 256)
    assert parse(frame) == 0
    good = FrameInfo.from_buffer_copy(bytes(info))
    # header fields: chunk_syms @24, n_syms @32, n_chunks @40 (include/rcb200.h)
    crafted = frame.copy()
    crafted[8:12] = np.frombuffer(struct.pack("<I", 2), dtype=np.uint8)            # sym_bytes = 2
    crafted[24:48] = np.frombuffer(struct.pack("<QQQ", 1 << 63, 1 << 63, 1), dtype=np.uint8)
    assert parse(crafted) != 0
    crafted = frame.copy()
    crafted[24:48] = np.frombuffer(struct.pack("<QQQ", (1 << 30) + 1, 4000, 1), dtype=np.uint8)
    assert parse(crafted) != 0
    # a zero-length chunk / a chunk of 7 bytes
    for delta in (0, 7):
        broken = frame.copy()
        offs = broken[good.offsets_off:good.offsets_off + 8 * 5].view("<u8")
        offs[2] = offs[1] + delta
        assert parse(broken) != 0
    # offsets not starting at zero
    broken = frame.copy()
    broken[good.offsets_off:good.offsets_off + 8].view("<u8")[0] = 1
    assert parse(broken) != 0


@pytest.mark.gpu
@pytest.mark.parametrize("K,chunk,n", [(256, 4096, 10 * 4096 + 77), (256, 65536, 3 * 65536), (4096, 2048, 50000),
                                       (2, 100, 1001)])
def test_gpu_frame_equals_checker_frame_shared_model(ctx, oracle, K, chunk, n):
    syms = _data(oracle, n, K)
    c, cum, total = oracle.model_from_symbols(syms, K)
    model = ctx.model_from_tables(c, cum, total)
    frame = ctx.frame_encode(syms, chunk, model)
    ref = frame_ref.write_frame(syms, chunk, [(c, cum, total)], K)
    assert frame.tobytes() == ref.tobytes()
    # GPU frame -> oracle, oracle frame -> GPU
    assert np.array_equal(frame_ref.decode_frame(frame), syms)
    assert np.array_equal(ctx.frame_decode(ref), syms)
    info = ctx.frame_info(frame)
    assert info.n_syms == n and info.K == K and info.model_mode == 0


@pytest.mark.gpu
def test_gpu_frame_per_chunk_models(ctx, oracle):
    import torch

    K, chunk, n = 256, 8192, 9 * 8192 + 5
    thr = np.stack([oracle.zipf_thresholds(K, s) for s in (0.0, 1.1, 3.0)])
    syms = oracle.generate(n, K, 0x5EED0002, thr, chunk_syms=chunk)
    d = torch.from_numpy(syms).to(ctx.device)
    model = ctx.model_from_counts(ctx.histogram(d, K, chunk_syms=chunk))
    frame = ctx.frame_encode(syms, chunk, model)
    per = [oracle.model_from_symbols(syms[j * chunk:(j + 1) * chunk], K) for j in range(10)]
    ref = frame_ref.write_frame(syms, chunk, per, K)
    assert frame.tobytes() == ref.tobytes()
    assert np.array_equal(ctx.frame_decode(ref), syms)
    assert np.array_equal(frame_ref.decode_frame(frame), syms)


@pytest.mark.gpu
@pytest.mark.parametrize("K,chunk,n,rs,per_chunk", [(256, 65536, 5 * 65536 + 777, 16384, False),
                                                    (256, 8192, 20 * 8192 + 5, 2048, True),
                                                    (4096, 32768, 3 * 32768 + 100, 8192, False),
                                                    (256, 65536, 40 * 65536, 8192, False)])
def test_gpu_frame_with_restart_points(ctx, oracle, K, chunk, n, rs, per_chunk):
    """Version-2 frames: the restart section (every record from the oracle's Encoder state) makes the GPU frame
    byte-identical to the checker's; either side's frame decodes on the other; the 40-chunk case runs the sliced
    host pipeline."""
    import torch

    if per_chunk:
        thr = np.stack([oracle.zipf_thresholds(K, s) for s in (0.0, 1.1, 3.0)])
        syms = oracle.generate(n, K, 0x5EED0002, thr, chunk_syms=chunk)
        d = torch.from_numpy(syms).to(ctx.device)
        model = ctx.model_from_counts(ctx.histogram(d, K, chunk_syms=chunk))
        models = [oracle.model_from_symbols(syms[j * chunk:(j + 1) * chunk], K) for j in range((n + chunk - 1) // chunk)]
    else:
        syms = _data(oracle, n, K)
        c, cum, total = oracle.model_from_symbols(syms, K)
        model = ctx.model_from_tables(c, cum, total)
        models = [(c, cum, total)]
    frame = ctx.frame_encode(syms, chunk, model, restart_syms=rs)
    info = ctx.frame_info(frame)
    assert info.version == 2 and info.restart_syms == rs and info.frame_bytes == frame.size
    ref = frame_ref.write_frame(syms, chunk, models, K, restart_syms=rs)
    assert frame.tobytes() == ref.tobytes()
    assert np.array_equal(ctx.frame_decode(ref), syms)
    assert np.array_equal(frame_ref.decode_frame(frame), syms)
    # the same symbols without restart points: a version-1 frame with the same payload
    plain = ctx.frame_encode(syms, chunk, model)
    f1, f2 = frame_ref.read_frame(plain), frame_ref.read_frame(frame)
    assert ctx.frame_info(plain).version == 1 and np.array_equal(f1["stream"], f2["stream"])
    assert np.array_equal(f1["offsets"], f2["offsets"])
    # a damaged record is reported (RCB_ERR_RESTART_POINT), not silently decoded
    bad = frame.copy()
    bad[info.restart_off] ^= 0x40
    from range_coder_rust_b200 import RcbError
    with pytest.raises(RcbError) as e:
        ctx.frame_decode(bad)
    assert e.value.code == -14


@pytest.mark.gpu
def test_gpu_frame_empty_and_capacity(ctx, oracle):
    c = np.array([3, 1], dtype=np.uint32)
    cum = np.array([0, 3], dtype=np.uint32)
    model = ctx.model_from_tables(c, cum, 4)
    frame = ctx.frame_encode(np.zeros(0, dtype=np.uint8), 16, model)
    assert ctx.frame_info(frame).n_chunks == 0 and ctx.frame_decode(frame).size == 0
    syms = np.array([0, 1, 0, 0, 1, 0, 0], dtype=np.uint8)
    frame = ctx.frame_encode(syms, 4, model)
    out = np.empty(3, dtype=np.uint8)
    n = ctypes.c_uint64()
    rc = ctx.lib.rcb_frame_decode_host(ctx.h, frame.ctypes.data_as(ctypes.c_void_p), frame.size,
                                       out.ctypes.data_as(ctypes.c_void_p), out.nbytes, ctypes.byref(n))
    assert rc == -8  # RCB_ERR_OUT_CAPACITY
    assert n.value == 7


@pytest.mark.gpu
@pytest.mark.parametrize("adaptive", [False, True])
def test_file_front_end_round_trip(tmp_path, oracle, adaptive):
    """python -m range_coder_rust_b200 compress / decompress: the frame on disk decodes with the checker too."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    data = _data(oracle, 300_007, 256, s=1.3, seed=0x5EED0007)
    src, frm, dst = tmp_path / "in.bin", tmp_path / "in.rcb2", tmp_path / "out.bin"
    data.tofile(src)
    cmd = [sys.executable, "-m", "range_coder_rust_b200"]
    subprocess.run(cmd + ["compress", str(src), str(frm), "--chunk", "32768"] + (["--adaptive"] if adaptive else []),
                   check=True, cwd=root)
    subprocess.run(cmd + ["decompress", str(frm), str(dst)], check=True, cwd=root)
    assert np.array_equal(np.fromfile(dst, dtype=np.uint8), data)
    assert np.array_equal(frame_ref.decode_frame(np.fromfile(frm, dtype=np.uint8)), data)
    out = subprocess.run(cmd + ["info", str(frm)], check=True, cwd=root, capture_output=True, text=True).stdout
    assert "300007 symbols" in out and ("per-chunk tables" in out) == adaptive
    assert "RCB2 v2" in out and "restart points every 2048 symbols" in out  # the front end's default: chunk / 16
