"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU
oracle on the same inputs -- bit-exact bytes per chunk, bit-exact symbols back.

Modelled on the reference's only check, the round trip of
examples/sample_impl.rs:72-128, widened to the configurations of BASELINE.json.
"""
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VECTORS = json.load(open(os.path.join(ROOT, "tests", "golden", "vectors.json")))


def to_dev(ctx, a):
    t = torch.from_numpy(np.ascontiguousarray(a).view(np.int16) if a.dtype == np.uint16 else np.ascontiguousarray(a))
    return t.to(ctx.device)


def dev_to_np(t, n=None, dtype=None):
    a = t.cpu().numpy()
    if dtype is not None:
        a = a.view(dtype)
    return a if n is None else a[:n]


def gpu_encode(ctx, syms_np, chunk, model):
    d_syms = to_dev(ctx, syms_np)
    stream, offsets, nbytes = ctx.encode_chunks(d_syms, chunk, model)
    return stream, offsets, nbytes, d_syms


def assert_streams_equal(stream_t, offsets_t, nbytes, ref_stream, ref_offsets):
    offs = dev_to_np(offsets_t).astype(np.uint64)
    assert np.array_equal(offs, ref_offsets), "chunk offsets differ from the oracle"
    assert nbytes == int(ref_offsets[-1])
    got = dev_to_np(stream_t, nbytes)
    if not np.array_equal(got, ref_stream):
        bad = int(np.flatnonzero(got != ref_stream)[0])
        chunk = int(np.searchsorted(ref_offsets, bad, side="right") - 1)
        raise AssertionError(f"first differing byte {bad} (chunk {chunk})")


def vector_inputs(v):
    if "symbols" in v:
        syms = np.array(v["symbols"], dtype=np.uint16 if v["K"] > 256 else np.uint8)
    elif "symbols_repeat" in v:
        syms = np.full(v["symbols_repeat"][1], v["symbols_repeat"][0], dtype=np.uint8)
    else:
        syms = np.frombuffer(bytes.fromhex(v["symbols_hex"]), dtype="<u2" if v["K"] > 256 else np.uint8).copy()
    if v.get("c_from_symbols"):
        c = np.bincount(syms, minlength=v["K"]).astype(np.uint32)
        cum = np.concatenate([[0], np.cumsum(c)[:-1]]).astype(np.uint32)
    else:
        c = np.array(v["c"], dtype=np.uint32)
        cum = np.array(v["cum"], dtype=np.uint32)
    return syms, c, cum, v["total"]


# ------------------------------------------------------------- config 1: KATs
@pytest.mark.parametrize("v", VECTORS, ids=[v["name"] for v in VECTORS])
def test_golden_vectors_on_gpu(ctx, v):
    syms, c, cum, total = vector_inputs(v)
    model = ctx.model_from_tables(c, cum, total)
    n = syms.size
    chunk = max(n, 1)
    d_syms = to_dev(ctx, syms) if n else torch.empty(0, dtype=torch.uint8, device=ctx.device)
    if n == 0:
        # zero symbols -> zero chunks; a single empty Encoder run is a chunk of 0 symbols
        stream, offsets, nbytes = ctx.encode_chunks(d_syms, 1, model)
        assert nbytes == 0
        return
    stream, offsets, nbytes = ctx.encode_chunks(d_syms, chunk, model)
    code = dev_to_np(stream, nbytes).tobytes()
    assert nbytes == v["code_len"]
    assert hashlib.sha256(code).hexdigest() == v["code_sha256"]
    if "code_hex" in v:
        assert code.hex() == v["code_hex"]
    out = ctx.decode_chunks(stream, offsets, n, chunk, model, sym_bytes=syms.dtype.itemsize)
    assert np.array_equal(dev_to_np(out, dtype=syms.dtype), syms)


def test_sample_impl_round_trip(ctx, oracle):
    """examples/sample_impl.rs as shipped, with the table built on the GPU as well."""
    data = np.array([2, 1, 1, 4, 1, 4, 2, 1, 0, 1, 5, 9, 8, 7, 6, 5], dtype=np.uint8)
    d = to_dev(ctx, data)
    counts = ctx.histogram(d, 10)
    assert counts.cpu().tolist() == [1, 5, 2, 0, 2, 2, 1, 1, 1, 1]
    model = ctx.model_from_counts(counts)
    c, cum, total, flags = model.tables()
    assert total == 16 and list(cum) == [0, 1, 6, 8, 8, 10, 12, 13, 14, 15]
    stream, offsets, nbytes = ctx.encode_chunks(d, 16, model)
    assert dev_to_np(stream, nbytes).tobytes().hex() == "64475f8970365a2f83b20246c0"
    out = ctx.decode_chunks(stream, offsets, 16, 16, model)
    assert np.array_equal(dev_to_np(out), data)


# ------------------------------------- config 2: static global table, Zipf(1.1)
@pytest.mark.parametrize("n,chunk", [(8 << 20, 65536), (1_000_003, 65536), (300_000, 4096), (5000, 17),
                                     (777, 1), (100_000, 100_000)])
def test_static_zipf_matches_oracle(ctx, oracle, n, chunk):
    thr = oracle.zipf_thresholds(256, 1.1)
    syms = oracle.generate(n, 256, 0x5EED0001, thr)
    d_syms = to_dev(ctx, syms)
    counts = ctx.histogram(d_syms, 256)
    ref_counts = oracle.histogram(syms, 256)
    assert np.array_equal(dev_to_np(counts).astype(np.uint64), ref_counts)
    model = ctx.model_from_counts(counts)
    c, cum, total, flags = model.tables()
    rc, rcum, rtotal = oracle.model_from_symbols(syms, 256)
    assert total == rtotal and np.array_equal(c, rc) and np.array_equal(cum, rcum)
    stream, offsets, nbytes = ctx.encode_chunks(d_syms, chunk, model)
    ref_stream, ref_offsets = oracle.encode_chunks(syms, chunk, rc, rcum, rtotal)
    assert_streams_equal(stream, offsets, nbytes, ref_stream, ref_offsets)
    out = ctx.decode_chunks(stream, offsets, n, chunk, model)
    assert np.array_equal(dev_to_np(out), syms)
    # cross decode: the oracle's stream decoded on the GPU, the GPU's stream by the oracle
    pad = np.zeros(((ref_stream.size + 31) // 16) * 16, dtype=np.uint8)
    pad[:ref_stream.size] = ref_stream
    out2 = ctx.decode_chunks(to_dev(ctx, pad), to_dev(ctx, ref_offsets.view(np.int64)), n, chunk, model)
    assert np.array_equal(dev_to_np(out2), syms)
    dec, _ = oracle.decode_chunks(dev_to_np(stream, nbytes), ref_offsets, n, chunk, rc, rcum, rtotal)
    assert np.array_equal(dec, syms)


def test_generator_matches_oracle(ctx, oracle):
    for K, sb, s in [(256, 1, 1.1), (4096, 2, 1.1), (10, 1, 0.0)]:
        thr = oracle.zipf_thresholds(K, s)
        ref = oracle.generate(200_001, K, 0x5EED0003, thr, sym_bytes=sb, first=12345)
        got = ctx.generate(200_001, K, 0x5EED0003, thr, sym_bytes=sb, first=12345)
        assert np.array_equal(dev_to_np(got, dtype=ref.dtype), ref)
    # several tables cycling per chunk (config 3's mixed-entropy input)
    thr = np.stack([oracle.zipf_thresholds(256, s) for s in (0.0, 0.5, 1.1, 3.0)])
    ref = oracle.generate(70_000, 256, 0x5EED0002, thr, chunk_syms=4096)
    got = ctx.generate(70_000, 256, 0x5EED0002, thr, chunk_syms=4096)
    assert np.array_equal(dev_to_np(got), ref)


def test_histogram_small_alphabets_and_range_check(ctx, oracle):
    """Byte histogram kernel (replicated bins): K < 256 takes the range-checked flavour; vector loop,
    two-loads-per-trip loop and the scalar tail are all exercised; a symbol >= K is reported."""
    from range_coder_rust_b200 import _lib

    rng = np.random.default_rng(11)
    for K, n in ((100, (1 << 20) + 7), (2, 12345), (255, 3 * 4096 + 16)):
        syms = rng.integers(0, K, size=n).astype(np.uint8)
        d = to_dev(ctx, syms)
        counts = ctx.histogram(d, K)
        assert np.array_equal(dev_to_np(counts).astype(np.uint64), oracle.histogram(syms, K))
    syms = rng.integers(0, 100, size=1 << 16).astype(np.uint8)
    syms[40000] = 100  # out of range, inside the vector loop
    with pytest.raises(_lib.RcbError) as e:
        ctx.histogram(to_dev(ctx, syms), 100)
    assert e.value.code == _lib.RCB_ERR_SYMBOL_OUT_OF_RANGE


# ------------------------- config 3: per-chunk adaptive histograms, chunk sweep
S_CYCLE = (0.0, 0.25, 0.5, 0.8, 1.1, 1.5, 2.0, 3.0, 5.0)


@pytest.mark.parametrize("chunk", [16384, 65536, 262144])
def test_adaptive_per_chunk_models(ctx, oracle, chunk):
    n = 4 * 1024 * 1024 + 4321  # ragged last chunk
    thr = np.stack([oracle.zipf_thresholds(256, s) for s in S_CYCLE])
    syms = oracle.generate(n, 256, 0x5EED0002, thr, chunk_syms=chunk)
    d_syms = to_dev(ctx, syms)
    counts = ctx.histogram(d_syms, 256, chunk_syms=chunk)
    n_chunks = (n + chunk - 1) // chunk
    ref_c = np.zeros((n_chunks, 256), dtype=np.uint32)
    ref_cum = np.zeros((n_chunks, 256), dtype=np.uint32)
    ref_total = np.zeros(n_chunks, dtype=np.uint32)
    for i in range(n_chunks):
        ref_c[i], ref_cum[i], ref_total[i] = oracle.model_from_symbols(syms[i * chunk:(i + 1) * chunk], 256)
    assert np.array_equal(dev_to_np(counts).view(np.uint32), ref_c)
    model = ctx.model_from_counts(counts)
    for i in (0, n_chunks // 2, n_chunks - 1):
        c, cum, total, flags = model.tables(i)
        assert total == ref_total[i] and np.array_equal(c, ref_c[i]) and np.array_equal(cum, ref_cum[i])
    stream, offsets, nbytes = ctx.encode_chunks(d_syms, chunk, model)
    ref_stream, ref_offsets = oracle.encode_chunks(syms, chunk, ref_c, ref_cum, ref_total)
    assert_streams_equal(stream, offsets, nbytes, ref_stream, ref_offsets)
    out = ctx.decode_chunks(stream, offsets, n, chunk, model)
    assert np.array_equal(dev_to_np(out), syms)


# ----------------------------------------- config 4: 4096 symbols, u16 storage
def test_k4096_u16(ctx, oracle):
    n, chunk = 3 * 1024 * 1024 + 11, 32768
    thr = oracle.zipf_thresholds(4096, 1.1)
    syms = oracle.generate(n, 4096, 0x5EED0003, thr, sym_bytes=2)
    d_syms = to_dev(ctx, syms)
    counts = ctx.histogram(d_syms, 4096)
    assert np.array_equal(dev_to_np(counts).astype(np.uint64), oracle.histogram(syms, 4096))
    model = ctx.model_from_counts(counts)
    rc, rcum, rtotal = oracle.model_from_symbols(syms, 4096)
    stream, offsets, nbytes = ctx.encode_chunks(d_syms, chunk, model)
    ref_stream, ref_offsets = oracle.encode_chunks(syms, chunk, rc, rcum, rtotal)
    assert_streams_equal(stream, offsets, nbytes, ref_stream, ref_offsets)
    out = ctx.decode_chunks(stream, offsets, n, chunk, model, sym_bytes=2)
    assert np.array_equal(dev_to_np(out, dtype=np.uint16), syms)


# ----------------------------------------------- reciprocal path / odd tables
@pytest.mark.parametrize("kind", ["prime_total", "u32_max", "gaps", "single", "irregular", "slack_total"])
def test_tables_from_host(ctx, oracle, kind):
    rng = np.random.default_rng(11)
    if kind == "prime_total":
        w = np.arange(1, 257, dtype=np.float64) ** -1.1
        c = np.maximum(1, (w / w.sum() * 1_000_003).astype(np.uint32))
        c[0] += 1_000_003 - int(c.sum())
    elif kind == "u32_max":
        c = np.array([1, 0xFFFFFFFF - 3, 2], dtype=np.uint32)
    elif kind == "gaps":
        c = (rng.integers(0, 2, size=200) * rng.integers(1, 90000, size=200)).astype(np.uint32)
        c[17] = 5
    elif kind == "single":
        c = np.array([123457], dtype=np.uint32)
    else:  # unused code space between symbols: legal for the coder, not a prefix-sum table
        c = rng.integers(1, 1000, size=64).astype(np.uint32)
    cum, total = oracle.calc_cum(c)
    if kind == "irregular":
        cum = (cum + np.arange(64, dtype=np.uint32) * 3).astype(np.uint32)
        total = int(cum[-1] + c[-1] + 10)
    if kind == "slack_total":
        # proper prefix sums but unused code space ABOVE the last symbol (legal for a PModel): the
        # kernels that derive c[K-1] as total - cum[K-1] must not be chosen for it
        total = int(total + 11)  # odd: misses the power-of-two paths as well
    used = np.flatnonzero(c)
    n, chunk = 200_000, 8192
    p = c[used].astype(np.float64)
    syms = rng.choice(used, size=n, p=p / p.sum()).astype(np.uint8)
    model = ctx.model_from_tables(c, cum, total)
    _, _, _, flags = model.tables()
    assert bool(flags & 4) == (kind not in ("irregular", "slack_total"))  # RCB_MODEL_REGULAR
    d_syms = to_dev(ctx, syms)
    stream, offsets, nbytes = ctx.encode_chunks(d_syms, chunk, model)
    ref_stream, ref_offsets = oracle.encode_chunks(syms, chunk, c, cum, total)
    assert_streams_equal(stream, offsets, nbytes, ref_stream, ref_offsets)
    out = ctx.decode_chunks(stream, offsets, n, chunk, model)
    assert np.array_equal(dev_to_np(out), syms)


@pytest.mark.parametrize("kind", ["gen_2e30", "gen_prime", "pow2_2e20", "pow2_4096", "gen_small", "gen_u16",
                                  "gen_full_c"])
def test_fused_paths_for_general_totals(ctx, oracle, kind):
    """The word-speculative encoder and the fat-LUT decoder for totals that are not a power of two >= 2^24:
    the divide-free general-total step (reciprocal constant per symbol, exactness test, exact re-code on a
    flagged word), small powers of two (two shifts), and the table the divide-free step must refuse
    (a symbol with c == total).  Bytes against the oracle, symbols back, cross-decode of the oracle's stream."""
    rng = np.random.default_rng(17)
    K, sb = 256, 1
    if kind == "gen_2e30":
        total = (1 << 30) + 12345
    elif kind == "gen_prime":
        total = 1_000_003
    elif kind == "pow2_2e20":
        total = 1 << 20
    elif kind == "pow2_4096":
        K, total = 64, 4096
    elif kind == "gen_small":
        K, total = 50, 3001
    elif kind == "gen_u16":
        K, total, sb = 1000, (1 << 28) + 7, 2
    else:
        K, total = 3, 1_000_003
    if kind == "gen_full_c":
        c = np.array([0, total, 0], dtype=np.uint32)
    else:
        w = np.arange(1, K + 1, dtype=np.float64) ** -0.9
        floor_c = max(2, int(total / 4096 * 1.2) + 2)  # every symbol resolvable by the 4096-bucket table
        c = np.maximum(floor_c, (w / w.sum() * (total - K * floor_c)).astype(np.int64))
        c[0] += total - int(c.sum())
        assert c[0] > 0
        c = c.astype(np.uint32)
    cum, t = oracle.calc_cum(c)
    assert t == total
    n, chunk = 1_500_000 + 777, 65536
    used = np.flatnonzero(c)
    p = c[used].astype(np.float64)
    syms = rng.choice(used, size=n, p=p / p.sum()).astype(np.uint16 if sb == 2 else np.uint8)
    model = ctx.model_from_tables(c, cum, total)
    d_syms = to_dev(ctx, syms)
    stream, offsets, nbytes = ctx.encode_chunks(d_syms, chunk, model)
    ref_stream, ref_offsets = oracle.encode_chunks(syms, chunk, c, cum, total)
    assert_streams_equal(stream, offsets, nbytes, ref_stream, ref_offsets)
    out = ctx.decode_chunks(stream, offsets, n, chunk, model, sym_bytes=sb)
    assert np.array_equal(dev_to_np(out, dtype=syms.dtype), syms)
    pad = np.zeros(((ref_stream.size + 31) // 16) * 16, dtype=np.uint8)
    pad[:ref_stream.size] = ref_stream
    out2 = ctx.decode_chunks(to_dev(ctx, pad), to_dev(ctx, ref_offsets.view(np.int64)), n, chunk, model, sym_bytes=sb)
    assert np.array_equal(dev_to_np(out2, dtype=syms.dtype), syms)


@pytest.mark.parametrize("shape", ["u8_64k", "u8_odd_total", "u16_k4096", "tiny_rows", "five_vectors", "many_warps"])
def test_tma_symbol_staging_matches_oracle(ctx, oracle, shape, monkeypatch):
    """encode_tma_kernel (RCB_ENC_TMA=1): the symbols reach the lanes through cp.async.bulk.tensor.2d boxes
    of {64 bytes x 32 rows} instead of per-lane cp.async pieces; same bytes as the oracle.  Shapes cover a
    partly filled last warp (zero-filled rows), rows of one box column, a row length that is not a multiple
    of the box, u16 symbols and the general-total flavour."""
    monkeypatch.setenv("RCB_ENC_TMA", "1")
    K, sb, chunk, n_chunks, odd = 256, 1, 65536, 70, False
    if shape == "u8_odd_total":
        odd = True
    elif shape == "u16_k4096":
        K, sb, chunk, n_chunks = 4096, 2, 32768, 45
    elif shape == "tiny_rows":
        chunk, n_chunks = 64, 1000
    elif shape == "five_vectors":
        chunk, n_chunks = 80, 333
    elif shape == "many_warps":
        chunk, n_chunks = 4096, 5000
    n = chunk * n_chunks
    syms = oracle.generate(n, K, 0x5EED0001, oracle.zipf_thresholds(K, 1.1), sym_bytes=sb)
    d_syms = to_dev(ctx, syms)
    counts = ctx.histogram(d_syms, K)
    counts += 1  # every symbol codable whatever the sample missed
    if shape in ("tiny_rows", "five_vectors", "many_warps") or odd:
        counts[0] += (1 << 26) + 12345  # a general total >= 2^25: the table-wide reciprocal flavour
    else:
        counts[0] += (1 << 28) - int(counts.sum().item())  # a power-of-two total >= 2^24
    model = ctx.model_from_counts(counts)
    c, cum, total, _ = model.tables()
    stream, offsets, nbytes = ctx.encode_chunks(d_syms, chunk, model)
    ref_stream, ref_offsets = oracle.encode_chunks(syms, chunk, c, cum, total)
    assert_streams_equal(stream, offsets, nbytes, ref_stream, ref_offsets)
    out = ctx.decode_chunks(stream, offsets, n, chunk, model, sym_bytes=sb)
    assert np.array_equal(dev_to_np(out, dtype=syms.dtype), syms)


def test_slack_total_per_chunk_models(ctx, oracle):
    """Per-chunk tables whose last symbol does not end at total_freq (and the last symbol is coded a
    lot): bytes must match the oracle and decode must return the symbols, through the chunk API."""
    rng = np.random.default_rng(5)
    n_chunks, K, chunk = 40, 32, 3000
    c = rng.integers(1, 500, size=(n_chunks, K)).astype(np.uint32)
    cum = np.zeros_like(c)
    total = np.zeros(n_chunks, dtype=np.uint32)
    for j in range(n_chunks):
        cum[j], t = oracle.calc_cum(c[j])
        total[j] = t + 7 + 2 * j  # slack above the last symbol
    syms = rng.integers(0, K, size=n_chunks * chunk).astype(np.uint8)
    syms[::3] = K - 1
    model = ctx.model_from_tables(c, cum, total)
    for j in (0, n_chunks - 1):
        assert not (model.tables(j)[3] & 4)  # not RCB_MODEL_REGULAR
    d = to_dev(ctx, syms)
    stream, offsets, nbytes = ctx.encode_chunks(d, chunk, model)
    ref_stream, ref_offsets = oracle.encode_chunks(syms, chunk, c, cum, total)
    assert_streams_equal(stream, offsets, nbytes, ref_stream, ref_offsets)
    out = ctx.decode_chunks(stream, offsets, syms.size, chunk, model)
    assert np.array_equal(dev_to_np(out), syms)
    # the container stores per-chunk c only (total = sum c): such a model cannot be framed
    import range_coder_rust_b200 as rcb
    with pytest.raises(rcb.RcbError):
        ctx.frame_encode(syms, chunk, model)


def test_kat2_empty_encoder_run(ctx, oracle):
    """SURVEY KAT-2: an Encoder that saw no symbol finishes to eight zero bytes (src/encoder.rs:40-46),
    and a Decoder built on them decodes zero symbols having consumed exactly those 8 bytes
    (src/decoder.rs:14-23)."""
    import ctypes

    from test_gpu_stream_api import StreamState, _p

    lib = ctx.lib
    st = StreamState()
    lib.rcb_stream_state_init(ctypes.byref(st))
    out = np.full(16, 0xAA, dtype=np.uint8)
    n_out = ctypes.c_uint64()
    rc = lib.rcb_encode_stream(ctx.h, ctypes.byref(st), None, 0, 1, None, _p(out), out.size, ctypes.byref(n_out),
                               None, 1)
    assert rc == 0 and n_out.value == 8
    assert out[:8].tobytes() == bytes(8) == oracle.encode(np.zeros(0, np.uint8), [1], [0], 1)
    assert out[8] == 0xAA
    model = ctx.model_from_tables(np.array([1, 5, 2], np.uint32), np.array([0, 1, 6], np.uint32), 8)
    st = StreamState()
    lib.rcb_stream_state_init(ctypes.byref(st))
    code = out[:8].copy()
    rc = lib.rcb_decode_stream(ctx.h, ctypes.byref(st), _p(code), 8, 0, 1, model.h, None)
    assert rc == 0 and st.consumed == 8 and st.data == 0 and st.lower_bound == 0
    # the same run through the chunk API: one chunk whose only content is the flush
    one = np.zeros(1, dtype=np.uint8)
    m1 = ctx.model_from_tables(np.array([7], np.uint32), np.array([0], np.uint32), 7)
    stream, offsets, nbytes = ctx.encode_chunks(to_dev(ctx, one), 1, m1)
    assert dev_to_np(stream, nbytes).tobytes() == oracle.encode(one, [7], [0], 7)


def test_corrupt_offsets_are_contained(ctx, oracle):
    """Offsets are caller data (a damaged index, a crafted frame): a chunk shorter than the 8 bytes
    Decoder::new pops, offsets beyond the stream or running backwards must end in a status, never in
    a fault, and the context must stay usable -- through every decode kernel family."""
    import range_coder_rust_b200 as rcb
    from range_coder_rust_b200 import _lib

    chunk, n = 4096, 40 * 4096
    for kind in ("static_pow2", "static_odd", "adaptive"):
        syms = oracle.generate(n, 256, 0x5EED0001, oracle.zipf_thresholds(256, 1.1))
        d = to_dev(ctx, syms)
        if kind == "adaptive":
            model = ctx.model_from_counts(ctx.histogram(d, 256, chunk_syms=chunk))
        else:
            counts = ctx.histogram(d, 256)
            if kind == "static_odd":
                counts[0] += 999
            model = ctx.model_from_counts(counts)
        stream, offsets, nbytes = ctx.encode_chunks(d, chunk, model)
        good = dev_to_np(offsets).copy()
        tight = stream[:(nbytes + 15) // 16 * 16].clone()  # nothing readable past the padded end
        status = torch.zeros(40, dtype=torch.int32, device=ctx.device)
        cases = []
        o = good.copy(); o[7] = o[8]; cases.append((o, [7]))                      # empty chunk
        o = good.copy(); o[7] = o[8] - 5; cases.append((o, [7]))                  # 5-byte chunk
        o = good.copy(); o[20] = nbytes + (1 << 40); cases.append((o, [19, 20]))  # far beyond the stream
        o = good.copy(); o[30] = 0; cases.append((o, [29]))                       # runs backwards
        o = good.copy(); o[:] = nbytes; cases.append((o, list(range(40))))        # every chunk empty, at the end
        for offs, bad in cases:
            with pytest.raises(rcb.RcbError) as e:
                ctx.decode_chunks(tight, to_dev(ctx, offs), n, chunk, model, status=status)
            assert e.value.code == _lib.RCB_ERR_TRUNCATED_STREAM
            st = status.cpu().numpy()
            assert all(st[b] == 6 for b in bad), (kind, st)
        back = ctx.decode_chunks(stream, offsets, n, chunk, model)
        assert np.array_equal(dev_to_np(back), syms)
    # the host entry point refuses non-monotone offsets before sizing any copy with them
    h_stream = np.zeros((nbytes + 31) // 16 * 16, dtype=np.uint8)
    h_stream[:nbytes] = dev_to_np(stream, nbytes)
    o = good.astype(np.uint64).copy()
    o[5] = o[6] + 1
    with pytest.raises(rcb.RcbError) as e:
        ctx.decode_host(h_stream, o, n, chunk, model)
    assert e.value.code == _lib.RCB_ERR_INVALID_ARGUMENT


def test_config5_shard_on_one_gpu(ctx, oracle):
    """BASELINE.json configs[4] as one of its eight ranks sees it: an 8 GiB shard of the 64 GiB stream
    generated at first = 3 * n (rank 3's symbols), the all-reduced count table (sum 2^36 > u32, so the
    2^31 rescale of DESIGN.md fires) stood in for by 8 x the local counts, encode + decode of all
    131072 chunks, and >= 1 % of the chunks (every 64th: 2048 chunks) byte-compared with the oracle."""
    free, _ = torch.cuda.mem_get_info(ctx.device)
    if free < 30 << 30:
        pytest.skip("needs ~24 GiB of free device memory")
    n, chunk, rank = 8 << 30, 65536, 3
    thr = oracle.zipf_thresholds(256, 1.1)
    d_syms = ctx.generate(n, 256, 0x5EED0001, thr, first=rank * n)
    counts = ctx.histogram(d_syms, 256)
    assert int(counts.sum().item()) == n
    counts *= 8  # what the all-reduce over 8 statistically identical shards delivers: sum = 2^36
    model = ctx.model_from_counts(counts)
    c, cum, total, flags = model.tables()
    rc, scaled = oracle.normalise(dev_to_np(counts).astype(np.uint64))
    rcum, rtotal = oracle.calc_cum(rc)
    assert scaled == 1 and total == rtotal == 1 << 31 and np.array_equal(c, rc) and np.array_equal(cum, rcum)
    stream, offsets, nbytes = ctx.encode_chunks(d_syms, chunk, model)
    assert 0.70 < nbytes / n < 0.74
    out = ctx.decode_chunks(stream, offsets, n, chunk, model)
    assert torch.equal(out, d_syms)
    del out
    offs = dev_to_np(offsets).astype(np.uint64)
    n_chunks = n // chunk
    picked = np.arange(5, n_chunks, 64)
    assert picked.size >= n_chunks // 100
    # the oracle generates the same symbols from the global index (no device->host copy of the input)
    for i in picked[:: max(1, picked.size // 16)]:
        ref_syms = oracle.generate(chunk, 256, 0x5EED0001, thr, first=rank * n + int(i) * chunk)
        assert np.array_equal(dev_to_np(d_syms[int(i) * chunk:(int(i) + 1) * chunk]), ref_syms)
    idx = torch.from_numpy(picked).to(ctx.device)
    sample = d_syms.view(n_chunks, chunk)[idx].reshape(-1).cpu().numpy()
    ref_stream, ref_offsets = oracle.encode_chunks(sample, chunk, rc, rcum, rtotal)
    for k, i in enumerate(picked):
        got = dev_to_np(stream[int(offs[i]):int(offs[i + 1])])
        assert np.array_equal(got, ref_stream[int(ref_offsets[k]):int(ref_offsets[k + 1])]), f"chunk {i}"


def test_count_normalisation_above_u32(ctx, oracle):
    """64 GiB config: the global count sum (2^36) does not fit total_freq: u32."""
    rng = np.random.default_rng(3)
    w = np.arange(1, 257, dtype=np.float64) ** -1.1
    counts = (w / w.sum() * float(1 << 36)).astype(np.uint64)
    counts[200] = 0
    counts[201] = 3  # must stay >= 1 after the shift
    d_counts = torch.from_numpy(counts.view(np.int64)).to(ctx.device)
    model = ctx.model_from_counts(d_counts)
    c, cum, total, flags = model.tables()
    rc, scaled = oracle.normalise(counts)
    rcum, rtotal = oracle.calc_cum(rc)
    assert scaled == 1 and total == rtotal == 1 << 31
    assert np.array_equal(c, rc) and np.array_equal(cum, rcum)
    assert c[200] == 0 and c[201] == 1 and (flags & 1)  # rescaled totals are powers of two
    used = np.flatnonzero(rc)
    p = rc[used].astype(np.float64)
    syms = rng.choice(used, size=300_000, p=p / p.sum()).astype(np.uint8)
    stream, offsets, nbytes = ctx.encode_chunks(to_dev(ctx, syms), 65536, model)
    ref_stream, ref_offsets = oracle.encode_chunks(syms, 65536, rc, rcum, rtotal)
    assert_streams_equal(stream, offsets, nbytes, ref_stream, ref_offsets)
    out = ctx.decode_chunks(stream, offsets, syms.size, 65536, model)
    assert np.array_equal(dev_to_np(out), syms)


# ------------------------------------------------------------ error behaviour
def test_error_statuses(ctx, oracle):
    import range_coder_rust_b200 as rcb
    from range_coder_rust_b200 import _lib

    c = np.array([5, 0, 5, 6], dtype=np.uint32)
    cum, total = oracle.calc_cum(c)
    model = ctx.model_from_tables(c, cum, total)
    good = np.array([0, 2, 3, 3, 2, 0] * 100, dtype=np.uint8)
    bad_zero = good.copy()
    bad_zero[300] = 1  # c_freq == 0: the reference would never return
    status = torch.zeros(6, dtype=torch.int32, device=ctx.device)
    with pytest.raises(rcb.RcbError) as e:
        ctx.encode_chunks(to_dev(ctx, bad_zero), 100, model, status=status)
    assert e.value.code == _lib.RCB_ERR_ZERO_FREQ_SYMBOL
    assert status.cpu().tolist() == [0, 0, 0, 1, 0, 0]
    bad_range = good.copy()
    bad_range[599] = 4  # index >= alphabet size
    with pytest.raises(rcb.RcbError) as e:
        ctx.encode_chunks(to_dev(ctx, bad_range), 100, model, status=status)
    assert e.value.code == _lib.RCB_ERR_SYMBOL_OUT_OF_RANGE
    assert status.cpu().tolist() == [0, 0, 0, 0, 0, 4]
    # truncated stream: drop the tail of the last chunk
    stream, offsets, nbytes = ctx.encode_chunks(to_dev(ctx, good), 100, model)
    offs = dev_to_np(offsets).copy()
    offs[-1] -= 3
    with pytest.raises(rcb.RcbError) as e:
        ctx.decode_chunks(stream, to_dev(ctx, offs), good.size, 100, model, status=status)
    assert e.value.code == _lib.RCB_ERR_TRUNCATED_STREAM
    assert status.cpu().tolist() == [0, 0, 0, 0, 0, 6]
    # zero total, cum > total
    with pytest.raises(rcb.RcbError) as e:
        ctx.model_from_tables(np.zeros(4, np.uint32), np.zeros(4, np.uint32), 0)
    assert e.value.code == _lib.RCB_ERR_ZERO_TOTAL
    with pytest.raises(rcb.RcbError) as e:
        ctx.model_from_tables(np.array([1, 1], np.uint32), np.array([0, 9], np.uint32), 4)
    assert e.value.code == _lib.RCB_ERR_INVALID_MODEL
    # inconsistent table (cum + c > total): the reference's overflow errors, per chunk
    c2 = np.array([10, 10], dtype=np.uint32)
    cum2 = np.array([0, 15], dtype=np.uint32)
    m2 = ctx.model_from_tables(c2, cum2, 16)
    with pytest.raises(ValueError, match="OVERFLOW"):
        oracle.encode(np.ones(50, dtype=np.uint8), c2, cum2, 16)
    with pytest.raises(rcb.RcbError) as e:
        ctx.encode_chunks(to_dev(ctx, np.ones(64, dtype=np.uint8)), 64, m2)
    assert e.value.code in (_lib.RCB_ERR_LOWER_OVERFLOW, _lib.RCB_ERR_UPPER_OVERFLOW)
    # output capacity
    with pytest.raises(rcb.RcbError) as e:
        small = torch.empty(64, dtype=torch.uint8, device=ctx.device)
        ctx.encode_chunks(to_dev(ctx, good), 100, model, out=small)
    assert e.value.code == _lib.RCB_ERR_OUT_CAPACITY


@pytest.mark.timeout(120)
@pytest.mark.parametrize("kind", ["static_pow2", "static_odd", "static_odd_fat", "adaptive", "k4096"])
def test_garbage_streams_terminate(ctx, oracle, kind):
    """Corrupt input must never hang or fault the decoder (the reference would panic or spin): random
    bytes in place of the code stream, through every kernel family.  Any status is acceptable; the call
    has to return and the context has to stay usable."""
    import range_coder_rust_b200 as rcb

    K = 4096 if kind == "k4096" else 256
    sb = 2 if K > 256 else 1
    chunk, n = 4096, 64 * 4096 + 123
    syms = oracle.generate(n, K, 0x5EED0001, oracle.zipf_thresholds(K, 1.1), sym_bytes=sb)
    d = to_dev(ctx, syms)
    if kind == "adaptive":
        model = ctx.model_from_counts(ctx.histogram(d, K, chunk_syms=chunk))
    else:
        counts = ctx.histogram(d, K)
        if kind == "static_odd":
            counts[0] += 999
        if kind == "static_odd_fat":  # every count above one LUT bucket: the fat-LUT kernel, general total
            counts += 200
            counts[0] += 999
        model = ctx.model_from_counts(counts)
    stream, offsets, nbytes = ctx.encode_chunks(d, chunk, model)
    rng = np.random.default_rng(3)
    for trial in range(3):
        junk = torch.from_numpy(rng.integers(0, 256, size=stream.numel(), dtype=np.uint8)).to(ctx.device)
        if trial == 2:
            junk.fill_(0xFF)  # lower ends in ones everywhere: the worst case for loop 2
        try:
            ctx.decode_chunks(junk, offsets, n, chunk, model, sym_bytes=sb)
        except rcb.RcbError:
            pass
    back = ctx.decode_chunks(stream, offsets, n, chunk, model, sym_bytes=sb)  # still works afterwards
    assert np.array_equal(dev_to_np(back, dtype=syms.dtype), syms)


def test_staging_overflow_retry(ctx, oracle):
    """A model whose rarest symbol dominates the data needs more than the default
    staging estimate only in pathological cases; force one via a tiny estimate."""
    c = np.array([1 << 20, 1], dtype=np.uint32)  # bound comes from min c = 1
    cum, total = oracle.calc_cum(c)
    model = ctx.model_from_tables(c, cum, total)
    syms = np.ones(50_000, dtype=np.uint8)  # only the rare symbol: ~2.5 bytes each
    stream, offsets, nbytes = ctx.encode_chunks(to_dev(ctx, syms), 10_000, model)
    ref_stream, ref_offsets = oracle.encode_chunks(syms, 10_000, c, cum, total)
    assert_streams_equal(stream, offsets, nbytes, ref_stream, ref_offsets)


def test_host_buffer_entry_points(ctx, oracle):
    thr = oracle.zipf_thresholds(256, 1.1)
    syms = oracle.generate(1_000_000, 256, 0x5EED0001, thr)
    c, cum, total = oracle.model_from_symbols(syms, 256)
    model = ctx.model_from_tables(c, cum, total)
    out, offsets, nbytes = ctx.encode_host(syms, 65536, model)
    ref_stream, ref_offsets = oracle.encode_chunks(syms, 65536, c, cum, total)
    assert nbytes == ref_stream.size and np.array_equal(offsets, ref_offsets)
    assert np.array_equal(out[:nbytes], ref_stream)
    back = ctx.decode_host(out[:nbytes], offsets, syms.size, 65536, model)
    assert np.array_equal(back, syms)


def test_host_buffer_pipeline_slices(ctx, oracle):
    """Large host batches run as slices on several streams; results must not depend on that."""
    thr = oracle.zipf_thresholds(256, 1.1)
    n, chunk = 48 * 1024 * 1024 + 4096 * 3, 16384  # 3075 chunks -> 4 slices
    syms = oracle.generate(n, 256, 0x5EED0001, thr)
    c, cum, total = oracle.model_from_symbols(syms, 256)
    model = ctx.model_from_tables(c, cum, total)
    out, offsets, nbytes = ctx.encode_host(syms, chunk, model)
    ref_stream, ref_offsets = oracle.encode_chunks(syms, chunk, c, cum, total)
    assert nbytes == ref_stream.size and np.array_equal(offsets, ref_offsets)
    assert np.array_equal(out[:nbytes], ref_stream)
    back = ctx.decode_host(out[:nbytes], offsets, syms.size, chunk, model)
    assert np.array_equal(back, syms)
    # per-chunk models through the same pipeline
    d = to_dev(ctx, syms)
    counts = ctx.histogram(d, 256, chunk_syms=chunk)
    pm = ctx.model_from_counts(counts)
    out2, offsets2, nbytes2 = ctx.encode_host(syms, chunk, pm)
    s2, o2, nb2 = ctx.encode_chunks(d, chunk, pm)
    assert nbytes2 == nb2 and np.array_equal(offsets2, dev_to_np(o2).astype(np.uint64))
    assert np.array_equal(out2[:nbytes2], dev_to_np(s2, nb2))
    back2 = ctx.decode_host(out2[:nbytes2], offsets2, syms.size, chunk, pm)
    assert np.array_equal(back2, syms)


@pytest.mark.parametrize("K,chunk,kind", [(256, 65536, "static"), (256, 65536, "adaptive"), (4096, 32768, "static"),
                                          (256, 32768, "odd_total")])
def test_host_buffer_pipeline_segments(ctx, oracle, K, chunk, kind):
    """decode_host decodes big chunks in several launches per chunk (lane state handed over through
    device memory) so that copy-out overlaps decoding: every kernel family, ragged last chunk."""
    sb = 1 if K <= 256 else 2
    n = 600 * chunk + 12345  # 601 chunks -> slices x 4 segments per chunk, last chunk ragged
    thr = oracle.zipf_thresholds(K, 1.1)
    if kind == "adaptive":
        thr = np.stack([oracle.zipf_thresholds(K, s) for s in (0.0, 0.8, 1.1, 2.0, 5.0)])
        syms = oracle.generate(n, K, 0x5EED0002, thr, sym_bytes=sb, chunk_syms=chunk)
    else:
        syms = oracle.generate(n, K, 0x5EED0001, thr, sym_bytes=sb)
    d = to_dev(ctx, syms)
    if kind == "adaptive":
        model = ctx.model_from_counts(ctx.histogram(d, K, chunk_syms=chunk))
    else:
        counts = ctx.histogram(d, K)
        if kind == "odd_total":
            counts[0] += 12345
        model = ctx.model_from_counts(counts)
    stream, offsets, nbytes = ctx.encode_chunks(d, chunk, model)
    h_stream = np.zeros((nbytes + 15) // 16 * 16 + 16, dtype=np.uint8)
    h_stream[:nbytes] = dev_to_np(stream, nbytes)
    h_offsets = dev_to_np(offsets).astype(np.uint64)
    back = ctx.decode_host(h_stream, h_offsets, syms.size, chunk, model, sym_bytes=sb)
    assert np.array_equal(back, syms)
    # and the oracle agrees on a few chunks of that stream (first, middle, ragged last)
    for j in (0, 300, 600):
        c, cum, total, _ = model.tables(j if kind == "adaptive" else 0)
        code = h_stream[int(h_offsets[j]):int(h_offsets[j + 1])]
        ref, _ = oracle.decode(code, min(chunk, n - j * chunk), c, cum, total, sb)
        assert np.array_equal(ref, syms[j * chunk:(j + 1) * chunk])


# ------------------------------------------------ full-size properties (1 GiB)
def test_full_size_round_trip_and_sampled_parity(ctx, oracle):
    """BASELINE.json configs[1] at full size: 1 GiB Zipf(1.1), 64 KiB chunks.
    All chunks: GPU encode -> GPU decode == input (torch.equal on device).
    Sampled chunks (every 64th, 256 chunks = 16 MiB): bytes identical to the oracle."""
    n, chunk = 1 << 30, 65536
    thr = oracle.zipf_thresholds(256, 1.1)
    d_syms = ctx.generate(n, 256, 0x5EED0001, thr)
    counts = ctx.histogram(d_syms, 256)
    assert int(counts.sum().item()) == n
    model = ctx.model_from_counts(counts)
    c, cum, total, flags = model.tables()
    assert total == n and (flags & 1)
    stream, offsets, nbytes = ctx.encode_chunks(d_syms, chunk, model)
    ratio = nbytes / n
    assert 0.70 < ratio < 0.74  # ~0.722 for Zipf(1.1), K=256
    out = ctx.decode_chunks(stream, offsets, n, chunk, model)
    assert torch.equal(out, d_syms)
    offs = dev_to_np(offsets).astype(np.uint64)
    n_chunks = n // chunk
    for i in range(0, n_chunks, 64):
        s = dev_to_np(d_syms[i * chunk:(i + 1) * chunk])
        ref = oracle.encode(s, c, cum, total)
        got = dev_to_np(stream[int(offs[i]):int(offs[i + 1])]).tobytes()
        assert got == ref, f"chunk {i} differs from the oracle"


# ------------------------------------------------- multi-GPU exchange step (C ABI)
def test_comm_single_rank_allreduce(ctx, oracle):
    """rcb_comm_* / rcb_allreduce_counts with a one-rank communicator: NCCL is found at run time, the
    call is issued on the context's stream, and a sum over one rank leaves the counts unchanged --
    the model built after it is the single-GPU model."""
    import range_coder_rust_b200 as rcb

    syms = oracle.generate(1 << 20, 256, 0x5EED0001, oracle.zipf_thresholds(256, 1.1))
    d = to_dev(ctx, syms)
    counts = ctx.histogram(d, 256)
    before = counts.clone()
    comm = ctx.comm_init_rank(rcb.Comm.unique_id(), 1, 0)
    assert comm.nccl_version >= 21800
    n0 = ctx.launch_count
    ctx.allreduce_counts(counts, comm)
    model = ctx.model_from_counts(counts)  # same stream: ordered after the all-reduce
    assert ctx.launch_count > n0
    assert torch.equal(counts, before)
    c, cum, total, _ = model.tables()
    rc, rcum, rtotal = oracle.model_from_symbols(syms, 256)
    assert total == rtotal and np.array_equal(c, rc) and np.array_equal(cum, rcum)
    comm.close()


def test_comm_two_gpus_share_one_table(oracle):
    """Two GPUs driven by one thread (rcb_comm_init_all + rcb_allreduce_counts_multi): each codes its
    own half of the chunks under the table built from the summed counts; both tables equal the
    oracle's table of the whole stream and the concatenated streams equal the oracle's stream."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import range_coder_rust_b200 as rcb

    n, chunk, K = 64 * 65536, 65536, 256
    thr = oracle.zipf_thresholds(K, 1.1)
    whole = oracle.generate(n, K, 0x5EED0001, thr)
    ctxs = [rcb.Context(i) for i in range(2)]
    comms = rcb.Context.comm_init_all(ctxs)
    half = n // 2
    shards = [ctxs[i].generate(half, K, 0x5EED0001, thr, first=i * half) for i in range(2)]
    counts = [ctxs[i].histogram(shards[i], K) for i in range(2)]
    rcb.Context.allreduce_counts_multi(ctxs, comms, counts)
    rc, rcum, rtotal = oracle.model_from_symbols(whole, K)
    ref_stream, ref_offsets = oracle.encode_chunks(whole, chunk, rc, rcum, rtotal)
    parts = []
    for i in range(2):
        model = ctxs[i].model_from_counts(counts[i])
        c, cum, total, _ = model.tables()
        assert total == rtotal and np.array_equal(c, rc) and np.array_equal(cum, rcum)
        stream, offsets, nbytes = ctxs[i].encode_chunks(shards[i], chunk, model)
        parts.append(stream[:nbytes].cpu().numpy())
        back = ctxs[i].decode_chunks(stream, offsets, half, chunk, model)
        assert torch.equal(back, shards[i])
    assert np.array_equal(np.concatenate(parts), ref_stream)
    for k in comms:
        k.close()
    for c in ctxs:
        c.close()


# ------------------------------------------ f4: adaptive-per-symbol table on the GPU
@pytest.mark.parametrize("K,inc,limit,chunk,n", [(256, 24, 60000, 65536, 40 * 65536 + 777), (256, 32, 4096, 4096, 300_000),
                                                  (10, 5, 200, 1000, 123_456), (1000, 24, 65000, 8192, 200_000),
                                                  (256, 24, 60000, 1, 50), (2, 1, 2, 100, 10_000)])
def test_adaptive_per_symbol_model_matches_oracle(ctx, oracle, K, inc, limit, chunk, n):
    """the table changes after every symbol (counts from 1, +inc, halving at `limit`) and
    restarts with every chunk; rcb_adaptive_encode_chunks / decode_chunks against the oracle's caller-side
    loop over the reference semantics -- bytes per chunk, symbols back, and the oracle's stream decoded on
    the GPU."""
    rng = np.random.default_rng(K + chunk)
    sb = 2 if K > 256 else 1
    w = np.arange(1, K + 1, dtype=np.float64) ** -1.2
    syms = rng.choice(K, size=n, p=w / w.sum()).astype(np.uint16 if sb == 2 else np.uint8)
    syms[n // 3: n // 2] = K - 1 - syms[n // 3: n // 2]  # moving statistics
    d_syms = to_dev(ctx, syms)
    stream, offsets, nbytes = ctx.adaptive_encode_chunks(d_syms, chunk, K, inc, limit)
    ref_stream, ref_offsets = oracle.adaptive_encode_chunks(syms, chunk, K, inc, limit)
    assert_streams_equal(stream, offsets, nbytes, ref_stream, ref_offsets)
    out = ctx.adaptive_decode_chunks(stream, offsets, n, chunk, K, inc, limit, sym_bytes=sb)
    assert np.array_equal(dev_to_np(out, dtype=syms.dtype), syms)
    pad = np.zeros(((ref_stream.size + 31) // 16) * 16, dtype=np.uint8)
    pad[:ref_stream.size] = ref_stream
    out2 = ctx.adaptive_decode_chunks(to_dev(ctx, pad), to_dev(ctx, ref_offsets.view(np.int64)), n, chunk, K, inc, limit,
                                      sym_bytes=sb)
    assert np.array_equal(dev_to_np(out2, dtype=syms.dtype), syms)


def test_adaptive_per_symbol_errors(ctx, oracle):
    import range_coder_rust_b200 as rcb
    from range_coder_rust_b200 import _lib

    syms = np.array([0, 1, 2, 9, 1], dtype=np.uint8)
    with pytest.raises(rcb.RcbError) as e:  # symbol >= K
        ctx.adaptive_encode_chunks(to_dev(ctx, syms), 5, 4, 8, 1000)
    assert e.value.code == _lib.RCB_ERR_SYMBOL_OUT_OF_RANGE
    with pytest.raises(rcb.RcbError) as e:  # u16 counters: limit + inc must fit
        ctx.adaptive_encode_chunks(to_dev(ctx, syms), 5, 16, 100, 65500)
    assert e.value.code == _lib.RCB_ERR_UNSUPPORTED
    good = np.arange(200, dtype=np.uint8) % 16
    stream, offsets, nbytes = ctx.adaptive_encode_chunks(to_dev(ctx, good), 50, 16, 8, 1000)
    offs = dev_to_np(offsets).copy()
    offs[-1] -= 2  # truncated last chunk
    with pytest.raises(rcb.RcbError) as e:
        ctx.adaptive_decode_chunks(stream, to_dev(ctx, offs), good.size, 50, 16, 8, 1000)
    assert e.value.code == _lib.RCB_ERR_TRUNCATED_STREAM


def test_adaptive_golden_vectors_on_gpu(ctx):
    import hashlib as hl

    for v in json.load(open(os.path.join(ROOT, "tests", "golden", "adaptive_vectors.json"))):
        K, a = v["K"], v["adaptive"]
        syms = np.frombuffer(bytes.fromhex(v["symbols_hex"]), dtype="<u2" if K > 256 else np.uint8).copy()
        stream, offsets, nbytes = ctx.adaptive_encode_chunks(to_dev(ctx, syms), syms.size, K, a["inc"], a["limit"])
        code = dev_to_np(stream, nbytes).tobytes()
        assert nbytes == v["code_len"] and hl.sha256(code).hexdigest() == v["code_sha256"], v["name"]
        out = ctx.adaptive_decode_chunks(stream, offsets, syms.size, syms.size, K, a["inc"], a["limit"],
                                         sym_bytes=syms.dtype.itemsize)
        assert np.array_equal(dev_to_np(out, dtype=syms.dtype), syms)
