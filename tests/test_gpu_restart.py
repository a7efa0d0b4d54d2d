"""Restart points (include/rcb200.h: rcb_restart_point): several decoder lanes per chunk.

The code bytes stay the reference's: every test compares the stream with the oracle first.  The
records are checked against the oracle's Encoder state at the same symbol (rco_encode_state:
src/encoder.rs:24-37 without finish), the decode against the input, for every decode kernel
family, ragged last chunks included; damaged records must be reported, not silently decoded.
"""
import numpy as np
import pytest

from test_gpu_parity import S_CYCLE, VECTORS, assert_streams_equal, dev_to_np, to_dev, vector_inputs

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def records(restart_t, n_chunks, per):
    """int64[n_chunks*per*3] -> (lower u64[n_chunks][per], range u64, code_bytes u32)."""
    a = restart_t.cpu().numpy().view(np.uint64).reshape(n_chunks, per, 3)
    return a[:, :, 0], a[:, :, 1], (a[:, :, 2] & np.uint64(0xFFFFFFFF)).astype(np.uint64)


def check_records(oracle, restart_t, syms, chunk, rs, tables, picks):
    """Records of the picked chunks == the oracle's Encoder state in front of the same symbol."""
    n = syms.size
    n_chunks = (n + chunk - 1) // chunk
    per = (chunk + rs - 1) // rs - 1
    lo, rg, pos = records(restart_t, n_chunks, per)
    c, cum, total = tables
    for i in picks:
        ci, cumi, ti = (c[i], cum[i], int(total[i])) if np.ndim(total) else (c, cum, int(total))
        part = syms[i * chunk:(i + 1) * chunk]
        for r in range(per):
            j = (r + 1) * rs
            if j >= part.size:  # ragged last chunk: absent
                assert rg[i, r] == 0 and lo[i, r] == 0 and pos[i, r] == 0
                continue
            ref_lo, ref_rg, ref_n = oracle.encode_state(part[:j], ci, cumi, ti)
            assert int(lo[i, r]) == ref_lo, (i, r)
            assert int(pos[i, r]) == ref_n, (i, r)
            # the record holds range rounded down to a multiple of total_freq (only range / total is used)
            assert int(rg[i, r]) == ref_rg // ti * ti, (i, r)


def round_trip(ctx, oracle, syms, chunk, rs, model, tables, sb=1, picks=(0,)):
    n = syms.size
    n_chunks = (n + chunk - 1) // chunk
    d_syms = to_dev(ctx, syms)
    restart = ctx.restart_points(n_chunks, chunk, rs)
    assert restart is not None
    stream, offsets, nbytes = ctx.encode_chunks(d_syms, chunk, model, restart_syms=rs, restart=restart)
    ref_stream, ref_offsets = oracle.encode_chunks(syms, chunk, *tables)
    assert_streams_equal(stream, offsets, nbytes, ref_stream, ref_offsets)  # the reference's bytes, unchanged
    check_records(oracle, restart, syms, chunk, rs, tables, picks)
    status = torch.full((n_chunks,), 99, dtype=torch.int32, device=ctx.device)
    out = ctx.decode_chunks(stream, offsets, n, chunk, model, sym_bytes=sb, status=status, restart_syms=rs,
                            restart=restart)
    assert np.array_equal(dev_to_np(out, dtype=syms.dtype), syms)
    assert not status.any()
    # the plain decoder on the same stream (restart points are optional side information)
    out1 = ctx.decode_chunks(stream, offsets, n, chunk, model, sym_bytes=sb)
    assert np.array_equal(dev_to_np(out1, dtype=syms.dtype), syms)
    return d_syms, stream, offsets, restart


@pytest.mark.parametrize("n,chunk,rs", [(4 << 20, 65536, 16384), ((2 << 20) + 70001, 65536, 8192),
                                        (1_000_003, 65536, 1024), (300_000, 4096, 64), (70_000, 100_000, 4096),
                                        (65536 * 3 + 64, 65536, 16384), (65536 * 2 + 16384, 65536, 16384)])
def test_static_table_restart_points(ctx, oracle, n, chunk, rs):
    """Shared static table, power-of-two total (the fused fat-LUT decoder): bytes, records, symbols."""
    thr = oracle.zipf_thresholds(256, 1.1)
    syms = oracle.generate(n, 256, 0x5EED0001, thr)
    # a total of exactly 2^k needs n = 2^k symbols: build the table from a 1 MiB prefix scaled to 2^24
    c0 = np.maximum(1, np.bincount(syms[:1 << 20], minlength=256)).astype(np.int64)
    c = (c0 * (1 << 24) // c0.sum()).astype(np.int64)
    c = np.maximum(c, 4096 + 600)
    c[0] += (1 << 24) - int(c.sum())
    c = c.astype(np.uint32)
    cum, total = oracle.calc_cum(c)
    assert total == 1 << 24
    model = ctx.model_from_tables(c, cum, total)
    n_chunks = (n + chunk - 1) // chunk
    round_trip(ctx, oracle, syms, chunk, rs, model, (c, cum, total), picks=sorted({0, n_chunks // 2, n_chunks - 1}))


@pytest.mark.parametrize("v", [v for v in VECTORS if "restart" in v], ids=[v["name"] for v in VECTORS if "restart" in v])
def test_golden_restart_records_on_gpu(ctx, v):
    """tests/golden/vectors.json: the records of one chunk = the whole vector against the committed fixture
    (which rust/pin_reference checks against the unmodified crate), then both decoders."""
    syms, c, cum, total = vector_inputs(v)
    rs = v["restart"]["restart_syms"]
    recs = v["restart"]["records"]
    n = syms.size
    model = ctx.model_from_tables(c, cum, total)
    restart = ctx.restart_points(1, n, rs)
    stream, offsets, nbytes = ctx.encode_chunks(to_dev(ctx, syms), n, model, restart_syms=rs, restart=restart)
    assert nbytes == v["code_len"]
    lo, rg, pos = records(restart, 1, len(recs))
    for r, (ref_lo, ref_rg, ref_n) in enumerate(recs):
        assert int(lo[0, r]) == int(ref_lo, 16), (v["name"], r)
        assert int(rg[0, r]) == int(ref_rg, 16) // total * total, (v["name"], r)  # stored as a multiple of total_freq
        assert int(pos[0, r]) == ref_n, (v["name"], r)
    for kw in ({"restart_syms": rs, "restart": restart}, {}):
        out = ctx.decode_chunks(stream, offsets, n, n, model, sym_bytes=syms.dtype.itemsize, **kw)
        assert np.array_equal(dev_to_np(out, dtype=syms.dtype), syms)


@pytest.mark.parametrize("kind", ["gen_2e30", "gen_prime", "pow2_2e20", "gen_u16", "gen_full_c", "irregular",
                                  "gaps"])
def test_restart_points_other_tables(ctx, oracle, kind):
    """The other shared-table kernels: general totals (table-wide and per-symbol reciprocal), small powers of
    two, the row kernel (u16, K = 1000), the generic kernels (irregular table, zero-frequency gaps)."""
    rng = np.random.default_rng(23)
    K, sb = 256, 1
    if kind == "irregular":
        c = rng.integers(1, 1000, size=64).astype(np.uint32)
        cum, total = oracle.calc_cum(c)
        cum = (cum + np.arange(64, dtype=np.uint32) * 3).astype(np.uint32)
        total = int(cum[-1] + c[-1] + 10)
    elif kind == "gaps":
        c = (rng.integers(0, 2, size=200) * rng.integers(1, 90000, size=200)).astype(np.uint32)
        c[17] = 5
        cum, total = oracle.calc_cum(c)
    else:
        if kind == "gen_2e30":
            total = (1 << 30) + 12345
        elif kind == "gen_prime":
            total = 1_000_003
        elif kind == "pow2_2e20":
            total = 1 << 20
        elif kind == "gen_u16":
            K, total, sb = 1000, (1 << 28) + 7, 2
        else:
            K, total = 3, 1_000_003
        if kind == "gen_full_c":
            c = np.array([0, total, 0], dtype=np.uint32)
        else:
            w = np.arange(1, K + 1, dtype=np.float64) ** -0.9
            floor_c = max(2, int(total / 4096 * 1.2) + 2)
            c = np.maximum(floor_c, (w / w.sum() * (total - K * floor_c)).astype(np.int64))
            c[0] += total - int(c.sum())
            c = c.astype(np.uint32)
        cum, t = oracle.calc_cum(c)
        assert t == total
    n, chunk, rs = 700_000 + 333, 32768, 4096
    used = np.flatnonzero(c)
    p = c[used].astype(np.float64)
    syms = rng.choice(used, size=n, p=p / p.sum()).astype(np.uint16 if sb == 2 else np.uint8)
    model = ctx.model_from_tables(c, cum, total)
    n_chunks = (n + chunk - 1) // chunk
    round_trip(ctx, oracle, syms, chunk, rs, model, (c, cum, total), sb=sb, picks=(0, n_chunks - 1))


@pytest.mark.parametrize("chunk,rs", [(65536, 16384), (16384, 4096), (262144, 16384), (65536, 32768)])
def test_restart_points_per_chunk_tables(ctx, oracle, chunk, rs):
    """configs[2]: one table per chunk; the lanes of a chunk share its shared-memory row."""
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
    model = ctx.model_from_counts(counts)
    round_trip(ctx, oracle, syms, chunk, rs, model, (ref_c, ref_cum, ref_total), picks=(0, 7, n_chunks - 1))


def test_restart_points_k4096(ctx, oracle):
    """configs[3]: 4096 symbols, u16 storage (row kernel with a block-wide table)."""
    n, chunk, rs = 3 * 1024 * 1024 + 11, 32768, 8192
    thr = oracle.zipf_thresholds(4096, 1.1)
    syms = oracle.generate(n, 4096, 0x5EED0003, thr, sym_bytes=2)
    model = ctx.model_from_counts(ctx.histogram(to_dev(ctx, syms), 4096))
    tables = oracle.model_from_symbols(syms, 4096)
    round_trip(ctx, oracle, syms, chunk, rs, model, tables, sb=2, picks=(0, (n + chunk - 1) // chunk - 1))


def test_damaged_restart_points_are_reported(ctx, oracle):
    """A record that does not belong to the stream: the lane before it arrives elsewhere (RCB_ST_RESTART),
    a position outside the chunk decodes nothing; no other chunk is touched and nothing is read out of bounds."""
    n, chunk, rs = 40 * 65536, 65536, 16384
    thr = oracle.zipf_thresholds(256, 1.1)
    syms = oracle.generate(n, 256, 0x5EED0001, thr)
    c, cum, total = oracle.model_from_symbols(syms, 256)
    model = ctx.model_from_tables(c, cum, total)
    n_chunks, per = n // chunk, chunk // rs - 1
    d_syms, stream, offsets, restart = round_trip(ctx, oracle, syms, chunk, rs, model, (c, cum, total))
    bad = restart.clone().view(n_chunks, per, 3)
    bad[3, 1, 0] ^= 1 << 40            # lower_bound of chunk 3, record 1
    bad[9, 0, 2] = 0x7FFFFFF0          # code_bytes beyond the chunk
    bad[12, 2, 2] += 1                 # off by one byte
    bad[20, 1, 1] = 0                  # range 0 = "absent" inside a full chunk
    status = torch.zeros(n_chunks, dtype=torch.int32, device=ctx.device)
    from range_coder_rust_b200 import RcbError
    with pytest.raises(RcbError) as e:
        ctx.decode_chunks(stream, offsets, n, chunk, model, status=status, restart_syms=rs, restart=bad.view(-1))
    assert e.value.code == -14  # RCB_ERR_RESTART_POINT
    st = status.cpu().numpy()
    assert set(np.flatnonzero(st)) == {3, 9, 12, 20} and (st[[3, 9, 12, 20]] == 7).all()


def test_restart_argument_checks(ctx, oracle):
    syms = np.zeros(4096, dtype=np.uint8)
    c, cum, total = oracle.model_from_symbols(syms, 2)
    model = ctx.model_from_tables(c, cum, total)
    d = to_dev(ctx, syms)
    from range_coder_rust_b200 import RcbError
    buf = torch.zeros(3 * 64 * 8, dtype=torch.int64, device=ctx.device)
    with pytest.raises(RcbError):  # not a multiple of 64
        ctx.encode_chunks(d, 4096, model, restart_syms=100, restart=buf)
    with pytest.raises(RcbError):  # more than 64 parts per chunk
        ctx.encode_chunks(d, 4096 * 128, model, restart_syms=64, restart=buf)
    # a single part: plain call
    assert ctx.restart_points(1, 4096, 4096) is None
    stream, offsets, nbytes = ctx.encode_chunks(d, 4096, model, restart_syms=4096, restart=buf)
    assert nbytes == len(oracle.encode(syms, c, cum, total))
