"""CPU tests of the product's per-symbol arithmetic (rcb_core.cuh compiled for the
host by tests/hostcore) against the oracle: closed-form renormalisation, byte
sinks, reciprocal division, table-driven symbol lookup with exact fallback."""
import ctypes
import zlib

import numpy as np
import pytest


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def hc_encode(L, syms, c, cum, total, cap=None, checked=0):
    syms = np.ascontiguousarray(syms)
    cap = cap if cap is not None else 16 * syms.size + 64
    out = np.zeros(cap + 8, dtype=np.uint8)
    st = ctypes.c_uint32()
    n = L.hc_encode(_p(syms), syms.size, syms.dtype.itemsize, c.size, _p(c), _p(cum), total, _p(out), cap, checked,
                    ctypes.byref(st))
    return out[:min(n, cap)].tobytes(), int(n), st.value


def hc_decode(L, stream, off0, off1, n_syms, c, cum, total, sym_bytes=1, use_lut=1, checked=0, lut_cap=4096):
    stream = np.ascontiguousarray(stream, dtype=np.uint8)
    out = np.zeros(n_syms, dtype=np.uint8 if sym_bytes == 1 else np.uint16)
    st = ctypes.c_uint32()
    fb = ctypes.c_uint64()
    used = L.hc_decode(_p(stream), off0, off1, stream.size, n_syms, sym_bytes, c.size, _p(c), _p(cum), total,
                       _p(out), use_lut, checked, lut_cap, ctypes.byref(st), ctypes.byref(fb))
    return out, int(used), st.value, int(fb.value)


def random_model(rng, K, kind):
    if kind == "pow2":
        counts = rng.integers(1, 1000, size=K).astype(np.uint64)
        counts[0] += (1 << 24) - counts.sum()
    elif kind == "big":
        counts = rng.integers(1, 2 ** 32 // K, size=K).astype(np.uint64)
    elif kind == "gaps":
        counts = (rng.integers(0, 2, size=K) * rng.integers(1, 5000, size=K)).astype(np.uint64)
        counts[K // 2] = max(counts[K // 2], 1)
    else:  # zipf-like
        w = np.arange(1, K + 1, dtype=np.float64) ** -1.1
        counts = np.maximum(1, (w / w.sum() * (1 << 26)).astype(np.uint64))
    return counts.astype(np.uint32)


@pytest.mark.parametrize("kind", ["pow2", "big", "gaps", "zipf"])
@pytest.mark.parametrize("K", [2, 10, 256, 4096])
def test_encode_decode_match_oracle(oracle, hostcore, K, kind):
    rng = np.random.default_rng(zlib.crc32(f"{K}-{kind}".encode()))
    c = random_model(rng, K, kind)
    cum, total = oracle.calc_cum(c)
    used_syms = np.flatnonzero(c)
    p = c[used_syms].astype(np.float64)
    n = 6000
    syms = rng.choice(used_syms, size=n, p=p / p.sum()).astype(np.uint16 if K > 256 else np.uint8)
    ref = oracle.encode(syms, c, cum, total)
    code, length, st = hc_encode(hostcore, syms, c, cum, total)
    assert st == 0 and length == len(ref)
    assert code == ref
    # decode: table-driven and exact-only paths, chunk embedded at every byte alignment
    for shift in range(4):
        buf = np.zeros(shift + len(ref) + 32, dtype=np.uint8)
        buf[shift:shift + len(ref)] = np.frombuffer(ref, dtype=np.uint8)
        buf[shift + len(ref):] = 0xA5  # following chunk's bytes must not matter
        for use_lut in (1, 0):
            dec, used, st, fb = hc_decode(hostcore, buf, shift, shift + len(ref), n, c, cum, total,
                                          sym_bytes=syms.dtype.itemsize, use_lut=use_lut)
            assert st == 0 and used == len(ref)
            assert np.array_equal(dec, syms)
            if use_lut and K <= 256 and kind in ("pow2", "zipf"):
                assert fb < n * 0.05  # the LUT resolves nearly everything


def test_fused_path_slow_events_and_degenerate_streams(oracle, hostcore):
    """total = 2^30 (the benchmark's shape): loop-2 events inside the fused loop, and streams whose
    value sits at the very bottom of the range (data - lower == 0 for every symbol)."""
    rng = np.random.default_rng(77)
    w = np.arange(1, 257, dtype=np.float64) ** -1.1
    c = np.maximum(1, (w / w.sum() * (1 << 30)).astype(np.int64))
    c[0] += (1 << 30) - c.sum()
    c = c.astype(np.uint32)
    cum, total = oracle.calc_cum(c)
    assert total == 1 << 30
    n = 300_000
    syms = rng.choice(256, size=n, p=c / c.sum()).astype(np.uint8)
    ref = oracle.encode(syms, c, cum, total)
    code, length, st = hc_encode(hostcore, syms, c, cum, total)
    assert st == 0 and code == ref
    dec, used, st, fb = hc_decode(hostcore, np.frombuffer(ref + bytes(32), dtype=np.uint8), 0, len(ref), n, c, cum,
                                  total)
    assert st == 0 and used == len(ref) and np.array_equal(dec, syms)
    assert 0 < fb < n * 0.01  # loop-2 events take the exact path, nearly everything else the table
    for sym in (0, 255):
        syms = np.full(20_000, sym, dtype=np.uint8)
        ref = oracle.encode(syms, c, cum, total)
        code, length, st = hc_encode(hostcore, syms, c, cum, total)
        assert st == 0 and code == ref
        dec, used, st, fb = hc_decode(hostcore, np.frombuffer(ref + bytes(32), dtype=np.uint8), 0, len(ref),
                                      syms.size, c, cum, total)
        assert st == 0 and used == len(ref) and np.array_equal(dec, syms)
        assert fb < syms.size * 0.02


def test_small_lut_forces_fallbacks_but_stays_exact(oracle, hostcore):
    rng = np.random.default_rng(5)
    c = random_model(rng, 256, "zipf")
    cum, total = oracle.calc_cum(c)
    syms = rng.integers(0, 256, size=4000).astype(np.uint8)
    ref = np.frombuffer(oracle.encode(syms, c, cum, total), dtype=np.uint8)
    dec, used, st, fb = hc_decode(hostcore, np.concatenate([ref, np.zeros(16, np.uint8)]), 0, ref.size, syms.size,
                                  c, cum, total, lut_cap=16)
    assert st == 0 and used == ref.size and np.array_equal(dec, syms)
    assert fb > 0


def test_slow_path_events_are_exercised(oracle, hostcore):
    """total = 2^32-1 with c = 1 symbols: 4-5 bytes per symbol, loop 2 fires often."""
    c = np.array([1, 0xFFFFFFFF - 3, 2], dtype=np.uint32)
    cum, total = oracle.calc_cum(c)
    rng = np.random.default_rng(9)
    syms = rng.integers(0, 3, size=5000).astype(np.uint8)
    ref = oracle.encode(syms, c, cum, total)
    code, length, st = hc_encode(hostcore, syms, c, cum, total)
    assert st == 0 and code == ref
    dec, used, st, _ = hc_decode(hostcore, np.frombuffer(ref + bytes(16), dtype=np.uint8), 0, len(ref), syms.size,
                                 c, cum, total)
    assert st == 0 and used == len(ref) and np.array_equal(dec, syms)


def test_empty_and_single_symbol(oracle, hostcore):
    c = np.array([7], dtype=np.uint32)
    cum, total = oracle.calc_cum(c)
    code, length, st = hc_encode(hostcore, np.zeros(0, dtype=np.uint8), c, cum, total)
    assert code == bytes(8) and st == 0
    syms = np.zeros(10000, dtype=np.uint8)
    code, length, st = hc_encode(hostcore, syms, c, cum, total)
    assert code == oracle.encode(syms, c, cum, total) and length == 8
    dec, used, st, _ = hc_decode(hostcore, np.frombuffer(code + bytes(16), dtype=np.uint8), 0, 8, 10000, c, cum, total)
    assert st == 0 and used == 8 and not dec.any()


def test_status_codes(oracle, hostcore):
    c = np.array([1, 0, 1], dtype=np.uint32)
    cum, total = oracle.calc_cum(c)
    _, _, st = hc_encode(hostcore, np.array([0, 1, 2], dtype=np.uint8), c, cum, total)
    assert st == 1  # ST_ZERO_FREQ
    _, _, st = hc_encode(hostcore, np.array([0, 3], dtype=np.uint8), c, cum, total)
    assert st == 4  # ST_SYMBOL_RANGE
    # capacity: the reported length is the needed size
    rng = np.random.default_rng(3)
    c = np.ones(256, dtype=np.uint32)
    cum, total = oracle.calc_cum(c)
    syms = rng.integers(0, 256, size=1000).astype(np.uint8)
    ref = oracle.encode(syms, c, cum, total)
    _, length, st = hc_encode(hostcore, syms, c, cum, total, cap=100)
    assert st == 5 and length == len(ref)
    # inconsistent table: checked mode reports the reference's overflow errors
    c = np.array([10, 10], dtype=np.uint32)
    cum = np.array([0, 15], dtype=np.uint32)
    with pytest.raises(ValueError, match="OVERFLOW"):
        oracle.encode(np.ones(50, dtype=np.uint8), c, cum, 16)
    _, _, st = hc_encode(hostcore, np.ones(50, dtype=np.uint8), c, cum, 16, checked=1)
    assert st in (2, 3)
    # truncated stream
    c = np.array([1, 1], dtype=np.uint32)
    cum, total = oracle.calc_cum(c)
    code = np.frombuffer(oracle.encode(np.ones(100, dtype=np.uint8), c, cum, total), dtype=np.uint8)
    _, _, st, _ = hc_decode(hostcore, np.concatenate([code, np.zeros(16, np.uint8)]), 0, code.size - 1, 100, c, cum,
                            total)
    assert st == 6


@pytest.mark.parametrize("total", [2, 3, 5, 255, 65537, 1000003, (1 << 30) - 125, (1 << 31) + 1, 0xFFFFFFFF,
                                   1 << 16, 1 << 30, 1])
def test_reciprocal_division_is_exact(hostcore, total):
    rng = np.random.default_rng(total % 9973)
    r = rng.integers(1 << 48, (1 << 64) - 1, size=200_000, dtype=np.uint64, endpoint=True)
    edge = np.array([(1 << 48), (1 << 64) - 1, (1 << 48) + total - 1, ((1 << 64) - 1) // total * total,
                     ((1 << 64) - 1) // total * total - 1, 0, 1, total - 1, total, total + 1], dtype=np.uint64)
    # multiples of total and their neighbours are where an off-by-one would show
    k = rng.integers(1, ((1 << 64) - 1) // total, size=50_000, dtype=np.uint64)
    mult = k * np.uint64(total)
    allr = np.concatenate([r, edge, mult, mult - np.uint64(1), mult + np.uint64(1)])
    assert hostcore.hc_check_division(_p(allr), allr.size, total) == 0


@pytest.mark.parametrize("total", [3, 17, 255, 65537, 1000003, (1 << 30) - 125, (1 << 30) + 12345, (1 << 31) + 1,
                                   0xFFFFFFFF, 0xFFFFFFFE, 6, 1 << 20])
def test_divide_free_rpt_is_exact_or_flagged(hostcore, total):
    """fused_rpt_cs (general totals without a divide on the chain): whenever its exactness test passes the
    quotient equals floor((rpt * c << sh) / total); the cases it flags are rare for real tables.  Inputs:
    states as the coder produces them (range in [2^48, 2^64) before the symbol), plus quotients that sit
    exactly on a multiple of total (remainder 0 and total-1: where an off-by-one would show)."""
    rng = np.random.default_rng(total % 7919)
    n = 400_000
    rng_range = rng.integers(1 << 48, (1 << 64) - 1, size=n, dtype=np.uint64, endpoint=True)
    rpt = rng_range // np.uint64(total)
    c = rng.integers(1, max(2, total), size=n, dtype=np.uint64).astype(np.uint32)
    c[: n // 4] = rng.integers(1, min(total, 64), size=n // 4)          # rare symbols: the weakest margin
    c[n // 4: n // 2] = total - rng.integers(1, min(total, 64), size=n // 4)  # dominant symbols
    sh = (rng.integers(0, 4, size=n) * 8).astype(np.uint32)
    # keep only shifts the coder could take: (rpt * c) << sh < 2^64
    prod = rpt.astype(object) * c.astype(object)
    ok = np.array([(int(p) << int(s)) < (1 << 64) for p, s in zip(prod[:20000], sh[:20000])])
    sh[:20000][~ok] = 0
    sh[20000:] = 0
    # quotients on a multiple of total: rpt * c == k * total (+ 0 / - 1) needs rpt a multiple of total/gcd
    k = rng.integers(1, 1 << 20, size=1000, dtype=np.uint64)
    rpt_edge = k * np.uint64(total)
    rpts = np.concatenate([rpt, rpt_edge, rpt_edge + np.uint64(1), rpt_edge - np.uint64(1)])
    cs = np.concatenate([c, c[:1000], c[:1000], c[:1000]])
    shs = np.concatenate([sh, np.zeros(3000, np.uint32)])
    inexact = ctypes.c_uint64()
    wrong = hostcore.hc_check_cs(_p(rpts), _p(cs), _p(shs), rpts.size, total, ctypes.byref(inexact))
    assert wrong == 0
    if total & (total - 1):  # quotients on an exact multiple of total are (rightly) sent to the exact path
        assert inexact.value > 0
    # coder-like states and symbols a 4096-bucket table resolves (c >= total / 4096; a symbol's share of
    # flagged steps is ~ 2^-8 / c): far below the loop-2 rate (7e-4)
    keep = c >= max(1, total >> 12)
    r2, c2, s2 = (np.ascontiguousarray(a[keep]) for a in (rpt, c, sh))
    wrong = hostcore.hc_check_cs(_p(r2), _p(c2), _p(s2), r2.size, total, ctypes.byref(inexact))
    assert wrong == 0
    if total >= 1 << 20:
        assert inexact.value < max(3, r2.size * 2e-5)


@pytest.mark.parametrize("total", [(1 << 25) + 1, (1 << 26) - 1, (1 << 30) - 125, (1 << 30) + 12345, (1 << 31) + 1,
                                   0xFFFFFFFF, 0xFFFFFFFE, 3 << 29, 1_000_000_007])
def test_table_wide_reciprocal_is_exact_or_flagged(hostcore, total):
    """fused_rpt_m2 (decoder, totals >= 2^25): a passing test comes with floor((range' << sh) / total);
    inputs are post-symbol ranges at every legal shift plus ranges on exact multiples of total."""
    rng = np.random.default_rng(total % 7919)
    n = 400_000
    sh = (rng.integers(0, 4, size=n) * 8).astype(np.uint32)
    x = rng.integers(1 << 48, (1 << 64) - 1, size=n, dtype=np.uint64, endpoint=True)  # range' << sh
    rgp = x >> sh.astype(np.uint64)
    k = rng.integers(1 << 17, ((1 << 64) - 1) // total, size=3000, dtype=np.uint64)
    mult = k * np.uint64(total)
    rgps = np.concatenate([rgp, mult, mult - np.uint64(1), mult + np.uint64(1)])
    shs = np.concatenate([sh, np.zeros(9000, np.uint32)])
    inexact = ctypes.c_uint64()
    assert hostcore.hc_check_m2(_p(rgps), _p(shs), rgps.size, total, ctypes.byref(inexact)) == 0
    w = hostcore.hc_check_m2(_p(rgp), _p(sh), n, total, ctypes.byref(inexact))
    assert w == 0
    l = total.bit_length() - 1
    # flagged share: 2^-(l - sh) per symbol, uniform sh here -> dominated by sh = 24
    assert inexact.value < n * (2.0 ** -(l - 24)) * 0.5 + 50
