"""Regenerates tests/golden/vectors.json.

The reference crate ships no golden vectors and cannot be executed here (no
Rust toolchain), so these vectors come from oracle/rc_pyref.py -- the
type-by-type Python transliteration of the crate -- NOT from the reference
itself ("parity unpinned", see oracle/rc_oracle.h).  The `sample_impl` vector
additionally matches the value derived by hand in SURVEY.md App. B.1.

Run from the repository root:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import rc_pyref  # noqa: E402


def table_from_counts(counts):
    t = rc_pyref.FreqTable(len(counts))
    t.c = [int(x) for x in counts]
    t.calc_cum()
    return t


def vec(name, symbols, counts, note=""):
    t = table_from_counts(counts)
    code = rc_pyref.encode(symbols, t.c, t.cum, t.total_freq())
    dec = rc_pyref.decode(code, len(symbols), t.c, t.cum, t.total_freq())
    assert dec == list(symbols)
    v = {
        "name": name,
        "note": note,
        "K": len(counts),
        "c": t.c,
        "cum": t.cum,
        "total": t.total_freq(),
        "n_symbols": len(symbols),
        "code_len": len(code),
        "code_sha256": hashlib.sha256(code).hexdigest(),
    }
    if len(symbols) >= 128:
        # restart points (include/rcb200.h): the Encoder in front of symbol j = restart_syms, 2*restart_syms, ...
        # -- RangeCoder::lower_bound(), ::range() (src/range_coder.rs:28-35) and peek_code().len()
        # (src/encoder.rs:15-17), straight from the transliterated Encoder; hex because they are u64
        rs = 64 * -(-len(symbols) // (64 * 64))
        enc = rc_pyref.Encoder()
        recs = []
        for j, sym in enumerate(symbols):
            if j and j % rs == 0:
                recs.append(["%016x" % enc.range_coder.lower_bound, "%016x" % enc.range_coder.range,
                             len(enc.peek_code())])
            enc.encode(t, int(sym))
        assert bytes(enc.finish()) == bytes(code)
        v["restart"] = {"restart_syms": rs, "records": recs,
                        "note": "[lower_bound, range, code bytes so far] in front of symbol (r+1)*restart_syms"}
    if len(symbols) <= 64:
        v["symbols"] = [int(s) for s in symbols]
    elif len(set(symbols)) == 1:
        v["symbols_repeat"] = [int(symbols[0]), len(symbols)]
    if len(code) <= 256:
        v["code_hex"] = code.hex()
    return v


def main():
    out = []
    # 1. examples/sample_impl.rs:72-128 as shipped
    data = [2, 1, 1, 4, 1, 4, 2, 1, 0, 1, 5, 9, 8, 7, 6, 5]
    counts = np.bincount(data, minlength=10)
    v = vec("sample_impl", data, counts, "examples/sample_impl.rs:74,77")
    assert v["code_hex"] == "64475f8970365a2f83b20246c0", v["code_hex"]
    out.append(v)
    # 2. empty input: finish() alone (src/encoder.rs:40-46)
    out.append(vec("empty", [], [1, 1], "no symbols: 8 zero bytes"))
    # 3. single-symbol alphabet (c == total): nothing but the flush for a long time
    out.append(vec("single_symbol_10000", [0] * 10000, [7], "c == total"))
    # 4. totals that are not powers of two, incl. the u32 maximum (exercises the reciprocal)
    rng = np.random.default_rng(12345)
    syms = rng.integers(0, 3, size=48).tolist()
    out.append(vec("total_u32_max", syms, [1, 0xFFFFFFFF - 3, 2], "total = 2^32-1, c=1 symbols cost ~4 bytes"))
    syms = rng.integers(0, 5, size=64).tolist()
    out.append(vec("total_1000003", syms, [1, 999999, 1, 1, 1], "prime total"))
    # 5. zero-frequency symbols inside the alphabet (never coded)
    syms = [0, 2, 2, 5, 0, 5, 5, 2] * 6
    out.append(vec("zero_freq_gaps", syms, [12, 0, 18, 0, 0, 18, 0], "symbols 1,3,4,6 have c=0"))
    # 6. a 4096-symbol Zipf-ish chunk, K=256 (pinned by hash)
    w = np.arange(1, 257, dtype=np.float64) ** -1.1
    p = w / w.sum()
    syms = rng.choice(256, size=4096, p=p)
    counts = np.bincount(syms, minlength=256)
    v = vec("zipf_k256_4096", syms.tolist(), counts, "numpy default_rng(12345) continued; see symbols_hex")
    v["symbols_hex"] = bytes(int(s) for s in syms).hex()
    out.append(v)
    # 7. K=4096, u16 symbols
    w = np.arange(1, 4097, dtype=np.float64) ** -1.1
    p = w / w.sum()
    syms = rng.choice(4096, size=2048, p=p)
    counts = np.bincount(syms, minlength=4096)
    v = vec("zipf_k4096_2048", syms.tolist(), counts, "u16 symbols; see symbols_hex (little-endian u16)")
    v["symbols_hex"] = np.asarray(syms, dtype="<u2").tobytes().hex()
    v.pop("c")
    v.pop("cum")
    v["c_from_symbols"] = True
    out.append(v)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vectors.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path, len(out), "vectors")
    # the table follows the symbols (rc_pyref.AdaptiveFreqTable); separate file, same pinning
    arng = np.random.default_rng(4242)
    adaptive = []
    for name, K, inc, limit, n in (("adaptive_k16", 16, 8, 200, 300), ("adaptive_k256", 256, 24, 60000, 3000),
                                   ("adaptive_k3_tight", 3, 1, 3, 64), ("adaptive_k1000_u16", 1000, 24, 65000, 1500)):
        w = np.arange(1, K + 1, dtype=np.float64) ** -1.3
        syms = arng.choice(K, size=n, p=w / w.sum())
        syms[n // 2:] = K - 1 - syms[n // 2:]
        code = rc_pyref.adaptive_encode(syms.tolist(), K, inc, limit)
        assert rc_pyref.adaptive_decode(code, n, K, inc, limit) == syms.tolist()
        adaptive.append({"name": name, "K": K, "adaptive": {"inc": inc, "limit": limit}, "n_symbols": n,
                         "symbols_hex": (np.asarray(syms, dtype="<u2").tobytes() if K > 256
                                         else bytes(int(x) for x in syms)).hex(),
                         "code_len": len(code), "code_sha256": hashlib.sha256(code).hexdigest(),
                         "code_hex": code.hex() if len(code) <= 512 else None})
    apath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "adaptive_vectors.json")
    with open(apath, "w") as f:
        json.dump(adaptive, f, indent=1)
    print("wrote", apath, len(adaptive), "vectors")


if __name__ == "__main__":
    main()
