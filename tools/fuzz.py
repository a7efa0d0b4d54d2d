#!/usr/bin/env python
"""Randomised parity run on the GPU: random alphabets, tables (dense, sparse, power-of-two and odd
totals), stream lengths and chunk sizes (odd, tiny, ragged), shared and per-chunk models, with and without
restart points -- every chunk's bytes (and restart records) against the oracle, and the decoded symbols against
the input.  Not part of the test
suite (minutes, not seconds); run it after touching a kernel:  python tools/fuzz.py --iters 200"""
import argparse
import os
import sys
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_bind as oracle  # noqa: E402
import range_coder_rust_b200 as rcb  # noqa: E402


def random_counts(rng, K, kind):
    if kind == "zipf":
        w = np.arange(1, K + 1, dtype=np.float64) ** -rng.uniform(0.0, 4.0)
        rng.shuffle(w)
    elif kind == "uniform":
        w = np.ones(K)
    elif kind == "sparse":
        w = rng.random(K) ** 4
        w[rng.random(K) < 0.7] = 0
        w[rng.integers(0, K)] += 1.0
    else:  # spiky
        w = rng.random(K) * 1e-4
        w[rng.integers(0, K, size=max(1, K // 16))] += 1.0
    return w / w.sum()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    ctx = rcb.Context(0)
    rng = np.random.default_rng(a.seed)
    for it in range(a.iters):
        K = int(rng.choice([2, 3, 10, 64, 255, 256, 257, 1000, 4096]))
        sb = 1 if K <= 256 else 2
        kind = str(rng.choice(["zipf", "uniform", "sparse", "spiky"]))
        p = random_counts(rng, K, kind)
        n = int(rng.choice([1, 5, 1000, 65536, 200_000, 1_000_003]))
        chunk = int(rng.choice([1, 7, 100, 4096, 16384, 65536, 65537, 99_999]))
        big = rng.random() < 0.03  # enough chunks and bytes for the sliced / segmented host pipeline
        if big:
            n, chunk = 40_000_003, int(rng.choice([32768, 65536]))
        if n // chunk > 20000:
            chunk = 4096
        per_chunk = bool(rng.random() < 0.4) and n // chunk <= 2000
        syms = rng.choice(K, size=n, p=p).astype(np.uint8 if sb == 1 else np.uint16)
        d = torch.from_numpy(syms.view(np.int16) if sb == 2 else syms).to(ctx.device)
        n_chunks = (n + chunk - 1) // chunk
        tag = f"it={it} K={K} kind={kind} n={n} chunk={chunk} per_chunk={per_chunk}"
        if per_chunk:
            model = ctx.model_from_counts(ctx.histogram(d, K, chunk_syms=chunk))
        else:
            mode = str(rng.choice(["hist", "pow2", "odd"]))
            counts = ctx.histogram(d, K)
            if mode == "hist":
                model = ctx.model_from_counts(counts)
            else:  # host-made table over the symbols that occur: power-of-two or odd total
                cnt = counts.cpu().numpy().astype(np.float64)
                total = (1 << int(rng.integers(12, 32))) if mode == "pow2" else int(rng.integers(1 << 12, (1 << 32) - 1))
                nz = cnt > 0
                if total < 2 * int(nz.sum()):
                    total = 4 * int(nz.sum())
                c = np.zeros(K, dtype=np.int64)
                c[nz] = np.maximum(1, np.floor(cnt[nz] / cnt.sum() * (total - nz.sum())).astype(np.int64))
                c[np.argmax(cnt)] += total - c.sum()
                assert c.min() >= 0 and c.sum() == total
                c = c.astype(np.uint32)
                cum, tot = oracle.calc_cum(c)
                model = ctx.model_from_tables(c, cum, tot)
            tag += f" table={mode}"
        stream, offsets, nbytes = ctx.encode_chunks(d, chunk, model)
        h_stream = stream.cpu().numpy()[:nbytes]
        h_off = offsets.cpu().numpy().astype(np.uint64)
        pick = range(n_chunks) if n_chunks <= 64 else sorted(set(rng.integers(0, n_chunks, size=48).tolist()) | {0, n_chunks - 1})
        for j in pick:
            c, cum, total, _ = model.tables(j if per_chunk else 0)
            ref = oracle.encode(syms[j * chunk:(j + 1) * chunk], c, cum, total)
            got = h_stream[int(h_off[j]):int(h_off[j + 1])].tobytes()
            assert got == ref, f"encode mismatch chunk {j}: {tag}"
        back = ctx.decode_chunks(stream, offsets, n, chunk, model, sym_bytes=sb)
        b = back.cpu().numpy()
        b = b.view(np.uint16) if sb == 2 else b
        assert np.array_equal(b, syms), f"decode mismatch: {tag}"
        # restart points: same bytes, records == the oracle's Encoder state, several decoder lanes per chunk
        units = chunk // 64
        if units >= 2 and rng.random() < 0.6:
            lo_u = -(-units // 64)  # at most 64 parts per chunk
            rs = 64 * int(rng.integers(lo_u, max(lo_u, units // 2) + 1))
            per = (chunk + rs - 1) // rs - 1
            restart = ctx.restart_points(n_chunks, chunk, rs)
            if restart is not None:
                s2, o2, nb2 = ctx.encode_chunks(d, chunk, model, restart_syms=rs, restart=restart)
                assert nb2 == nbytes and torch.equal(o2, offsets) and torch.equal(s2[:nb2], stream[:nbytes]), \
                    f"restart encode changed the stream: {tag} rs={rs}"
                rec = restart.cpu().numpy().view(np.uint64).reshape(n_chunks, per, 3)
                for j in list(pick)[:6]:
                    c, cum, total, _ = model.tables(j if per_chunk else 0)
                    part = syms[j * chunk:(j + 1) * chunk]
                    for r in sorted(set(rng.integers(0, per, size=min(per, 4)).tolist())):
                        k = (r + 1) * rs
                        if k >= part.size:
                            assert not rec[j, r].any(), f"absent record not zero: {tag} rs={rs} chunk {j} rec {r}"
                            continue
                        lo, rg, nb_ = oracle.encode_state(part[:k], c, cum, int(total))
                        assert (int(rec[j, r, 0]), int(rec[j, r, 1]), int(rec[j, r, 2]) & 0xFFFFFFFF) == \
                            (lo, rg // int(total) * int(total), nb_), f"restart record: {tag} rs={rs} chunk {j} rec {r}"
                st = torch.full((n_chunks,), 99, dtype=torch.int32, device=ctx.device)
                back2 = ctx.decode_chunks(s2, o2, n, chunk, model, sym_bytes=sb, status=st, restart_syms=rs, restart=restart)
                b2 = back2.cpu().numpy()
                b2 = b2.view(np.uint16) if sb == 2 else b2
                assert np.array_equal(b2, syms) and not st.any(), f"restart decode mismatch: {tag} rs={rs}"
                tag += f" restart={rs}"
                if big or rng.random() < 0.2:  # host entry points with restart points
                    rs_np = np.zeros(n_chunks * per * 3, dtype=np.uint64)
                    o, off2, nbh, _ = ctx.encode_host(syms, chunk, model, restart_syms=rs, restart_np=rs_np)
                    assert nbh == nbytes and o[:nbh].tobytes() == h_stream.tobytes(), tag
                    assert np.array_equal(rs_np.reshape(n_chunks, per, 3), rec), tag
                    pad = np.zeros(nbh + 32, dtype=np.uint8)
                    pad[:nbh] = o[:nbh]
                    assert np.array_equal(ctx.decode_host(pad, off2, n, chunk, model, sym_bytes=sb, restart_syms=rs,
                                                          restart_np=rs_np), syms), tag
        if big or rng.random() < 0.3:  # host entry points (pipeline + segments when big enough)
            o, off2, nb2 = ctx.encode_host(syms, chunk, model)
            assert nb2 == nbytes and np.array_equal(off2, h_off) and o[:nb2].tobytes() == h_stream.tobytes(), tag
            pad = np.zeros(nb2 + 32, dtype=np.uint8)
            pad[:nb2] = o[:nb2]
            assert np.array_equal(ctx.decode_host(pad, off2, n, chunk, model, sym_bytes=sb), syms), tag
        print(f"ok {tag} bytes={nbytes} crc={zlib.crc32(h_stream.tobytes()):08x}", flush=True)
    print("fuzz: all iterations bit-exact")


if __name__ == "__main__":
    main()
