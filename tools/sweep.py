#!/usr/bin/env python
"""Throughput of every BASELINE.json configuration shape on one GPU (device-resident data,
CUDA events, best of --reps).  The headline number is bench.py's; this sweep documents the
other rows of SURVEY 8(d): adaptive per-chunk models over a chunk sweep, the 4096-symbol
alphabet, non-power-of-two totals and a many-lane batch."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import range_coder_rust_b200 as rcb  # noqa: E402

S_CYCLE = (0.0, 0.25, 0.5, 0.8, 1.1, 1.5, 2.0, 3.0, 5.0)


def timed(fn, reps):
    best = None
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        t = a.elapsed_time(b)
        best = t if best is None else min(best, t)
    return best


def run(ctx, name, n_bytes, K, chunk, mode, zipf, reps, odd_total=False, parts=(1,)):
    sb = 1 if K <= 256 else 2
    n = n_bytes // sb
    if mode == "adaptive":
        thr = np.stack([rcb.zipf_thresholds(K, s) for s in S_CYCLE])
        d = ctx.generate(n, K, 0x5EED0002, thr, sym_bytes=sb, chunk_syms=chunk)
    else:
        d = ctx.generate(n, K, 0x5EED0001 if K <= 256 else 0x5EED0003, rcb.zipf_thresholds(K, zipf), sym_bytes=sb)
    n_chunks = (n + chunk - 1) // chunk
    res = {"config": name, "bytes": n_bytes, "K": K, "chunk_syms": chunk, "model": mode, "lanes": n_chunks}
    if mode == "adaptive":
        counts = ctx.histogram(d, K, chunk_syms=chunk)
        model = ctx.model_from_counts(counts)
        res["hist_ms"] = timed(lambda: ctx.histogram(d, K, chunk_syms=chunk, out=counts), reps)
        res["model_ms"] = timed(lambda: ctx.model_from_counts(counts, model=model), reps)
    else:
        counts = ctx.histogram(d, K)
        if odd_total:  # a static table whose total is not a power of two (reciprocal path)
            counts[0] += 12345
        model = ctx.model_from_counts(counts)
        res["hist_ms"] = timed(lambda: ctx.histogram(d, K, out=counts), reps)
    c, cum, total, flags = model.tables(0)
    res["total_model0"] = int(total)
    cap = ctx.encode_bound(model, n, sb, chunk) + 16
    stream = torch.empty(cap, dtype=torch.uint8, device=ctx.device)
    offsets = torch.empty(n_chunks + 1, dtype=torch.int64, device=ctx.device)
    back = torch.empty_like(d)
    ctx.encode_chunks(d, chunk, model, out=stream, offsets=offsets)
    res["encode_ms"] = timed(lambda: ctx.encode_chunks(d, chunk, model, out=stream, offsets=offsets, sync=False), reps)
    nbytes = ctx.encode_result()
    res["ratio"] = nbytes / n_bytes
    ctx.decode_chunks(stream, offsets, n, chunk, model, sym_bytes=sb, out=back)
    res["decode_ms"] = timed(
        lambda: ctx.decode_chunks(stream, offsets, n, chunk, model, sym_bytes=sb, out=back, sync=False), reps)
    ctx.decode_result()
    res["round_trip_ok"] = bool(torch.equal(back, d))
    res["encode_gbs"] = n_bytes / res["encode_ms"] / 1e6
    res["decode_gbs"] = n_bytes / res["decode_ms"] / 1e6
    # restart points: `p` decoder lanes per chunk (the encoder records its state every chunk / p symbols)
    for p in parts:
        if p <= 1 or chunk % (64 * p):
            continue
        rs = chunk // p
        pts = ctx.restart_points(n_chunks, chunk, rs)
        kw = {"restart_syms": rs, "restart": pts}
        ctx.encode_chunks(d, chunk, model, out=stream, offsets=offsets, **kw)
        enc = timed(lambda: ctx.encode_chunks(d, chunk, model, out=stream, offsets=offsets, sync=False, **kw), reps)
        ctx.encode_result()
        back.zero_()
        ctx.decode_chunks(stream, offsets, n, chunk, model, sym_bytes=sb, out=back, **kw)
        dec = timed(lambda: ctx.decode_chunks(stream, offsets, n, chunk, model, sym_bytes=sb, out=back, sync=False,
                                              **kw), reps)
        ctx.decode_result()
        res[f"restart_x{p}"] = {"restart_syms": rs, "encode_ms": enc, "decode_ms": dec,
                                "decode_gbs": n_bytes / dec / 1e6, "round_trip_ok": bool(torch.equal(back, d))}
    print(json.dumps(res), flush=True)
    del d, stream, offsets, back, model
    torch.cuda.empty_cache()


def run_f4(ctx, n_bytes, chunk, reps, inc=24, limit=60000):
    """the table follows the symbols (counts from 1, +inc, halving at limit), restart per chunk."""
    K = 256
    thr = np.stack([rcb.zipf_thresholds(K, s) for s in S_CYCLE])
    d = ctx.generate(n_bytes, K, 0x5EED0002, thr, sym_bytes=1, chunk_syms=chunk)
    n_chunks = (n_bytes + chunk - 1) // chunk
    stream, offsets, nbytes = ctx.adaptive_encode_chunks(d, chunk, K, inc, limit)
    res = {"config": f"f4: adaptive-per-symbol table, mixed entropy, K=256, chunk {chunk // 1024} KiB", "bytes": n_bytes,
           "K": K, "chunk_syms": chunk, "model": "per-symbol", "lanes": n_chunks, "ratio": nbytes / n_bytes}
    res["encode_ms"] = timed(lambda: ctx.adaptive_encode_chunks(d, chunk, K, inc, limit, out=stream, offsets=offsets), reps)
    back = torch.empty_like(d)
    res["decode_ms"] = timed(lambda: ctx.adaptive_decode_chunks(stream, offsets, n_bytes, chunk, K, inc, limit, out=back), reps)
    res["round_trip_ok"] = bool(torch.equal(back, d))
    res["encode_gbs"] = n_bytes / res["encode_ms"] / 1e6
    res["decode_gbs"] = n_bytes / res["decode_ms"] / 1e6
    print(json.dumps(res), flush=True)


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--reps", type=int, default=3)
    p.add_argument("--gib", type=float, default=1.0)
    p.add_argument("--big", action="store_true", help="also an 8 GiB many-lane batch (config 5's per-GPU shard)")
    p.add_argument("--parts", default="1", help="comma-separated decoder lanes per chunk to add (restart points)")
    p.add_argument("--only", default="", help="comma-separated config tags to run (2,3-16,3-64,3-256,4,2b,2c,f4)")
    a = p.parse_args()
    only = set(x for x in a.only.split(",") if x)
    parts = tuple(int(x) for x in a.parts.split(",") if x)

    def want(tag):
        return not only or tag in only
    ctx = rcb.Context(0)
    nb = int(a.gib * (1 << 30))
    if want("2"):
        run(ctx, "2: static global table, Zipf 1.1, K=256", nb, 256, 65536, "static", 1.1, a.reps, parts=parts)
    for chunk in (16384, 65536, 262144):
        if want(f"3-{chunk // 1024}"):
            run(ctx, f"3: adaptive per-chunk, mixed entropy, K=256, chunk {chunk // 1024} KiB", nb, 256, chunk,
                "adaptive", None, a.reps, parts=parts)
    if want("4"):
        run(ctx, "4: K=4096 (u16), Zipf 1.1, static global table", nb, 4096, 32768, "static", 1.1, a.reps, parts=parts)
    if want("2b"):
        run(ctx, "2b: static table with a non-power-of-two total", nb, 256, 65536, "static", 1.1, a.reps,
            odd_total=True, parts=parts)
    if want("2c"):
        run(ctx, "2c: static global table, 16 KiB chunks (65536 lanes)", nb, 256, 16384, "static", 1.1, a.reps, parts=parts)
    if want("f4"):
        run_f4(ctx, nb, 65536, a.reps)
    if a.big:
        run(ctx, "5: 8 GiB per-GPU shard, static table, 64 KiB chunks (131072 lanes)", 8 << 30, 256, 65536, "static",
            1.1, max(1, a.reps - 1), parts=parts)


if __name__ == "__main__":
    main()
