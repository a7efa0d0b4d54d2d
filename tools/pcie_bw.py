#!/usr/bin/env python
"""Host<->device copy bandwidth of this box (pinned memory) next to the host-buffer entry points:
the bus is the roofline of `e2e` in bench.py."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import range_coder_rust_b200 as rcb  # noqa: E402


def wall(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    return best


def main():
    n = 1 << 30
    dev = torch.device("cuda:0")
    h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    res["h2d_gbs"] = n / wall(lambda: d_a.copy_(h_a, non_blocking=True)) / 1e9
    res["d2h_gbs"] = n / wall(lambda: h_b.copy_(d_b, non_blocking=True)) / 1e9

    def both():
        with torch.cuda.stream(s1):
            d_a.copy_(h_a, non_blocking=True)
        with torch.cuda.stream(s2):
            h_b.copy_(d_b, non_blocking=True)

    res["bidir_each_gbs"] = n / wall(both) / 1e9

    def sliced():
        k = 8
        for i in range(k):
            sl = slice(i * n // k, (i + 1) * n // k)
            with torch.cuda.stream(s1 if i % 2 == 0 else s2):
                d_a[sl].copy_(h_a[sl], non_blocking=True)
                h_b[sl].copy_(d_a[sl], non_blocking=True)

    res["sliced_roundtrip_gbs_each"] = n / wall(sliced) / 1e9

    ctx = rcb.Context(0)
    K, chunk = 256, 65536
    d = ctx.generate(n, K, 0x5EED0001, rcb.zipf_thresholds(K, 1.1))
    model = ctx.model_from_counts(ctx.histogram(d, K))
    h_a.copy_(d)
    syms = h_a.numpy()
    cap = ctx.encode_bound(model, n, 1, chunk) + 16
    h_stream = torch.empty(cap, dtype=torch.uint8).pin_memory().numpy()
    out = h_b.numpy()
    st, offs, nbytes = ctx.encode_host(syms, chunk, model, out_np=h_stream)
    res["compressed_bytes"] = nbytes
    res["encode_host_ms"] = wall(lambda: ctx.encode_host(syms, chunk, model, out_np=h_stream)) * 1e3
    res["decode_host_ms"] = wall(lambda: ctx.decode_host(h_stream, offs, n, chunk, model, out_np=out)) * 1e3
    res["round_trip_ok"] = bool(np.array_equal(out, syms))
    res["encode_host_bus_gbs"] = (n + nbytes) / res["encode_host_ms"] / 1e6
    res["decode_host_bus_gbs"] = (n + nbytes) / res["decode_host_ms"] / 1e6
    print(json.dumps(res))


if __name__ == "__main__":
    main()
