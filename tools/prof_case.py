#!/usr/bin/env python
"""One BASELINE.json configuration, coded twice on one GPU (first pass = warm-up): the command ncu
captures.  `ncu -k regex:"encode_kernel|decode" -s 2 -c 2 python tools/prof_case.py --case 2b`
profiles the second pass's encode and decode kernels."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import range_coder_rust_b200 as rcb  # noqa: E402

S_CYCLE = (0.0, 0.25, 0.5, 0.8, 1.1, 1.5, 2.0, 3.0, 5.0)


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--case", default="2", choices=["2", "2b", "3-16", "3-64", "3-256", "4"])
    p.add_argument("--gib", type=float, default=1.0)
    p.add_argument("--restart", type=int, default=0, help="restart points every this many symbols (0 = none)")
    a = p.parse_args()
    ctx = rcb.Context(0)
    nb = int(a.gib * (1 << 30))
    K = 4096 if a.case == "4" else 256
    sb = 2 if K > 256 else 1
    n = nb // sb
    if a.case.startswith("3"):
        chunk = int(a.case.split("-")[1]) * 1024
        thr = np.stack([rcb.zipf_thresholds(K, s) for s in S_CYCLE])
        d = ctx.generate(n, K, 0x5EED0002, thr, sym_bytes=sb, chunk_syms=chunk)
        model = ctx.model_from_counts(ctx.histogram(d, K, chunk_syms=chunk))
    else:
        chunk = 65536 // sb
        d = ctx.generate(n, K, 0x5EED0001 if K <= 256 else 0x5EED0003, rcb.zipf_thresholds(K, 1.1), sym_bytes=sb)
        counts = ctx.histogram(d, K)
        if a.case == "2b":
            counts[0] += 12345
        model = ctx.model_from_counts(counts)
    n_chunks = (n + chunk - 1) // chunk
    cap = ctx.encode_bound(model, n, sb, chunk) + 16
    stream = torch.empty(cap, dtype=torch.uint8, device=ctx.device)
    offsets = torch.empty(n_chunks + 1, dtype=torch.int64, device=ctx.device)
    back = torch.empty_like(d)
    restart = ctx.restart_points(n_chunks, chunk, a.restart) if a.restart else None
    kw = {"restart_syms": a.restart, "restart": restart} if restart is not None else {}
    for _ in range(2):
        ctx.encode_chunks(d, chunk, model, out=stream, offsets=offsets, **kw)
        ctx.decode_chunks(stream, offsets, n, chunk, model, sym_bytes=sb, out=back, **kw)
    assert torch.equal(back, d)
    print("ok", a.case)


if __name__ == "__main__":
    main()
