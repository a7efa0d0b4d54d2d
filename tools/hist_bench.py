#!/usr/bin/env python
"""Times rcb_histogram on the BASELINE.json shapes (1 GiB, device resident, CUDA events, best of --reps) and
checks the counts against torch.bincount.  RCB_HIST_SHARED=0 selects the per-warp-copy kernels for comparison
(the variable is read once per process: run it twice)."""
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


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--bytes", type=int, default=1 << 30)
    p.add_argument("--reps", type=int, default=5)
    a = p.parse_args()
    ctx = rcb.Context(0)
    out = {"RCB_HIST_SHARED": os.environ.get("RCB_HIST_SHARED", "(default)")}
    for name, K, chunk, adaptive in (("global_u8_K256", 256, 0, False), ("global_u8_K200", 200, 0, False),
                                     ("global_u16_K4096", 4096, 0, False), ("chunks_u8_64KiB", 256, 65536, True),
                                     ("chunks_u8_16KiB", 256, 16384, True), ("chunks_u16_K4096_32Ki", 4096, 32768, False)):
        sb = 1 if K <= 256 else 2
        n = a.bytes // sb
        if adaptive:
            thr = np.stack([rcb.zipf_thresholds(K, s) for s in S_CYCLE])
            d = ctx.generate(n, K, 0x5EED0002, thr, sym_bytes=sb, chunk_syms=chunk)
        else:
            d = ctx.generate(n, K, 0x5EED0001, rcb.zipf_thresholds(K, 1.1), sym_bytes=sb)
        counts = ctx.histogram(d, K, chunk_syms=chunk)
        ms = timed(lambda: ctx.histogram(d, K, chunk_syms=chunk, out=counts), a.reps)
        flat = d.view(torch.int16).to(torch.int32) & 0xFFFF if sb == 2 else d
        if chunk:
            nck = 64  # the first chunks against bincount
            ref = torch.stack([torch.bincount(flat[i * chunk:(i + 1) * chunk].to(torch.int64), minlength=K) for i in range(nck)])
            ok = bool((counts.view(-1, K)[:nck].to(torch.int64) == ref).all())
        else:
            ref = torch.bincount(flat.to(torch.int64), minlength=K)
            ok = bool((counts.to(torch.int64) == ref).all())
        out[name] = {"ms": round(ms, 4), "gbs": round(a.bytes / ms / 1e6, 1), "ok": ok}
        del d, counts, flat, ref
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
