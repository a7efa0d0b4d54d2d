#!/usr/bin/env python
"""Timeline of the pipelined end-to-end leg of bench.py: per-call wall times of rcb_encode_host (E threads, one
context each) and rcb_decode_host (D threads, one context each) when the encodes of later steps overlap the
decodes of earlier ones.  Every step copies its input host->device and its results device->host inside its calls;
--enc / --dec set how many calls of each kind may be in flight."""
import argparse
import json
import os
import queue
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import range_coder_rust_b200 as rcb  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--steps", type=int, default=6)
    p.add_argument("--gib", type=float, default=1.0)
    p.add_argument("--enc", type=int, default=1, help="encoder threads (contexts)")
    p.add_argument("--dec", type=int, default=1, help="decoder threads (contexts)")
    p.add_argument("--restart", type=int, default=0, help="restart points every N symbols (0: none)")
    p.add_argument("--quiet", action="store_true")
    a = p.parse_args()
    n, chunk, K = int(a.gib * (1 << 30)), 65536, 256
    dev = torch.device("cuda:0")
    ctx0 = rcb.Context(0)
    d = ctx0.generate(n, K, 0x5EED0001, rcb.zipf_thresholds(K, 1.1))
    model0 = ctx0.model_from_counts(ctx0.histogram(d, K))
    c, cum, total, _ = model0.tables()
    cap = ctx0.encode_bound(model0, n, 1, chunk) + 16
    h_syms = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_syms.copy_(d)
    syms = h_syms.numpy()
    del d

    def make_ctx():
        cx = rcb.Context(0, stream=torch.cuda.Stream(dev))
        return cx, cx.model_from_tables(c, cum, total)

    encs = [make_ctx() for _ in range(a.enc)]
    decs = [make_ctx() for _ in range(a.dec)]
    n_buf = a.enc + a.dec
    n_chunks = n // chunk
    per = (chunk + a.restart - 1) // a.restart - 1 if a.restart else 0
    bufs = [torch.empty(cap, dtype=torch.uint8, pin_memory=True).numpy() for _ in range(n_buf)]
    rsb = [torch.zeros(max(1, n_chunks * per * 3), dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
           for _ in range(n_buf)]
    backs = [torch.empty(n, dtype=torch.uint8, pin_memory=True).numpy() for _ in range(a.dec)]
    log = []
    t00 = [0.0]

    def run(k_steps):
        q = queue.Queue()
        free = queue.Queue()
        for b in range(n_buf):
            free.put(b)
        nxt = [0]
        lock = threading.Lock()

        def dec(j):
            cx, m = decs[j]
            with torch.cuda.device(dev):
                while True:
                    it = q.get()
                    if it is None:
                        return
                    i, b, offs, nb = it
                    t0 = time.perf_counter()
                    cx.decode_host(bufs[b][:nb], offs, n, chunk, m, out_np=backs[j], restart_syms=a.restart,
                                   restart_np=rsb[b] if per else None)
                    log.append(("D%d" % j, i, (t0 - t00[0]) * 1e3, (time.perf_counter() - t00[0]) * 1e3))
                    free.put(b)

        def enc(j):
            cx, m = encs[j]
            with torch.cuda.device(dev):
                while True:
                    with lock:
                        i = nxt[0]
                        nxt[0] += 1
                    if i >= k_steps:
                        return
                    b = free.get()
                    t0 = time.perf_counter()
                    r = cx.encode_host(syms, chunk, m, out_np=bufs[b], restart_syms=a.restart if per else 0,
                                       restart_np=rsb[b] if per else None)
                    log.append(("E%d" % j, i, (t0 - t00[0]) * 1e3, (time.perf_counter() - t00[0]) * 1e3))
                    q.put((i, b, r[1], r[2]))

        dts = [threading.Thread(target=dec, args=(j,)) for j in range(a.dec)]
        ets = [threading.Thread(target=enc, args=(j,)) for j in range(a.enc)]
        for t in dts + ets:
            t.start()
        for t in ets:
            t.join()
        for _ in dts:
            q.put(None)
        for t in dts:
            t.join()

    run(max(2, n_buf))
    log.clear()
    torch.cuda.synchronize()
    t00[0] = time.perf_counter()
    run(a.steps)
    dt = time.perf_counter() - t00[0]
    for b in backs:
        assert np.array_equal(b, syms)
    if not a.quiet:
        for who, i, t0, t1 in sorted(log, key=lambda x: x[2]):
            print(f"{who} step {i}: {t0:8.2f} -> {t1:8.2f}  ({t1 - t0:6.2f} ms)")
    print(json.dumps({"enc": a.enc, "dec": a.dec, "restart": a.restart, "steps": a.steps,
                      "ms_per_step": dt / a.steps * 1e3, "gbs": n * a.steps / dt / 1e9}))


if __name__ == "__main__":
    main()
