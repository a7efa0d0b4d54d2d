#!/usr/bin/env python
"""Timeline of the pipelined end-to-end leg of bench.py: per-call wall times of rcb_encode_host (thread E,
context 1) and rcb_decode_host (thread D, context 2) when step i+1's encode overlaps step i's decode."""
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
    p.add_argument("--stagger-ms", type=float, default=0.0)
    p.add_argument("--enc-delay-ms", type=float, default=0.0, help="start encode i+1 this long after decode i was handed over")
    a = p.parse_args()
    n, chunk, K = int(a.gib * (1 << 30)), 65536, 256
    dev = torch.device("cuda:0")
    ctx = rcb.Context(0)
    ctx2 = rcb.Context(0, stream=torch.cuda.Stream(dev))
    d = ctx.generate(n, K, 0x5EED0001, rcb.zipf_thresholds(K, 1.1))
    model = ctx.model_from_counts(ctx.histogram(d, K))
    c, cum, total, _ = model.tables()
    model2 = ctx2.model_from_tables(c, cum, total)
    cap = ctx.encode_bound(model, n, 1, chunk) + 16
    h_syms = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_syms.copy_(d)
    bufs = [torch.empty(cap, dtype=torch.uint8, pin_memory=True).numpy() for _ in range(2)]
    h_back = torch.empty(n, dtype=torch.uint8, pin_memory=True).numpy()
    syms = h_syms.numpy()
    log = []
    t00 = [0.0]

    def run(k_steps):
        q = queue.Queue()
        free = threading.Semaphore(2)

        def dec():
            with torch.cuda.device(dev):
                while True:
                    it = q.get()
                    if it is None:
                        return
                    i, buf, offs, nb = it
                    if a.stagger_ms:
                        time.sleep(a.stagger_ms * 1e-3)
                    t0 = time.perf_counter()
                    ctx2.decode_host(buf[:nb], offs, n, chunk, model2, out_np=h_back)
                    log.append(("D", i, (t0 - t00[0]) * 1e3, (time.perf_counter() - t00[0]) * 1e3))
                    free.release()

        t = threading.Thread(target=dec)
        t.start()
        for i in range(k_steps):
            free.acquire()
            if i and a.enc_delay_ms:
                time.sleep(a.enc_delay_ms * 1e-3)
            t0 = time.perf_counter()
            buf, offs, nb = ctx.encode_host(syms, chunk, model, out_np=bufs[i & 1])
            log.append(("E", i, (t0 - t00[0]) * 1e3, (time.perf_counter() - t00[0]) * 1e3))
            q.put((i, buf, offs, nb))
        q.put(None)
        t.join()

    run(2)
    log.clear()
    torch.cuda.synchronize()
    t00[0] = time.perf_counter()
    run(a.steps)
    dt = time.perf_counter() - t00[0]
    assert np.array_equal(h_back, syms)
    for who, i, t0, t1 in sorted(log, key=lambda x: x[2]):
        print(f"{who}{i}: {t0:8.2f} -> {t1:8.2f}  ({t1 - t0:6.2f} ms)")
    print(json.dumps({"steps": a.steps, "ms_per_step": dt / a.steps * 1e3, "gbs": n * a.steps / dt / 1e9}))


if __name__ == "__main__":
    main()
