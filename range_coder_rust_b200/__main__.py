"""File front end of the RCB2 container: bytes of a file are the symbols (K = 256).

  python -m range_coder_rust_b200 compress   IN OUT [--chunk 65536] [--adaptive] [--restart N] [--device 0]
  python -m range_coder_rust_b200 decompress IN OUT
  python -m range_coder_rust_b200 info       IN

`compress` builds the frequency table on the GPU (one table for the file, or one per chunk with
--adaptive), codes every chunk as one reference Encoder run and writes the frame; `decompress` needs
nothing but the frame.  All coding happens in librcb200.so; there is no CPU path."""
import argparse
import sys

import numpy as np
import torch

from . import Context


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m range_coder_rust_b200")
    sub = ap.add_subparsers(dest="cmd", required=True)
    c = sub.add_parser("compress")
    c.add_argument("src")
    c.add_argument("dst")
    c.add_argument("--chunk", type=int, default=65536, help="symbols per chunk")
    c.add_argument("--adaptive", action="store_true", help="one frequency table per chunk")
    c.add_argument("--restart", type=int, default=-1,
                   help="restart points every N symbols of a chunk (several decoder lanes per chunk; multiple of 64); "
                        "0 = none (version-1 frame), default: chunk / 16 for chunks of >= 32 Ki symbols, else chunk / 4")
    c.add_argument("--device", type=int, default=0)
    d = sub.add_parser("decompress")
    d.add_argument("src")
    d.add_argument("dst")
    d.add_argument("--device", type=int, default=0)
    i = sub.add_parser("info")
    i.add_argument("src")
    i.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)

    ctx = Context(a.device)
    if a.cmd == "compress":
        syms = np.fromfile(a.src, dtype=np.uint8)
        if syms.size == 0:
            sys.exit("empty input: nothing to model")
        dev = torch.from_numpy(syms).to(ctx.device)
        counts = ctx.histogram(dev, 256, chunk_syms=a.chunk if a.adaptive else 0)
        model = ctx.model_from_counts(counts)
        restart = a.restart if a.restart >= 0 else (a.chunk // 16 if a.chunk >= 32768 and a.chunk % 1024 == 0 else
                                                    (a.chunk // 4 if a.chunk % 256 == 0 else 0))
        frame = ctx.frame_encode(syms, a.chunk, model, restart_syms=restart)
        frame.tofile(a.dst)
        print(f"{a.src}: {syms.size} -> {frame.size} bytes ({frame.size / syms.size:.4f}), "
              f"{(syms.size + a.chunk - 1) // a.chunk} chunks, {'per-chunk tables' if a.adaptive else 'one table'}"
              f"{', restart points every %d symbols' % restart if restart else ''}")
    elif a.cmd == "decompress":
        frame = np.fromfile(a.src, dtype=np.uint8)
        out = ctx.frame_decode(frame)
        out.tofile(a.dst)
        print(f"{a.src}: {frame.size} -> {out.nbytes} bytes")
    else:
        frame = np.fromfile(a.src, dtype=np.uint8)
        f = ctx.frame_info(frame)
        print(f"RCB2 v{f.version}: {f.n_syms} symbols of {f.sym_bytes} byte(s), K={f.K}, {f.n_chunks} chunks of "
              f"{f.chunk_syms}, {'per-chunk tables' if f.model_mode else 'one table'}, payload {f.payload_bytes} bytes, "
              f"frame {f.frame_bytes} bytes"
              f"{', restart points every %d symbols' % f.restart_syms if f.restart_syms else ''}")


if __name__ == "__main__":
    main()
