#!/usr/bin/env python
"""Small end-to-end pass over every kernel family for compute-sanitizer (memcheck): tiny inputs, all paths --
shared / per-chunk / general-total / K=4096 models, restart points (intact and damaged), TMA input, the adaptive-per-symbol table, corrupt offsets,
host-buffer pipeline, container.  `compute-sanitizer --tool memcheck python tools/sanitize_case.py`"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import range_coder_rust_b200 as rcb  # noqa: E402


def restart_syms(chunk):
    """A restart spacing for the chunk (multiple of 64, several parts), 0 when the chunk is too short."""
    return 64 * max(1, chunk // 64 // 5) if chunk >= 128 else 0


def main():
    ctx = rcb.Context(0)
    rng = np.random.default_rng(1)
    for K, sb, chunk, n in ((256, 1, 4096, 70 * 4096 + 333), (256, 1, 64, 64 * 200), (4096, 2, 2048, 2048 * 40 + 17),
                            (10, 1, 1000, 12345)):
        syms = ctx.generate(n, K, 0x5EED0001, rcb.zipf_thresholds(K, 1.1), sym_bytes=sb)
        counts = ctx.histogram(syms, K)
        counts += 1
        variants = [("hist", counts.clone())]
        c2 = counts.clone()
        c2[0] += (1 << 26) + 12345
        variants.append(("general_m2", c2))
        c3 = counts.clone()
        c3[0] += 999
        variants.append(("general_small", c3))
        c4 = counts.clone()
        c4[0] += (1 << 28) - int(c4.sum().item())
        variants.append(("pow2", c4))
        for name, cnt in variants:
            for tma in ("", "1"):
                if tma:
                    os.environ["RCB_ENC_TMA"] = "1"
                else:
                    os.environ.pop("RCB_ENC_TMA", None)
                model = ctx.model_from_counts(cnt)
                stream, offsets, nbytes = ctx.encode_chunks(syms, chunk, model)
                back = ctx.decode_chunks(stream, offsets, n, chunk, model, sym_bytes=sb)
                assert torch.equal(back, syms), (K, name, tma)
                rs = restart_syms(chunk)
                if rs:  # restart points: several decoder lanes per chunk, then damaged records (status only)
                    n_chunks = (n + chunk - 1) // chunk
                    rp = ctx.restart_points(n_chunks, chunk, rs)
                    s2, o2, nb2 = ctx.encode_chunks(syms, chunk, model, restart_syms=rs, restart=rp)
                    assert nb2 == nbytes and torch.equal(s2[:nb2], stream[:nbytes]), (K, name, tma, "restart stream")
                    assert torch.equal(ctx.decode_chunks(s2, o2, n, chunk, model, sym_bytes=sb, restart_syms=rs,
                                                         restart=rp), syms), (K, name, tma, "restart")
                    bad = rp.clone()
                    bad[2::3] += 1 << 20  # code_bytes far past the chunk
                    bad[0::7] ^= 0x5A5A5A5A5A
                    try:
                        ctx.decode_chunks(s2, o2, n, chunk, model, sym_bytes=sb, restart_syms=rs, restart=bad)
                    except rcb.RcbError:
                        pass
        os.environ.pop("RCB_ENC_TMA", None)
        # per-chunk tables
        pm = ctx.model_from_counts(ctx.histogram(syms, K, chunk_syms=chunk))
        stream, offsets, nbytes = ctx.encode_chunks(syms, chunk, pm)
        assert torch.equal(ctx.decode_chunks(stream, offsets, n, chunk, pm, sym_bytes=sb), syms)
        rs = restart_syms(chunk)
        if rs:
            rp = ctx.restart_points((n + chunk - 1) // chunk, chunk, rs)
            s2, o2, nb2 = ctx.encode_chunks(syms, chunk, pm, restart_syms=rs, restart=rp)
            assert nb2 == nbytes and torch.equal(s2[:nb2], stream[:nbytes])
            assert torch.equal(ctx.decode_chunks(s2, o2, n, chunk, pm, sym_bytes=sb, restart_syms=rs, restart=rp), syms)
        # corrupt offsets: contained, status only
        offs = offsets.clone()
        offs[1] = offs[2]
        offs[3] = nbytes + (1 << 40)
        try:
            ctx.decode_chunks(stream[: (nbytes + 15) // 16 * 16].clone(), offs, n, chunk, pm, sym_bytes=sb)
        except rcb.RcbError:
            pass
        # garbage stream
        junk = torch.randint(0, 256, (stream.numel(),), dtype=torch.uint8, device=ctx.device)
        try:
            ctx.decode_chunks(junk, offsets, n, chunk, pm, sym_bytes=sb)
        except rcb.RcbError:
            pass
        # adaptive-per-symbol table
        if K <= 1024:
            a_s, a_o, a_n = ctx.adaptive_encode_chunks(syms, chunk, K, 24, 60000)
            assert torch.equal(ctx.adaptive_decode_chunks(a_s, a_o, n, chunk, K, 24, 60000, sym_bytes=sb), syms)
        # host-buffer entry points + container
        h = syms.cpu().numpy()
        if sb == 2:
            h = h.view(np.uint16)
        model = ctx.model_from_counts(counts)
        out, off2, nb2 = ctx.encode_host(h, chunk, model)
        assert np.array_equal(ctx.decode_host(np.concatenate([out[:nb2], np.zeros(32, np.uint8)]), off2, n, chunk, model,
                                              sym_bytes=sb), h)
        frame = ctx.frame_encode(h, chunk, model)
        assert np.array_equal(ctx.frame_decode(frame), h)
        print("ok", K, sb, chunk, n, flush=True)
    ctx.close()
    print("sanitize_case: done")


if __name__ == "__main__":
    main()
