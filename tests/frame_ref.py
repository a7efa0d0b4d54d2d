"""Checker-side reader/writer of the RCB2 container (include/rcb200.h), independent of the
library's C++ implementation: numpy slicing + the oracle's chunk coder.  Test infrastructure."""
import struct

import numpy as np

import oracle_bind as oracle

HDR = struct.Struct("<4sIIIIIQQQQ")  # 56 bytes


def _al8(x):
    return (x + 7) & ~7


def read_frame(frame):
    frame = np.ascontiguousarray(frame, dtype=np.uint8)
    magic, ver, sb, K, mode, r_units, chunk, n, n_chunks, payload = HDR.unpack(frame[:HDR.size].tobytes())
    assert magic == b"RCB2" and ver in (1, 2) and (ver == 1) == (r_units == 0)
    off = HDR.size
    if mode == 0:
        total = int(frame[off:off + 4].view("<u4")[0])
        cum = frame[off + 8:off + 8 + 4 * K].view("<u4").copy()
        c = frame[off + 8 + 4 * K:off + 8 + 8 * K].view("<u4").copy()
        off += _al8(8 + 8 * K)
        models = [(c, cum, total)]
    else:
        cc = frame[off:off + 4 * K * n_chunks].view("<u4").reshape(n_chunks, K)
        models = []
        for j in range(n_chunks):
            cum, total = oracle.calc_cum(cc[j].copy())
            models.append((cc[j].copy(), cum, total))
        off += _al8(4 * K * n_chunks)
    offsets = frame[off:off + 8 * (n_chunks + 1)].view("<u8").copy()
    off += 8 * (n_chunks + 1)
    restart_syms, restart = 64 * r_units, None
    if restart_syms:  # version 2: rcb_restart_point[n_chunks][per] = {lower u64, range u64, code_bytes u32, 0 u32}
        per = (chunk + restart_syms - 1) // restart_syms - 1
        restart = frame[off:off + 24 * n_chunks * per].view("<u8").reshape(n_chunks, per, 3).copy()
        off += 24 * n_chunks * per
    stream = frame[off:off + payload].copy()
    return dict(sym_bytes=sb, K=K, mode=mode, chunk=chunk, n=n, n_chunks=n_chunks, models=models, offsets=offsets,
                stream=stream, restart_syms=restart_syms, restart=restart)


def decode_frame(frame):
    f = read_frame(frame)
    dt = np.uint8 if f["sym_bytes"] == 1 else np.uint16
    out = np.empty(f["n"], dtype=dt)
    for j in range(f["n_chunks"]):
        c, cum, total = f["models"][0 if f["mode"] == 0 else j]
        lo, hi = j * f["chunk"], min(f["n"], (j + 1) * f["chunk"])
        code = f["stream"][int(f["offsets"][j]):int(f["offsets"][j + 1])]
        out[lo:hi] = oracle.decode(code, hi - lo, c, cum, total, f["sym_bytes"])[0]
    return out


def restart_records(part, restart_syms, per, c, cum, total):
    """The restart points of one chunk from the oracle's Encoder state (rco_encode_state): lower_bound, range
    rounded down to a multiple of total_freq, bytes emitted so far; absent records (ragged chunk) are zero."""
    rec = np.zeros((per, 3), dtype="<u8")
    for r in range(per):
        j = (r + 1) * restart_syms
        if j >= part.size:
            break
        lo, rg, nb = oracle.encode_state(part[:j], c, cum, total)
        rec[r] = (lo, rg // total * total, nb)
    return rec


def write_frame(syms, chunk, models, K, restart_syms=0):
    """models: one (c, cum, total) or one per chunk.  restart_syms != 0: version-2 frame with restart points."""
    syms = np.ascontiguousarray(syms)
    n = syms.size
    n_chunks = (n + chunk - 1) // chunk
    mode = 0 if len(models) == 1 else 1
    parts, offsets, recs = [], [0], []
    per = (chunk + restart_syms - 1) // restart_syms - 1 if restart_syms else 0
    if per <= 0:
        restart_syms = per = 0
    for j in range(n_chunks):
        c, cum, total = models[0 if mode == 0 else j]
        code = oracle.encode(syms[j * chunk:(j + 1) * chunk], c, cum, total)
        parts.append(code)
        offsets.append(offsets[-1] + len(code))
        if per:
            recs.append(restart_records(syms[j * chunk:(j + 1) * chunk], restart_syms, per, c, cum, int(total)))
    body = bytearray(HDR.pack(b"RCB2", 2 if per else 1, syms.dtype.itemsize, K, mode, restart_syms // 64, chunk, n,
                              n_chunks, offsets[-1]))
    if mode == 0:
        c, cum, total = models[0]
        ms = struct.pack("<II", total, 0) + np.asarray(cum, "<u4").tobytes() + np.asarray(c, "<u4").tobytes()
    else:
        ms = b"".join(np.asarray(m[0], "<u4").tobytes() for m in models)
    body += ms + b"\0" * (_al8(len(ms)) - len(ms))
    body += np.asarray(offsets, "<u8").tobytes()
    for rec in recs:
        body += rec.tobytes()
    for p in parts:
        body += p
    return np.frombuffer(bytes(body), dtype=np.uint8)
