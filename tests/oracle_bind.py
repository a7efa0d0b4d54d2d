"""ctypes binding of the CPU oracle (oracle/librc_oracle.so) -- test infrastructure.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
legs use this module; the product package never imports it.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "librc_oracle.so")

_vp = ctypes.c_void_p
_u64 = ctypes.c_uint64
_u32 = ctypes.c_uint32
_i64 = ctypes.c_int64
_ci = ctypes.c_int

RCO_ERR = {
    -1: "LOWER_OVERFLOW", -2: "UPPER_OVERFLOW", -3: "ZERO_TOTAL", -4: "ZERO_FREQ",
    -5: "CAPACITY", -6: "TRUNCATED", -7: "SYMBOL_RANGE",
}


def build_oracle(force=False):
    src = [os.path.join(ORACLE_DIR, f) for f in ("rc_oracle.c", "rc_oracle.h", "Makefile")]
    if force or not os.path.exists(ORACLE_LIB) or any(
            os.path.getmtime(s) > os.path.getmtime(ORACLE_LIB) for s in src):
        subprocess.run(["make", "-C", ORACLE_DIR, "-B"], check=True, capture_output=True)
    return ORACLE_LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        L = ctypes.CDLL(ORACLE_LIB)
        L.rco_encode.restype = _i64
        L.rco_encode.argtypes = [_vp, _u64, _ci, _u32, _vp, _vp, _u32, _vp, _u64]
        L.rco_encode_state.restype = _i64
        L.rco_encode_state.argtypes = [_vp, _u64, _ci, _u32, _vp, _vp, _u32, _vp, _vp]
        L.rco_decode.restype = _i64
        L.rco_decode.argtypes = [_vp, _u64, _u64, _ci, _u32, _vp, _vp, _u32, _vp]
        L.rco_histogram.restype = None
        L.rco_histogram.argtypes = [_vp, _u64, _ci, _u32, _vp]
        L.rco_calc_cum.restype = _u32
        L.rco_calc_cum.argtypes = [_vp, _u32, _vp]
        L.rco_normalise.restype = _ci
        L.rco_normalise.argtypes = [_vp, _u32, _vp]
        L.rco_encode_chunks.restype = _ci
        L.rco_encode_chunks.argtypes = [_vp, _u64, _ci, _u64, _u32, _vp, _vp, _vp, _ci, _vp, _u64, _vp, _ci]
        L.rco_decode_chunks.restype = _ci
        L.rco_decode_chunks.argtypes = [_vp, _vp, _u64, _ci, _u64, _u32, _vp, _vp, _vp, _ci, _vp, _vp, _ci]
        L.rco_adaptive_encode.restype = _i64
        L.rco_adaptive_encode.argtypes = [_vp, _u64, _ci, _u32, _u32, _u32, _vp, _u64]
        L.rco_adaptive_decode.restype = _i64
        L.rco_adaptive_decode.argtypes = [_vp, _u64, _u64, _ci, _u32, _u32, _u32, _vp]
        L.rco_adaptive_encode_chunks.restype = _ci
        L.rco_adaptive_encode_chunks.argtypes = [_vp, _u64, _ci, _u64, _u32, _u32, _u32, _vp, _u64, _vp, _ci]
        L.rco_adaptive_decode_chunks.restype = _ci
        L.rco_adaptive_decode_chunks.argtypes = [_vp, _vp, _u64, _ci, _u64, _u32, _u32, _u32, _vp, _vp, _ci]
        L.rco_generate.restype = None
        L.rco_generate.argtypes = [_vp, _u64, _u64, _ci, _u32, _u64, _vp, _u32, _u64, _ci]
        L.rco_hardware_threads.restype = _ci
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_vp)


def hardware_threads():
    return int(lib().rco_hardware_threads())


def histogram(syms, K):
    syms = np.ascontiguousarray(syms)
    counts = np.zeros(K, dtype=np.uint64)
    lib().rco_histogram(_p(syms), syms.size, syms.dtype.itemsize, K, _p(counts))
    return counts


def normalise(counts):
    counts = np.ascontiguousarray(counts, dtype=np.uint64)
    c = np.zeros(counts.size, dtype=np.uint32)
    scaled = lib().rco_normalise(_p(counts), counts.size, _p(c))
    return c, scaled


def calc_cum(c):
    c = np.ascontiguousarray(c, dtype=np.uint32)
    cum = np.zeros(c.size, dtype=np.uint32)
    total = lib().rco_calc_cum(_p(c), c.size, _p(cum))
    return cum, int(total)


def model_from_symbols(syms, K):
    """histogram -> (identity or shift) normalise -> calc_cum; returns (c, cum, total)."""
    c, _ = normalise(histogram(syms, K))
    cum, total = calc_cum(c)
    return c, cum, total


def encode(syms, c, cum, total, cap=None):
    """Whole Encoder run; returns bytes, or raises ValueError(name) on an oracle error."""
    syms = np.ascontiguousarray(syms)
    c = np.ascontiguousarray(c, dtype=np.uint32)
    cum = np.ascontiguousarray(cum, dtype=np.uint32)
    if cap is None:
        cap = 16 * syms.size + 64
    out = np.zeros(cap, dtype=np.uint8)
    n = lib().rco_encode(_p(syms), syms.size, syms.dtype.itemsize, c.size, _p(c), _p(cum), total, _p(out), cap)
    if n < 0:
        raise ValueError(RCO_ERR.get(n, str(n)))
    return out[:n].tobytes()


def encode_state(syms, c, cum, total):
    """Encoder state after `syms` and before finish(): (lower_bound, range, bytes emitted so far)."""
    syms = np.ascontiguousarray(syms)
    c = np.ascontiguousarray(c, dtype=np.uint32)
    cum = np.ascontiguousarray(cum, dtype=np.uint32)
    lo = np.zeros(1, dtype=np.uint64)
    rg = np.zeros(1, dtype=np.uint64)
    n = lib().rco_encode_state(_p(syms), syms.size, syms.dtype.itemsize, c.size, _p(c), _p(cum), total, _p(lo), _p(rg))
    if n < 0:
        raise ValueError(RCO_ERR.get(n, str(n)))
    return int(lo[0]), int(rg[0]), int(n)


def decode(code, n_syms, c, cum, total, sym_bytes=1):
    code = np.frombuffer(bytes(code), dtype=np.uint8) if not isinstance(code, np.ndarray) else code
    code = np.ascontiguousarray(code, dtype=np.uint8)
    c = np.ascontiguousarray(c, dtype=np.uint32)
    cum = np.ascontiguousarray(cum, dtype=np.uint32)
    out = np.zeros(n_syms, dtype=np.uint8 if sym_bytes == 1 else np.uint16)
    used = lib().rco_decode(_p(code), code.size, n_syms, sym_bytes, c.size, _p(c), _p(cum), total, _p(out))
    if used < 0:
        raise ValueError(RCO_ERR.get(used, str(used)))
    return out, int(used)


def encode_chunks(syms, chunk_syms, c, cum, total, threads=None, pitch=None):
    """Chunked multi-thread encode; returns (stream bytes array, offsets uint64[n_chunks+1])."""
    syms = np.ascontiguousarray(syms)
    c = np.ascontiguousarray(c, dtype=np.uint32)
    cum = np.ascontiguousarray(cum, dtype=np.uint32)
    total = np.ascontiguousarray(np.atleast_1d(total), dtype=np.uint32)
    per_chunk = 1 if c.ndim == 2 else 0
    K = c.shape[-1]
    n = syms.size
    n_chunks = (n + chunk_syms - 1) // chunk_syms
    if pitch is None:
        pitch = 4 * chunk_syms * syms.dtype.itemsize + 64
    out = np.zeros(n_chunks * pitch, dtype=np.uint8)
    lens = np.zeros(n_chunks, dtype=np.int64)
    threads = threads or hardware_threads()
    bad = lib().rco_encode_chunks(_p(syms), n, syms.dtype.itemsize, chunk_syms, K, _p(c), _p(cum), _p(total),
                                  per_chunk, _p(out), pitch, _p(lens), threads)
    if bad:
        first = int(np.argmax(lens < 0))
        raise ValueError(f"{bad} chunks failed; chunk {first}: {RCO_ERR.get(int(lens[first]), lens[first])}")
    offsets = np.zeros(n_chunks + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lens)
    stream = np.empty(int(offsets[-1]), dtype=np.uint8)
    for i in range(n_chunks):
        stream[int(offsets[i]):int(offsets[i + 1])] = out[i * pitch:i * pitch + int(lens[i])]
    return stream, offsets


def decode_chunks(stream, offsets, n_syms, chunk_syms, c, cum, total, sym_bytes=1, threads=None):
    stream = np.ascontiguousarray(stream, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    c = np.ascontiguousarray(c, dtype=np.uint32)
    cum = np.ascontiguousarray(cum, dtype=np.uint32)
    total = np.ascontiguousarray(np.atleast_1d(total), dtype=np.uint32)
    per_chunk = 1 if c.ndim == 2 else 0
    K = c.shape[-1]
    n_chunks = (n_syms + chunk_syms - 1) // chunk_syms
    out = np.zeros(n_syms, dtype=np.uint8 if sym_bytes == 1 else np.uint16)
    used = np.zeros(n_chunks, dtype=np.int64)
    threads = threads or hardware_threads()
    bad = lib().rco_decode_chunks(_p(stream), _p(offsets), n_syms, sym_bytes, chunk_syms, K, _p(c), _p(cum),
                                  _p(total), per_chunk, _p(out), _p(used), threads)
    if bad:
        first = int(np.argmax(used < 0))
        raise ValueError(f"{bad} chunks failed; chunk {first}: {RCO_ERR.get(int(used[first]), used[first])}")
    return out, used


def adaptive_encode(syms, K, inc, limit, cap=None):
    """f4: whole Encoder run over a table that follows the symbols (rc_oracle.h)."""
    syms = np.ascontiguousarray(syms)
    cap = cap or 16 * syms.size + 64
    out = np.zeros(cap, dtype=np.uint8)
    n = lib().rco_adaptive_encode(_p(syms), syms.size, syms.dtype.itemsize, K, inc, limit, _p(out), cap)
    if n < 0:
        raise ValueError(RCO_ERR.get(n, str(n)))
    return out[:n].tobytes()


def adaptive_decode(code, n_syms, K, inc, limit, sym_bytes=1):
    code = np.ascontiguousarray(np.frombuffer(bytes(code), dtype=np.uint8) if not isinstance(code, np.ndarray) else code)
    out = np.zeros(n_syms, dtype=np.uint8 if sym_bytes == 1 else np.uint16)
    used = lib().rco_adaptive_decode(_p(code), code.size, n_syms, sym_bytes, K, inc, limit, _p(out))
    if used < 0:
        raise ValueError(RCO_ERR.get(used, str(used)))
    return out, int(used)


def adaptive_encode_chunks(syms, chunk_syms, K, inc, limit, threads=None):
    """(stream, offsets) like encode_chunks; every chunk restarts coder and table."""
    syms = np.ascontiguousarray(syms)
    n = syms.size
    n_chunks = (n + chunk_syms - 1) // chunk_syms
    pitch = 4 * chunk_syms * syms.dtype.itemsize + 64
    out = np.zeros(n_chunks * pitch, dtype=np.uint8)
    lens = np.zeros(n_chunks, dtype=np.int64)
    bad = lib().rco_adaptive_encode_chunks(_p(syms), n, syms.dtype.itemsize, chunk_syms, K, inc, limit, _p(out), pitch,
                                           _p(lens), threads or hardware_threads())
    if bad:
        raise ValueError(f"{bad} chunks failed")
    offsets = np.zeros(n_chunks + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lens)
    stream = np.empty(int(offsets[-1]), dtype=np.uint8)
    for i in range(n_chunks):
        stream[int(offsets[i]):int(offsets[i + 1])] = out[i * pitch:i * pitch + int(lens[i])]
    return stream, offsets


def adaptive_decode_chunks(stream, offsets, n_syms, chunk_syms, K, inc, limit, sym_bytes=1, threads=None):
    stream = np.ascontiguousarray(stream, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n_chunks = (n_syms + chunk_syms - 1) // chunk_syms
    out = np.zeros(n_syms, dtype=np.uint8 if sym_bytes == 1 else np.uint16)
    used = np.zeros(n_chunks, dtype=np.int64)
    bad = lib().rco_adaptive_decode_chunks(_p(stream), _p(offsets), n_syms, sym_bytes, chunk_syms, K, inc, limit,
                                           _p(out), _p(used), threads or hardware_threads())
    if bad:
        raise ValueError(f"{bad} chunks failed")
    return out, used


def histogram_mt(syms, K, threads=None):
    """The same histogram with the symbols cut into one slice per host thread (ctypes releases the GIL
    during the C call); the u64 partial tables are summed.  Used by the timed CPU baselines so that the
    histogram does not run single-threaded next to multi-threaded coding."""
    from concurrent.futures import ThreadPoolExecutor

    syms = np.ascontiguousarray(syms)
    threads = max(1, min(threads or hardware_threads(), (syms.size + (1 << 16) - 1) >> 16))
    if threads == 1:
        return histogram(syms, K)
    cuts = [syms.size * t // threads for t in range(threads + 1)]
    with ThreadPoolExecutor(threads) as ex:
        parts = list(ex.map(lambda t: histogram(syms[cuts[t]:cuts[t + 1]], K), range(threads)))
    return np.sum(parts, axis=0, dtype=np.uint64)


class RoundTripBuffers:
    """Preallocated staging / output arrays for repeated timed passes over the same batch shape: the
    timed window then holds only the reference algorithm (histogram, table, encode, decode), no numpy
    allocation or stream concatenation."""

    def __init__(self, n_syms, chunk_syms, sym_bytes=1):
        self.n, self.chunk, self.sb = n_syms, chunk_syms, sym_bytes
        self.n_chunks = (n_syms + chunk_syms - 1) // chunk_syms
        self.pitch = 4 * chunk_syms * sym_bytes + 64
        self.staging = np.zeros(self.n_chunks * self.pitch, dtype=np.uint8)
        self.lens = np.zeros(self.n_chunks, dtype=np.int64)
        self.offsets = (np.arange(self.n_chunks + 1, dtype=np.uint64) * np.uint64(self.pitch))
        self.back = np.zeros(n_syms, dtype=np.uint8 if sym_bytes == 1 else np.uint16)
        self.used = np.zeros(self.n_chunks, dtype=np.int64)

    def encode(self, syms, c, cum, total, threads):
        c = np.ascontiguousarray(c, dtype=np.uint32)
        cum = np.ascontiguousarray(cum, dtype=np.uint32)
        total = np.ascontiguousarray(np.atleast_1d(total), dtype=np.uint32)
        bad = lib().rco_encode_chunks(_p(syms), self.n, self.sb, self.chunk, c.shape[-1], _p(c), _p(cum), _p(total),
                                      1 if c.ndim == 2 else 0, _p(self.staging), self.pitch, _p(self.lens), threads)
        if bad:
            raise ValueError(f"{bad} chunks failed")

    def decode(self, c, cum, total, threads):
        """Decodes every chunk from its staging row (row i starts at i * pitch; a chunk reads no byte
        past its own length, so the row pitch serves as the offset table)."""
        c = np.ascontiguousarray(c, dtype=np.uint32)
        cum = np.ascontiguousarray(cum, dtype=np.uint32)
        total = np.ascontiguousarray(np.atleast_1d(total), dtype=np.uint32)
        bad = lib().rco_decode_chunks(_p(self.staging), _p(self.offsets), self.n, self.sb, self.chunk, c.shape[-1],
                                      _p(c), _p(cum), _p(total), 1 if c.ndim == 2 else 0, _p(self.back),
                                      _p(self.used), threads)
        if bad:
            raise ValueError(f"{bad} chunks failed")
        return self.back

    def stream(self):
        """(stream, offsets) in the dense layout of encode_chunks (not timed)."""
        offsets = np.zeros(self.n_chunks + 1, dtype=np.uint64)
        offsets[1:] = np.cumsum(self.lens)
        stream = np.empty(int(offsets[-1]), dtype=np.uint8)
        for i in range(self.n_chunks):
            stream[int(offsets[i]):int(offsets[i + 1])] = self.staging[i * self.pitch:i * self.pitch + int(self.lens[i])]
        return stream, offsets


def generate(n, K, seed, thresholds, sym_bytes=1, chunk_syms=0, first=0, threads=None):
    thr = np.ascontiguousarray(thresholds, dtype=np.uint32)
    if thr.ndim == 1:
        thr = thr[None, :]
    out = np.zeros(n, dtype=np.uint8 if sym_bytes == 1 else np.uint16)
    lib().rco_generate(_p(out), first, n, sym_bytes, K, seed, _p(thr), thr.shape[0], chunk_syms or 1,
                       threads or hardware_threads())
    return out


def zipf_thresholds(K, s):
    """Same formula as range_coder_rust_b200.api.zipf_thresholds (kept separate on purpose:
    the oracle side must not import the product)."""
    w = np.arange(1, K + 1, dtype=np.float64) ** (-float(s))
    cdf = np.cumsum(w) / np.sum(w)
    thr = np.floor(cdf[: K - 1] * 4294967296.0)
    return np.minimum(thr, 4294967295.0).astype(np.uint32)
