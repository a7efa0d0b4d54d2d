"""Thin Python front end over the C ABI: torch supplies device memory and
streams (plumbing); every byte of coding work happens in librcb200.so.

Names follow the reference's domain: symbols, chunks, frequency tables
(`c_freq`, `cum_freq`, `total_freq` -- src/pmodel.rs:4-13), code streams.
"""
import ctypes
import weakref

import numpy as np
import torch

from . import _lib
from ._lib import RcbError


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _np_ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


class FrameInfo(ctypes.Structure):
    """struct rcb_frame_info (include/rcb200.h)."""
    _fields_ = [(n, ctypes.c_uint32) for n in ("version", "sym_bytes", "K", "model_mode")] + \
               [(n, ctypes.c_uint64) for n in ("chunk_syms", "n_syms", "n_chunks", "payload_bytes", "model_off",
                                               "offsets_off", "payload_off", "frame_bytes", "restart_syms",
                                               "restart_off")]


class Model:
    """Dense snapshot of a PModel (src/pmodel.rs:4-13): one table shared by all
    chunks (n_models == 1) or one per chunk."""

    def __init__(self, ctx, K, n_models=1):
        self.ctx = ctx
        self.K = int(K)
        self.n_models = int(n_models)
        h = ctypes.c_void_p()
        ctx._check(ctx.lib.rcb_model_create(ctx.h, self.K, self.n_models, ctypes.byref(h)), "rcb_model_create")
        self.h = h
        ctx._models.add(self)

    def close(self):
        # a model must be destroyed before its context (the C side keeps a ctx pointer)
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.rcb_model_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def tables(self, index=0):
        """(c[K], cum[K], total, flags) of model `index`, copied to the host."""
        c = np.zeros(self.K, dtype=np.uint32)
        cum = np.zeros(self.K, dtype=np.uint32)
        total = ctypes.c_uint32()
        flags = ctypes.c_uint32()
        self.ctx._check(
            self.ctx.lib.rcb_model_get_tables(self.ctx.h, self.h, index, _np_ptr(c), _np_ptr(cum),
                                              ctypes.byref(total), ctypes.byref(flags)),
            "rcb_model_get_tables")
        return c, cum, total.value, flags.value


class AdaptiveParams(ctypes.Structure):
    """struct rcb_adaptive_params (include/rcb200.h): the adaptive-per-symbol table of SURVEY 8 f4."""
    _fields_ = [("K", ctypes.c_uint32), ("inc", ctypes.c_uint32), ("limit", ctypes.c_uint32)]


class Comm:
    """An NCCL communicator owned by the library (include/rcb200.h, multi-GPU section): the path's only
    exchange step is `Context.allreduce_counts` over it."""

    def __init__(self, ctx, h, n_ranks, rank):
        self.ctx, self.h, self.n_ranks, self.rank = ctx, h, n_ranks, rank

    @staticmethod
    def unique_id(lib=None):
        """128 opaque bytes from rank 0 (ncclGetUniqueId); ship them to the other ranks out of band."""
        lib = lib or _lib.load()
        buf = (ctypes.c_uint8 * 128)()
        rc = lib.rcb_comm_unique_id(buf)
        if rc:
            raise RcbError(rc, "rcb_comm_unique_id", (lib.rcb_comm_last_error(None) or b"").decode())
        return bytes(buf)

    @property
    def nccl_version(self):
        v = ctypes.c_int()
        self.ctx.lib.rcb_comm_info(self.h, None, None, ctypes.byref(v))
        return v.value

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.rcb_comm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One context per GPU (include/rcb200.h); issues work on the torch stream
    that is current when it is created (or on `stream`)."""

    def __init__(self, device=0, stream=None):
        self.lib = _lib.load()
        self.h = None
        self._models = weakref.WeakSet()
        if not torch.cuda.is_available():
            raise RcbError(_lib.RCB_ERR_NO_DEVICE, "Context")
        self.device = torch.device("cuda", device)
        with torch.cuda.device(self.device):
            s = stream if stream is not None else torch.cuda.current_stream(self.device)
        self.stream = s
        h = ctypes.c_void_p()
        rc = self.lib.rcb_ctx_create(device, ctypes.c_void_p(s.cuda_stream), ctypes.byref(h))
        if rc:
            raise RcbError(rc, "rcb_ctx_create")
        self.h = h

    # ------------------------------------------------------------ plumbing
    def _check(self, rc, where):
        if rc:
            detail = ""
            if rc == _lib.RCB_ERR_CUDA:
                msg = ctypes.c_char_p()
                self.lib.rcb_last_cuda_error(self.h, ctypes.byref(msg))
                detail = (msg.value or b"").decode()
            raise RcbError(rc, where, detail)

    def close(self):
        if getattr(self, "h", None):
            for m in list(self._models):
                m.close()
            self.lib.rcb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        self._check(self.lib.rcb_ctx_synchronize(self.h), "rcb_ctx_synchronize")

    def set_block_threads(self, encode_threads=0, decode_threads=0):
        self._check(self.lib.rcb_ctx_set_block_threads(self.h, encode_threads, decode_threads),
                    "rcb_ctx_set_block_threads")

    def enable_timing(self, on=True):
        self._check(self.lib.rcb_ctx_enable_timing(self.h, 1 if on else 0), "rcb_ctx_enable_timing")

    def timings(self):
        """ms of (encode kernel, length scan, gather, decode kernel, status summary)."""
        ms = (ctypes.c_float * 5)()
        self._check(self.lib.rcb_ctx_get_timings(self.h, ms, 5), "rcb_ctx_get_timings")
        return dict(zip(("encode_kernel", "scan", "gather", "decode_kernel", "summary"), [float(x) for x in ms]))

    @property
    def launch_count(self):
        return int(self.lib.rcb_ctx_launch_count(self.h))

    @staticmethod
    def _sym_bytes(t):
        if t.dtype == torch.uint8:
            return 1
        if t.dtype in (torch.int16, torch.uint16):
            return 2
        raise TypeError("symbols must be uint8 or (u)int16 tensors")

    # --------------------------------------------------------------- model
    def histogram(self, syms, K, chunk_syms=0, out=None):
        """FreqTable::add_alphabet_freq loop (examples/sample_impl.rs:58-60,78-80).
        chunk_syms == 0 -> int64[K]; else int32[n_chunks][K] (bit patterns of u64/u32)."""
        n = syms.numel()
        sb = self._sym_bytes(syms)
        counts = out
        if counts is None:
            if chunk_syms == 0:
                counts = torch.empty(K, dtype=torch.int64, device=self.device)
            else:
                n_chunks = (n + chunk_syms - 1) // chunk_syms
                counts = torch.empty((n_chunks, K), dtype=torch.int32, device=self.device)
        self._check(self.lib.rcb_histogram(self.h, _ptr(syms), n, sb, K, chunk_syms, _ptr(counts)),
                    "rcb_histogram")
        return counts

    def model_from_counts(self, counts, K=None, model=None):
        """FreqTable::calc_cum (examples/sample_impl.rs:61-69) on device counts.
        `model` rebuilds an existing Model of the same shape in place."""
        if counts.dim() == 1:
            n_models, k = 1, counts.shape[0]
        else:
            n_models, k = counts.shape
        K = int(K or k)
        assert k == K
        cb = 8 if counts.dtype == torch.int64 else 4
        assert counts.dtype in (torch.int64, torch.int32)
        m = model if model is not None else Model(self, K, n_models)
        assert m.K == K and m.n_models == n_models
        self._check(self.lib.rcb_model_from_counts(self.h, m.h, _ptr(counts.contiguous()), cb),
                    "rcb_model_from_counts")
        return m

    def model_from_tables(self, c, cum, total):
        """Snapshot of PModel tables given as host arrays: c/cum [K] or [n_models][K]."""
        c = np.ascontiguousarray(c, dtype=np.uint32)
        cum = np.ascontiguousarray(cum, dtype=np.uint32)
        if c.ndim == 1:
            n_models, K = 1, c.shape[0]
        else:
            n_models, K = c.shape
        total = np.ascontiguousarray(np.atleast_1d(total), dtype=np.uint32)
        assert cum.shape == c.shape and total.shape[0] == n_models
        m = Model(self, K, n_models)
        self._check(self.lib.rcb_model_from_tables(self.h, m.h, _np_ptr(c), _np_ptr(cum), _np_ptr(total)),
                    "rcb_model_from_tables")
        return m

    # ------------------------------------------------------------ multi-GPU
    def comm_init_rank(self, unique_id, n_ranks, rank):
        """One process (or thread) per GPU: every rank calls this with rank 0's `Comm.unique_id()`."""
        assert len(unique_id) == 128
        buf = (ctypes.c_uint8 * 128).from_buffer_copy(unique_id)
        h = ctypes.c_void_p()
        rc = self.lib.rcb_comm_init_rank(self.h, buf, n_ranks, rank, ctypes.byref(h))
        if rc:
            raise RcbError(rc, "rcb_comm_init_rank", (self.lib.rcb_comm_last_error(None) or b"").decode())
        return Comm(self, h, n_ranks, rank)

    @staticmethod
    def comm_init_all(ctxs):
        """One thread driving several GPUs (ncclCommInitAll): one Comm per context, same order."""
        lib = ctxs[0].lib
        n = len(ctxs)
        hs = (ctypes.c_void_p * n)(*[c.h for c in ctxs])
        out = (ctypes.c_void_p * n)()
        rc = lib.rcb_comm_init_all(hs, n, out)
        if rc:
            raise RcbError(rc, "rcb_comm_init_all", (lib.rcb_comm_last_error(None) or b"").decode())
        return [Comm(c, ctypes.c_void_p(out[i]), n, i) for i, c in enumerate(ctxs)]

    def allreduce_counts(self, counts, comm):
        """Sum the ranks' u64 count tables in place (one ncclAllReduce on this context's stream)."""
        assert counts.dtype == torch.int64 and counts.dim() == 1 and counts.is_contiguous()
        rc = self.lib.rcb_allreduce_counts(self.h, comm.h, _ptr(counts), counts.numel())
        if rc:
            raise RcbError(rc, "rcb_allreduce_counts", (self.lib.rcb_comm_last_error(comm.h) or b"").decode())
        return counts

    @staticmethod
    def allreduce_counts_multi(ctxs, comms, counts_list):
        lib = ctxs[0].lib
        n = len(ctxs)
        K = counts_list[0].numel()
        assert all(t.dtype == torch.int64 and t.numel() == K and t.is_contiguous() for t in counts_list)
        hs = (ctypes.c_void_p * n)(*[c.h for c in ctxs])
        ks = (ctypes.c_void_p * n)(*[k.h for k in comms])
        ps = (ctypes.c_void_p * n)(*[t.data_ptr() for t in counts_list])
        rc = lib.rcb_allreduce_counts_multi(hs, ks, ps, K, n)
        if rc:
            raise RcbError(rc, "rcb_allreduce_counts_multi", (lib.rcb_comm_last_error(comms[0].h) or b"").decode())

    # -------------------------------------------------------- encode/decode
    def encode_bound(self, model, n_syms, sym_bytes, chunk_syms):
        return int(self.lib.rcb_encode_bound(self.h, model.h, n_syms, sym_bytes, chunk_syms))

    def restart_points(self, n_chunks, chunk_syms, restart_syms):
        """Device buffer for the restart points of a call: rcb_restart_point[n_chunks][per chunk] as
        int64[n_chunks * per_chunk * 3] (24-byte records), or None when the chunk has a single part."""
        per = int(self.lib.rcb_restart_points_per_chunk(chunk_syms, restart_syms))
        if per == 0:
            return None
        return torch.zeros(n_chunks * per * 3, dtype=torch.int64, device=self.device)

    def encode_chunks(self, syms, chunk_syms, model, out=None, offsets=None, status=None, sync=True,
                      restart_syms=0, restart=None):
        """Every chunk through Encoder::encode ... finish (src/encoder.rs:24-46).
        Returns (stream uint8[cap], offsets int64[n_chunks+1], n_bytes or None).
        restart_syms / restart (see restart_points): also record the coder state every restart_syms symbols."""
        n = syms.numel()
        sb = self._sym_bytes(syms)
        n_chunks = (n + chunk_syms - 1) // chunk_syms
        if out is None:
            out = torch.empty(self.encode_bound(model, n, sb, chunk_syms) + 16, dtype=torch.uint8,
                              device=self.device)
        if offsets is None:
            offsets = torch.empty(n_chunks + 1, dtype=torch.int64, device=self.device)
        rs = int(restart_syms) if restart is not None else 0
        if sync:
            nbytes = ctypes.c_uint64()
            rc = self.lib.rcb_encode_chunks_restart(self.h, _ptr(syms), n, sb, chunk_syms, model.h, _ptr(out),
                                                    out.numel(), _ptr(offsets), _ptr(status), rs, _ptr(restart),
                                                    ctypes.byref(nbytes))
            self._check(rc, "rcb_encode_chunks")
            return out, offsets, int(nbytes.value)
        rc = self.lib.rcb_encode_chunks_restart_async(self.h, _ptr(syms), n, sb, chunk_syms, model.h, _ptr(out),
                                                      out.numel(), _ptr(offsets), _ptr(status), rs, _ptr(restart))
        self._check(rc, "rcb_encode_chunks_async")
        return out, offsets, None

    def encode_result(self):
        nbytes = ctypes.c_uint64()
        self._check(self.lib.rcb_encode_result(self.h, ctypes.byref(nbytes)), "rcb_encode_result")
        return int(nbytes.value)

    def decode_chunks(self, stream, offsets, n_syms, chunk_syms, model, sym_bytes=1, out=None, status=None,
                      sync=True, restart_syms=0, restart=None):
        """Every chunk through Decoder::new + decode (src/decoder.rs:14-54); with restart points
        (restart_syms, restart as written by encode_chunks) several lanes per chunk."""
        if out is None:
            dt = torch.uint8 if sym_bytes == 1 else torch.int16
            out = torch.empty(n_syms, dtype=dt, device=self.device)
        rs = int(restart_syms) if restart is not None else 0
        fn = self.lib.rcb_decode_chunks_restart if sync else self.lib.rcb_decode_chunks_restart_async
        rc = fn(self.h, _ptr(stream), _ptr(offsets), n_syms, sym_bytes, chunk_syms, model.h, _ptr(out),
                _ptr(status), rs, _ptr(restart))
        self._check(rc, "rcb_decode_chunks")
        return out

    def decode_result(self):
        self._check(self.lib.rcb_decode_result(self.h), "rcb_decode_result")

    # ------------------------------------------- adaptive-per-symbol model (f4)
    def adaptive_encode_chunks(self, syms, chunk_syms, K, inc=24, limit=60000, out=None, offsets=None, status=None):
        """Every chunk through Encoder::encode under a table the caller updates after each symbol
        (counts from 1, +inc for the coded symbol, halving at `limit`; rcb200.h)."""
        n = syms.numel()
        sb = self._sym_bytes(syms)
        p = AdaptiveParams(K, inc, limit)
        n_chunks = (n + chunk_syms - 1) // chunk_syms
        if out is None:
            cap = int(self.lib.rcb_adaptive_encode_bound(ctypes.byref(p), n, chunk_syms))
            if cap == 0:
                raise RcbError(_lib.RCB_ERR_UNSUPPORTED, "rcb_adaptive_encode_bound")
            out = torch.empty(cap + 16, dtype=torch.uint8, device=self.device)
        if offsets is None:
            offsets = torch.empty(n_chunks + 1, dtype=torch.int64, device=self.device)
        nbytes = ctypes.c_uint64()
        rc = self.lib.rcb_adaptive_encode_chunks(self.h, _ptr(syms), n, sb, chunk_syms, ctypes.byref(p), _ptr(out),
                                                 out.numel(), _ptr(offsets), _ptr(status), ctypes.byref(nbytes))
        self._check(rc, "rcb_adaptive_encode_chunks")
        return out, offsets, int(nbytes.value)

    def adaptive_decode_chunks(self, stream, offsets, n_syms, chunk_syms, K, inc=24, limit=60000, sym_bytes=1, out=None,
                               status=None):
        p = AdaptiveParams(K, inc, limit)
        if out is None:
            out = torch.empty(n_syms, dtype=torch.uint8 if sym_bytes == 1 else torch.int16, device=self.device)
        rc = self.lib.rcb_adaptive_decode_chunks(self.h, _ptr(stream), _ptr(offsets), n_syms, sym_bytes, chunk_syms,
                                                 ctypes.byref(p), _ptr(out), _ptr(status))
        self._check(rc, "rcb_adaptive_decode_chunks")
        return out

    # ------------------------------------------------- host-buffer entry points
    def encode_host(self, syms_np, chunk_syms, model, out_np=None, restart_syms=0, restart_np=None):
        """rcb_encode_host[_restart]: host symbols in, host code stream + offsets out.  With restart_syms the
        restart points land in restart_np (uint64[n_chunks * per_chunk * 3], allocated when None) and are
        returned as a fourth element."""
        syms_np = np.ascontiguousarray(syms_np)
        sb, n = syms_np.dtype.itemsize, syms_np.size
        n_chunks = (n + chunk_syms - 1) // chunk_syms
        if out_np is None:
            out_np = np.empty(self.encode_bound(model, n, sb, chunk_syms) + 16, dtype=np.uint8)
        offsets = np.zeros(n_chunks + 1, dtype=np.uint64)
        nbytes = ctypes.c_uint64()
        per = int(self.lib.rcb_restart_points_per_chunk(chunk_syms, restart_syms)) if restart_syms else 0
        if per:
            if restart_np is None:
                restart_np = np.zeros(max(1, n_chunks * per * 3), dtype=np.uint64)
            rc = self.lib.rcb_encode_host_restart(self.h, _np_ptr(syms_np), n, sb, chunk_syms, model.h,
                                                  _np_ptr(out_np), out_np.size, _np_ptr(offsets), ctypes.byref(nbytes),
                                                  restart_syms, _np_ptr(restart_np))
            self._check(rc, "rcb_encode_host_restart")
            return out_np, offsets, int(nbytes.value), restart_np
        rc = self.lib.rcb_encode_host(self.h, _np_ptr(syms_np), n, sb, chunk_syms, model.h, _np_ptr(out_np),
                                      out_np.size, _np_ptr(offsets), ctypes.byref(nbytes))
        self._check(rc, "rcb_encode_host")
        return out_np, offsets, int(nbytes.value)

    def decode_host(self, stream_np, offsets_np, n_syms, chunk_syms, model, sym_bytes=1, out_np=None, restart_syms=0,
                    restart_np=None):
        stream_np = np.ascontiguousarray(stream_np, dtype=np.uint8)
        offsets_np = np.ascontiguousarray(offsets_np, dtype=np.uint64)
        if out_np is None:
            out_np = np.empty(n_syms, dtype=np.uint8 if sym_bytes == 1 else np.uint16)
        if restart_syms and restart_np is not None:
            rc = self.lib.rcb_decode_host_restart(self.h, _np_ptr(stream_np), _np_ptr(offsets_np), n_syms, sym_bytes,
                                                  chunk_syms, model.h, _np_ptr(out_np), restart_syms,
                                                  _np_ptr(np.ascontiguousarray(restart_np, dtype=np.uint64)))
            self._check(rc, "rcb_decode_host_restart")
            return out_np
        rc = self.lib.rcb_decode_host(self.h, _np_ptr(stream_np), _np_ptr(offsets_np), n_syms, sym_bytes,
                                      chunk_syms, model.h, _np_ptr(out_np))
        self._check(rc, "rcb_decode_host")
        return out_np

    # ------------------------------------------------------- framed container
    def frame_encode(self, syms_np, chunk_syms, model, restart_syms=0):
        """encode_host + the RCB2 container (rcb200.h); returns the frame as uint8[].  restart_syms != 0 writes
        a version-2 frame that carries restart points (several decoder lanes per chunk)."""
        syms_np = np.ascontiguousarray(syms_np)
        sb, n = syms_np.dtype.itemsize, syms_np.size
        n_chunks = (n + chunk_syms - 1) // chunk_syms
        cap = self.lib.rcb_frame_bound_restart(model.K, n_chunks, int(model.n_models != 1),
                                               self.encode_bound(model, n, sb, chunk_syms), chunk_syms, restart_syms)
        frame = np.empty(cap, dtype=np.uint8)
        nbytes = ctypes.c_uint64()
        rc = self.lib.rcb_frame_encode_host_restart(self.h, _np_ptr(syms_np), n, sb, chunk_syms, model.h,
                                                    restart_syms, _np_ptr(frame), frame.size, ctypes.byref(nbytes))
        self._check(rc, "rcb_frame_encode_host")
        return frame[:int(nbytes.value)]

    def frame_info(self, frame_np):
        frame_np = np.ascontiguousarray(frame_np, dtype=np.uint8)
        info = FrameInfo()
        self._check(self.lib.rcb_frame_parse(_np_ptr(frame_np), frame_np.size, ctypes.byref(info)), "rcb_frame_parse")
        return info

    def frame_decode(self, frame_np):
        frame_np = np.ascontiguousarray(frame_np, dtype=np.uint8)
        info = self.frame_info(frame_np)
        out = np.empty(info.n_syms, dtype=np.uint8 if info.sym_bytes == 1 else np.uint16)
        n = ctypes.c_uint64()
        rc = self.lib.rcb_frame_decode_host(self.h, _np_ptr(frame_np), frame_np.size, _np_ptr(out), out.nbytes,
                                            ctypes.byref(n))
        self._check(rc, "rcb_frame_decode_host")
        return out

    # ------------------------------------------------------- synthetic data
    def generate(self, n, K, seed, thresholds, sym_bytes=1, chunk_syms=0, first=0, out=None):
        """Counter-based synthetic symbols; thresholds uint32[n_tables][K-1]."""
        thr = np.ascontiguousarray(thresholds, dtype=np.uint32)
        if thr.ndim == 1:
            thr = thr[None, :]
        assert thr.shape[1] == K - 1
        if out is None:
            dt = torch.uint8 if sym_bytes == 1 else torch.int16
            out = torch.empty(n, dtype=dt, device=self.device)
        rc = self.lib.rcb_generate(self.h, _ptr(out), first, n, sym_bytes, K, seed, _np_ptr(thr), thr.shape[0],
                                   chunk_syms)
        self._check(rc, "rcb_generate")
        return out


def zipf_thresholds(K, s):
    """floor(2^32 * CDF(i)) for i < K-1 of P(i) ~ (i+1)^-s, computed once on the
    host in double; shared verbatim by the CPU and GPU generators."""
    w = np.arange(1, K + 1, dtype=np.float64) ** (-float(s))
    cdf = np.cumsum(w) / np.sum(w)
    thr = np.floor(cdf[: K - 1] * 4294967296.0)
    return np.minimum(thr, 4294967295.0).astype(np.uint32)
