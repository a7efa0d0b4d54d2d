"""ctypes binding of librcb200.so (the C ABI in include/rcb200.h).

The library is the product: there is no Python or CPU implementation behind
it.  Loading fails loudly if the shared object has not been built.
"""
import ctypes
import os

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librcb200.so")

u8p = ctypes.POINTER(ctypes.c_uint8)
u32p = ctypes.POINTER(ctypes.c_uint32)
u64p = ctypes.POINTER(ctypes.c_uint64)
vp = ctypes.c_void_p
u64 = ctypes.c_uint64
u32 = ctypes.c_uint32
ci = ctypes.c_int

# name -> (restype, argtypes); mirrors include/rcb200.h one to one
SIGNATURES = {
    "rcb_strerror": (ctypes.c_char_p, [ci]),
    "rcb_version": (ci, []),
    "rcb_last_cuda_error": (ci, [vp, ctypes.POINTER(ctypes.c_char_p)]),
    "rcb_device_count": (ci, []),
    "rcb_ctx_create": (ci, [ci, vp, ctypes.POINTER(vp)]),
    "rcb_ctx_destroy": (ci, [vp]),
    "rcb_ctx_set_stream": (ci, [vp, vp]),
    "rcb_ctx_synchronize": (ci, [vp]),
    "rcb_ctx_set_block_threads": (ci, [vp, ci, ci]),
    "rcb_ctx_launch_count": (u64, [vp]),
    "rcb_ctx_enable_timing": (ci, [vp, ci]),
    "rcb_ctx_get_timings": (ci, [vp, ctypes.POINTER(ctypes.c_float), ci]),
    "rcb_device_alloc": (ci, [vp, u64, ctypes.POINTER(vp)]),
    "rcb_device_free": (ci, [vp, vp]),
    "rcb_copy_to_device": (ci, [vp, vp, vp, u64]),
    "rcb_copy_to_host": (ci, [vp, vp, vp, u64]),
    "rcb_model_create": (ci, [vp, u32, u64, ctypes.POINTER(vp)]),
    "rcb_model_destroy": (ci, [vp]),
    "rcb_histogram": (ci, [vp, vp, u64, ci, u32, u64, vp]),
    "rcb_model_from_counts": (ci, [vp, vp, vp, ci]),
    "rcb_model_from_tables": (ci, [vp, vp, vp, vp, vp]),
    "rcb_model_get_tables": (ci, [vp, vp, u64, vp, vp, vp, vp]),
    "rcb_encode_chunks": (ci, [vp, vp, u64, ci, u64, vp, vp, u64, vp, vp, u64p]),
    "rcb_encode_chunks_async": (ci, [vp, vp, u64, ci, u64, vp, vp, u64, vp, vp]),
    "rcb_encode_result": (ci, [vp, u64p]),
    "rcb_encode_bound": (u64, [vp, vp, u64, ci, u64]),
    "rcb_decode_chunks": (ci, [vp, vp, vp, u64, ci, u64, vp, vp, vp]),
    "rcb_decode_chunks_async": (ci, [vp, vp, vp, u64, ci, u64, vp, vp, vp]),
    "rcb_decode_result": (ci, [vp]),
    "rcb_restart_points_per_chunk": (u64, [u64, u64]),
    "rcb_encode_chunks_restart": (ci, [vp, vp, u64, ci, u64, vp, vp, u64, vp, vp, u64, vp, u64p]),
    "rcb_encode_chunks_restart_async": (ci, [vp, vp, u64, ci, u64, vp, vp, u64, vp, vp, u64, vp]),
    "rcb_decode_chunks_restart": (ci, [vp, vp, vp, u64, ci, u64, vp, vp, vp, u64, vp]),
    "rcb_decode_chunks_restart_async": (ci, [vp, vp, vp, u64, ci, u64, vp, vp, vp, u64, vp]),
    "rcb_encode_host": (ci, [vp, vp, u64, ci, u64, vp, vp, u64, vp, u64p]),
    "rcb_decode_host": (ci, [vp, vp, vp, u64, ci, u64, vp, vp]),
    "rcb_encode_host_restart": (ci, [vp, vp, u64, ci, u64, vp, vp, u64, vp, u64p, u64, vp]),
    "rcb_decode_host_restart": (ci, [vp, vp, vp, u64, ci, u64, vp, vp, u64, vp]),
    "rcb_stream_state_init": (None, [vp]),
    "rcb_encode_stream": (ci, [vp, vp, vp, u64, ci, vp, vp, u64, u64p, vp, ci]),
    "rcb_decode_stream": (ci, [vp, vp, vp, u64, u64, ci, vp, vp]),
    "rcb_frame_bound": (u64, [u32, u64, ci, u64]),
    "rcb_frame_write": (ci, [vp, vp, ci, u64, u64, vp, vp, vp, u64, u64p]),
    "rcb_frame_parse": (ci, [vp, u64, vp]),
    "rcb_frame_model": (ci, [vp, vp, vp, ctypes.POINTER(vp)]),
    "rcb_frame_encode_host": (ci, [vp, vp, u64, ci, u64, vp, vp, u64, u64p]),
    "rcb_frame_decode_host": (ci, [vp, vp, u64, vp, u64, u64p]),
    "rcb_frame_bound_restart": (u64, [u32, u64, ci, u64, u64, u64]),
    "rcb_frame_write_restart": (ci, [vp, vp, ci, u64, u64, vp, vp, u64, vp, vp, u64, u64p]),
    "rcb_frame_encode_host_restart": (ci, [vp, vp, u64, ci, u64, vp, u64, vp, u64, u64p]),
    "rcb_generate": (ci, [vp, vp, u64, u64, ci, u32, u64, vp, u32, u64]),
    "rcb_adaptive_encode_bound": (u64, [vp, u64, u64]),
    "rcb_adaptive_encode_chunks": (ci, [vp, vp, u64, ci, u64, vp, vp, u64, vp, vp, u64p]),
    "rcb_adaptive_decode_chunks": (ci, [vp, vp, vp, u64, ci, u64, vp, vp, vp]),
    "rcb_comm_unique_id": (ci, [vp]),
    "rcb_comm_init_rank": (ci, [vp, vp, ci, ci, ctypes.POINTER(vp)]),
    "rcb_comm_init_all": (ci, [ctypes.POINTER(vp), ci, ctypes.POINTER(vp)]),
    "rcb_comm_destroy": (ci, [vp]),
    "rcb_comm_info": (ci, [vp, ctypes.POINTER(ci), ctypes.POINTER(ci), ctypes.POINTER(ci)]),
    "rcb_comm_last_error": (ctypes.c_char_p, [vp]),
    "rcb_allreduce_counts": (ci, [vp, vp, vp, u32]),
    "rcb_allreduce_counts_multi": (ci, [ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp), u32, ci]),
}

_lib = None


def load(build_if_missing=True):
    """Load librcb200.so, building it in-tree with nvcc first if needed."""
    global _lib
    if _lib is not None:
        return _lib
    if _build.needs_build():
        # stale or missing: rebuild in-tree (content hash, not mtimes, decides)
        if not build_if_missing:
            raise RuntimeError(
                "librcb200.so is missing or stale (run `python -m range_coder_rust_b200.build`); "
                "range_coder_rust_b200 has no CPU fallback"
            )
        _build.build(force=True)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and the header diverge
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class RcbError(RuntimeError):
    """A non-zero rcb_error returned by the C ABI."""

    def __init__(self, code, where="", detail=""):
        self.code = code
        msg = load().rcb_strerror(code).decode()
        super().__init__(f"{where}: {msg} ({code}){(' - ' + detail) if detail else ''}")


# rcb_error values (include/rcb200.h)
RCB_OK = 0
RCB_ERR_INVALID_ARGUMENT = -1
RCB_ERR_CUDA = -2
RCB_ERR_ZERO_TOTAL = -3
RCB_ERR_ZERO_FREQ_SYMBOL = -4
RCB_ERR_LOWER_OVERFLOW = -5
RCB_ERR_UPPER_OVERFLOW = -6
RCB_ERR_SYMBOL_OUT_OF_RANGE = -7
RCB_ERR_OUT_CAPACITY = -8
RCB_ERR_TRUNCATED_STREAM = -9
RCB_ERR_INVALID_MODEL = -10
RCB_ERR_UNSUPPORTED = -11
RCB_ERR_NO_DEVICE = -12
RCB_ERR_NCCL = -13

RCB_MODEL_POW2 = 1
RCB_MODEL_CONSISTENT = 2
RCB_MODEL_REGULAR = 4
