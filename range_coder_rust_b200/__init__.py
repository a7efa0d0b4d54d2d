"""range_coder_rust_b200 -- B200-native chunk-parallel range coder.

Bit-exact, per chunk, with the reference crate diegodox/range_coder_rust
(`range_coder` v0.1.0); the hot path lives in hand-written sm_100a CUDA behind
the C ABI in include/rcb200.h.  This package only binds that ABI.
"""
from ._lib import LIB_PATH, RcbError, load  # noqa: F401

__all__ = ["Context", "Model", "Comm", "RcbError", "zipf_thresholds", "load", "LIB_PATH"]


def __getattr__(name):
    # torch is imported lazily so that `import range_coder_rust_b200` (and the
    # ABI symbol checks) work without initialising CUDA.
    if name in ("Context", "Model", "Comm", "zipf_thresholds"):
        from . import api

        return getattr(api, name)
    raise AttributeError(name)
