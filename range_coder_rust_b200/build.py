"""In-tree build of librcb200.so (sm_100a only) with nvcc.

The shared library is built next to this file so that it travels with the
repository snapshot to the GPU box; nothing is installed into site-packages.
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librcb200.so")
SOURCES = ["rcb_api.cu"]
HEADERS = ["rcb_core.cuh", "rcb_kernels.cuh", "rcb_encode.cuh", "rcb_decode.cuh", "rcb_decode_row.cuh", "rcb_stream.cuh", "rcb_comm.cuh", "rcb_adaptive.cuh", os.path.join("..", "..", "include", "rcb200.h")]

NVCC_FLAGS = [
    "-std=c++17",
    "-O3",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-shared",
    "-cudart", "shared",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: librcb200.so cannot be built (there is no CPU fallback)")


STAMP = LIB + ".srchash"  # travels with the .so; file mtimes do not survive a snapshot copy


def source_hash():
    h = hashlib.sha256()
    for dep in [os.path.join(CSRC, s) for s in SOURCES + HEADERS]:
        h.update(open(dep, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def needs_build():
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    return open(STAMP).read().strip() != source_hash()


def build(force=False, verbose=False, extra=()):
    if not force and not needs_build():
        return LIB
    cmd = [find_nvcc()] + NVCC_FLAGS + list(extra)
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building librcb200.so")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    with open(STAMP, "w") as f:
        f.write(source_hash() + "\n")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
