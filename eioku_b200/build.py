"""Build libesd.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libesd.so")
SOURCES = ["esd.cu"]
DEPS = ["esd.cu", "esd_kernels.cuh", "ingest_gather.h", "host_tables.h", "guard_alloc.h", os.path.join("..", "..", "include", "esd.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-shared",
    "-fmad=false",  # float64 score math must not be contracted (SURVEY.md hard parts)
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden",
    "-Xptxas", "-v",
    "-ldl",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libesd.so cannot be built (there is no CPU fallback)")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


DECODE_LIB = os.path.join(HERE, "libesd_decode.so")
DECODE_DEPS = ["esd_decode.cu", "jpeg_core.h", "jpeg_parse.h", "guard_alloc.h", os.path.join("..", "..", "include", "esd_decode.h")]


def build_decode(force: bool = False, verbose: bool = False) -> str:
    """libesd_decode.so: Motion-JPEG (AVI) -> device BGR frames through nvJPEG (include/esd_decode.h)."""
    stale = not os.path.exists(DECODE_LIB) or any(os.path.getmtime(os.path.join(CSRC, d)) > os.path.getmtime(DECODE_LIB) for d in DECODE_DEPS)
    if not force and not stale:
        return DECODE_LIB
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
           "-Xcompiler", "-fPIC,-fvisibility=hidden", "-o", DECODE_LIB, os.path.join(CSRC, "esd_decode.cu"), "-lnvjpeg"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libesd_decode.so")
    return DECODE_LIB


def build(force: bool = False, verbose: bool = False) -> str:
    build_decode(force, verbose)
    if not force and not is_stale():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libesd.so")
    with open(os.path.join(HERE, "libesd.ptxas.log"), "w") as f:
        f.write(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
