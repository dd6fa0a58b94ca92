"""Build libesd_synth.so (the CUDA filler of the synthetic clips) in-tree with nvcc for sm_100a."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libesd_synth.so")
DEPS = ["synth_fill.cu", "synth_core.h"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libesd_synth.so cannot be built")


def build(force: bool = False) -> str:
    stale = not os.path.exists(LIB) or any(os.path.getmtime(os.path.join(HERE, d)) > os.path.getmtime(LIB) for d in DEPS)
    if force or stale:
        cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
               "-Xcompiler", "-fPIC,-fvisibility=hidden", "-o", LIB, os.path.join(HERE, "synth_fill.cu")]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("nvcc failed building libesd_synth.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
