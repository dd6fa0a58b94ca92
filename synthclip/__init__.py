"""Synthetic clip generator: benchmark / test INPUT, not part of the product.

`schedule.py` builds the seeded, integer-only frame schedule (which scenes, blend, pan, noise stream per frame);
pixels come from `synth_core.h`, evaluated either on the GPU by `libesd_synth.so` (this package, `fill`) or on the CPU
by the oracle's twin (`oracle.c_oracle.synth_frames`, which includes the same header).  Nothing under `eioku_b200/`
imports this package and `libesd.so` exports no generator symbol; `bench.py --impl reference` uses only the CPU twin.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .schedule import (DESC_FIELDS, SCENE_BLACK, SCENE_WHITE, ClipSchedule, bgr_to_test_nv12, nv12_to_i420,  # noqa: F401
                       build_schedule)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libesd_synth.so")
_lib = None


def load_library():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not found: build it with `python -m synthclip.build` (nvcc, sm_100a)")
        L = C.CDLL(LIB_PATH)
        L.syn_fill.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_uint32, C.c_void_p, C.c_int64,
                               C.c_int, C.c_void_p]
        L.syn_fill.restype = C.c_int
        L.syn_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def fill(out_tensor, seed: int, descs: np.ndarray, chunk: int = 0):
    """Fill a CUDA uint8 tensor [N,H,W,3] with the clip frames described by `descs` (int32 [N,8])."""
    import torch

    L = load_library()
    descs = np.ascontiguousarray(descs, np.int32)
    n, h, w, _ = out_tensor.shape
    assert out_tensor.is_cuda and out_tensor.dtype == torch.uint8 and out_tensor.is_contiguous()
    assert descs.shape == (n, DESC_FIELDS)
    stream = torch.cuda.current_stream(out_tensor.device).cuda_stream
    step = chunk if chunk > 0 else n
    for a in range(0, n, step):
        part = out_tensor[a:a + step]
        rc = L.syn_fill(C.c_void_p(part.data_ptr()), w, h, out_tensor.stride(1), out_tensor.stride(0), seed & 0xFFFFFFFF,
                        descs[a:a + step].ctypes.data_as(C.c_void_p), part.shape[0], out_tensor.device.index, C.c_void_p(stream))
        if rc != 0:
            raise RuntimeError("syn_fill failed: " + (L.syn_last_error() or b"").decode())
    return out_tensor
