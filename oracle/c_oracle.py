"""CPU ORACLE (test infrastructure only) -- ctypes binding of oracle/esd_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libesd_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "esd_oracle.c")
    hdr = os.path.join(_HERE, "..", "synthclip", "synth_core.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, i32, i64, u32 = C.c_void_p, C.c_int, C.c_int64, C.c_uint32
        L.orc_axis_tables.argtypes = [i32, i32, vp, vp, vp, vp]
        L.orc_resize_linear.argtypes = [vp, i32, i32, i64, vp, i32, i32]
        L.orc_hsv_tables.argtypes = [vp, vp]
        L.orc_bgr2hsv.argtypes = [vp, i64, vp]
        L.orc_bgr2y.argtypes = [vp, i64, vp]
        L.orc_score_frames.argtypes = [vp, i64, i32, i32, i64, i64, i32, i32, vp, vp, vp, vp, i32]
        L.orc_synth_frames.argtypes = [u32, i32, i32, vp, i64, vp]
        for f in (L.orc_axis_tables, L.orc_resize_linear, L.orc_hsv_tables, L.orc_bgr2hsv, L.orc_bgr2y,
                  L.orc_score_frames, L.orc_synth_frames):
            f.restype = None
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def axis_tables(src: int, dst: int):
    o0 = np.empty(dst, np.int32); o1 = np.empty(dst, np.int32)
    c0 = np.empty(dst, np.int16); c1 = np.empty(dst, np.int16)
    lib().orc_axis_tables(src, dst, _p(o0), _p(o1), _p(c0), _p(c1))
    return o0, o1, c0, c1


def resize_linear(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    img = np.ascontiguousarray(img)
    out = np.empty((dh, dw, 3), np.uint8)
    lib().orc_resize_linear(_p(img), img.shape[0], img.shape[1], img.strides[0], _p(out), dh, dw)
    return out


def bgr2hsv(img: np.ndarray) -> np.ndarray:
    img = np.ascontiguousarray(img)
    out = np.empty_like(img)
    lib().orc_bgr2hsv(_p(img), img.size // 3, _p(out))
    return out


def bgr2y(img: np.ndarray) -> np.ndarray:
    img = np.ascontiguousarray(img)
    out = np.empty(img.shape[:-1], np.uint8)
    lib().orc_bgr2y(_p(img), img.size // 3, _p(out))
    return out


def score_frames(frames: np.ndarray, dw: int, dh: int, prev_hsv=None, bins: int = 0):
    """frames uint8 [N,H,W,3] -> (sums int64[N,3], hist uint32[N,bins] | None, last_hsv uint8[dh,dw,3])."""
    frames = np.ascontiguousarray(frames)
    n, h, w, _ = frames.shape
    sums = np.zeros((n, 3), np.int64)
    hist = np.zeros((n, bins), np.uint32) if bins else None
    last = np.empty((dh, dw, 3), np.uint8)
    if prev_hsv is not None:
        prev_hsv = np.ascontiguousarray(prev_hsv)
    lib().orc_score_frames(_p(frames), n, h, w, frames.strides[1], frames.strides[0], dh, dw,
                           _p(prev_hsv), _p(last), _p(sums), _p(hist), bins)
    return sums, hist, last


def synth_frames(seed: int, w: int, h: int, descs: np.ndarray) -> np.ndarray:
    """descs: int32 [N,8] (syn_frame_desc rows) -> uint8 [N,h,w,3]."""
    descs = np.ascontiguousarray(descs, np.int32)
    out = np.empty((descs.shape[0], h, w, 3), np.uint8)
    lib().orc_synth_frames(seed & 0xFFFFFFFF, w, h, _p(descs), descs.shape[0], _p(out))
    return out
