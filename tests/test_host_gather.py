"""CPU: the ingest path's host tap gather (eioku_b200/csrc/ingest_gather.h, CUDA-free) against a numpy statement of the
gathered layouts -- BGR24 and NV12, with and without non-temporal stores / software prefetch, clamped last columns, ragged
item ranges.  The GPU tests check that the fused kernel consumes these layouts bit-exactly."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import closed_form as cf


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = tmp_path_factory.mktemp("gather") / "gather_shim.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(so), os.path.join(ROOT, "tests", "gather_shim.cpp")])
    L = C.CDLL(str(so))
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    L.shim_gather.argtypes = [i32, i32, i32, vp, vp, i64, i32, i32, i32, i32, vp, i64, i64, vp, i64, i64, i32, i32]
    L.shim_gather.restype = None
    return L


def geometry(w, h, dw, dh):
    xo, _, _, _ = cf.linear_axis_tables(w, dw)
    yo, yo1, _, _ = cf.linear_axis_tables(h, dh)
    return xo.astype(np.int32), np.unique(np.concatenate([yo, yo1])).astype(np.int32)


def aligned(n, fill=0xEE):
    raw = np.full(n + 64, fill, np.uint8)
    o = (-raw.ctypes.data) % 64
    return raw[o:o + n]


@pytest.mark.parametrize("w,h,dw,dh", [(1920, 1080, 256, 144), (642, 362, 256, 144), (300, 200, 256, 171), (34, 18, 17, 9)])
@pytest.mark.parametrize("nt,pf,streams", [(0, 0, 1), (1, 4096, 1), (1, 0, 1), (0, 512, 1), (1, 4096, 8), (1, 0, 8)])
def test_bgr_tap_gather_layout(shim, w, h, dw, dh, nt, pf, streams):
    """streams = 8: the ingest path's default, eight rows gathered in lock-step (blocks of 16 items; ragged remainders fall back)."""
    rng = np.random.default_rng(w + dw)
    n = 3
    pitch = w * 3 + 5
    frames = rng.integers(0, 256, (n, h, pitch), dtype=np.uint8)
    xo, rows = geometry(w, h, dw, dh)
    off = (3 * xo).astype(np.int32)
    trb = (6 * dw + 15) & ~15
    nt_rows = len(rows)
    dst = aligned(n * nt_rows * trb)
    items = n * nt_rows
    for lo, hi in ((0, 7), (7, items - 3), (items - 3, items)):   # ragged ranges, as the worker pool hands them out
        shim.shim_gather(dw, trb, w * 3, off.ctypes.data, rows.ctypes.data, nt_rows, nt_rows, 0, min(pf, w * 3), nt,
                         frames.ctypes.data, frames.strides[0], pitch, dst.ctypes.data, lo, hi, 0, streams)
    got = dst.reshape(n, nt_rows, trb)
    px = frames[:, :, :w * 3].reshape(n, h, w, 3)
    x1 = np.minimum(xo + 1, w - 1)
    for d in range(dw):
        assert np.array_equal(got[:, :, 6 * d:6 * d + 3], px[:, rows, xo[d]]), d
        if xo[d] + 1 <= w - 1:   # a clamped last column carries weight 0 on tap 1: its bytes are unspecified
            assert np.array_equal(got[:, :, 6 * d + 3:6 * d + 6], px[:, rows, x1[d]]), d


@pytest.mark.parametrize("w,h,dw,dh", [(1920, 1080, 256, 144), (642, 362, 256, 144), (300, 200, 256, 171), (34, 18, 17, 9), (2050, 40, 1024, 20)])
@pytest.mark.parametrize("nt,pf", [(0, 0), (1, 4096), (0, 512)])
def test_nv12_tap_gather_layout(shim, w, h, dw, dh, nt, pf):
    rng = np.random.default_rng(w * 3 + dh)
    n = 2
    pitch = w + 6
    frames = rng.integers(0, 256, (n, h * 3 // 2, pitch), dtype=np.uint8)
    xo, yrows = geometry(w, h, dw, dh)
    uvrows = np.unique(yrows >> 1)
    touched = np.concatenate([yrows, h + uvrows]).astype(np.int32)
    trb = (4 * dw + 15) & ~15
    nt_rows = len(touched)
    dst = aligned(n * nt_rows * trb)
    items = n * nt_rows
    for lo, hi in ((0, 5), (5, items)):
        shim.shim_gather(dw, trb, w, xo.ctypes.data, touched.ctypes.data, nt_rows, len(yrows), 1, min(pf, w), nt,
                         frames.ctypes.data, frames.strides[0], pitch, dst.ctypes.data, lo, hi, 0, 1)
    got = dst.reshape(n, nt_rows, trb)
    x1 = np.minimum(xo + 1, w - 1)
    ysrc = frames[:, yrows, :w]
    uvsrc = frames[:, h + uvrows, :w]
    ny = len(yrows)
    for d in range(dw):
        assert np.array_equal(got[:, :ny, 2 * d], ysrc[:, :, xo[d]]) and np.array_equal(got[:, :ny, 2 * d + 1], ysrc[:, :, x1[d]]), d
        c0, c1 = xo[d] & ~1, x1[d] & ~1
        assert np.array_equal(got[:, ny:, 4 * d:4 * d + 2], uvsrc[:, :, c0:c0 + 2]), d
        assert np.array_equal(got[:, ny:, 4 * d + 2:4 * d + 4], uvsrc[:, :, c1:c1 + 2]), d
    if nt:
        assert not got[:, :ny, 2 * dw:].any()   # luma rows are zero-padded to the common pitch


@pytest.mark.parametrize("w,h,dw,dh", [(1920, 1080, 256, 144), (642, 362, 256, 144), (300, 200, 256, 171), (34, 18, 17, 9)])
@pytest.mark.parametrize("nt,pf", [(0, 0), (1, 4096)])
def test_i420_tap_gather_equals_nv12_gather_of_the_interleaved_frame(shim, w, h, dw, dh, nt, pf):
    """Planar I420 host frames gather into exactly the NV12 tap layout: same bytes as gathering the interleaved (NV12) form."""
    import synthclip as synth

    rng = np.random.default_rng(w + dh)
    n = 2
    nv12 = rng.integers(0, 256, (n, h * 3 // 2, w), dtype=np.uint8)
    i420 = synth.nv12_to_i420(nv12)
    xo, yrows = geometry(w, h, dw, dh)
    uvrows = np.unique(yrows >> 1)
    touched = np.concatenate([yrows, h + uvrows]).astype(np.int32)
    trb = (4 * dw + 15) & ~15
    nt_rows = len(touched)
    items = n * nt_rows
    out = []
    for fmt, frames in ((1, nv12), (2, i420)):
        dst = aligned(items * trb)
        for lo, hi in ((0, 3), (3, items)):
            shim.shim_gather(dw, trb, w, xo.ctypes.data, touched.ctypes.data, nt_rows, len(yrows), fmt, min(pf, w), nt,
                             frames.ctypes.data, frames.strides[0], w, dst.ctypes.data, lo, hi, h, 1)
        out.append(dst.reshape(n, nt_rows, trb)[:, :, :4 * dw].copy())
    ny = len(yrows)
    assert np.array_equal(out[0][:, :ny, :2 * dw], out[1][:, :ny, :2 * dw])
    assert np.array_equal(out[0][:, ny:], out[1][:, ny:])


def test_gather_loops_stay_inside_exact_size_buffers_under_asan(tmp_path):
    """tests/gather_asan.cpp: every gather variant (BGR24 one row / eight rows in lock-step, NV12, I420; plain and non-temporal
    stores) over dense frames and ring slots of exactly the advertised sizes, under AddressSanitizer + UBSan -- a read past the
    last row of a caller's frame would be a host crash in production and is invisible to the layout tests above."""
    import subprocess

    exe = str(tmp_path / "gather_asan")
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gather_asan.cpp")
    build = subprocess.run(["g++", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-o", exe, src],
                           capture_output=True, text=True)
    if build.returncode != 0 and "sanitize" in build.stderr:
        pytest.skip("this toolchain has no AddressSanitizer runtime")
    assert build.returncode == 0, build.stderr[-2000:]
    run = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0 and "gather asan ok: 60 runs" in run.stdout, (run.stdout[-300:], run.stderr[-3000:])
