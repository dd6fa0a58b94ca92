"""CPU: libesd's CUDA-free host logic (eioku_b200/csrc/host_tables.h) -- the INTER_LINEAR and INTER_AREA coefficient tables
against the cv2-pinned oracle's, and the fused kernel's work plan against its covering properties."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import closed_form as cf


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = tmp_path_factory.mktemp("tables") / "tables_shim.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so),
                           os.path.join(ROOT, "tests", "tables_shim.cpp")])
    L = C.CDLL(str(so))
    vp = C.c_void_p
    L.shim_axis_tables.argtypes = [C.c_int, C.c_int, vp, vp, vp, vp]
    L.shim_area_table.argtypes = [C.c_int, C.c_int, vp, vp, vp, C.c_int]
    L.shim_unit_plan.argtypes = [C.c_int, C.c_longlong, C.c_longlong, C.c_int, vp, C.c_int, vp, C.c_int, vp]
    L.shim_tap_wavefronts.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
    L.shim_tap_wavefronts.restype = C.c_longlong
    return L


@pytest.mark.parametrize("src,dst", [(1920, 256), (1080, 144), (1280, 256), (720, 144), (3840, 256), (2160, 144), (1920, 274), (1080, 154),
                                     (854, 285), (480, 160), (300, 256), (200, 171), (257, 256), (7, 3), (2, 1), (4096, 1024)])
def test_linear_axis_tables_equal_the_oracle(shim, src, dst):
    o0, o1, c0, c1 = (np.zeros(dst, np.int32) for _ in range(4))
    shim.shim_axis_tables(src, dst, o0.ctypes.data, o1.ctypes.data, c0.ctypes.data, c1.ctypes.data)
    w0, w1, a0, a1 = cf.linear_axis_tables(src, dst)
    assert np.array_equal(o0, w0) and np.array_equal(o1, w1)
    assert np.array_equal(c0, a0.astype(np.int32)) and np.array_equal(c1, a1.astype(np.int32))
    assert np.all(np.abs(c0 + c1 - 2048) <= 1)   # what the kernel's unsaturated vertical pass relies on


@pytest.mark.parametrize("src,dst", [(256, 32), (144, 32), (274, 32), (154, 32), (285, 32), (171, 32), (100, 32), (77, 32), (64, 32),
                                     (32, 32), (33, 32), (300, 64), (1920, 32), (1080, 32), (2000, 32), (320, 6), (120, 20)])
def test_area_axis_tables_equal_the_oracle(shim, src, dst):
    cap = src + 2 * dst + 8
    begin, idx, wt = np.zeros(dst + 1, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.float32)
    n = shim.shim_area_table(src, dst, begin.ctypes.data, idx.ctypes.data, wt.ctypes.data, cap)
    want = cf.area_axis_table(src, dst)
    assert n == len(want) and begin[0] == 0 and begin[-1] == n
    for d in range(dst):
        ent = [(d, int(idx[k]), np.float32(wt[k])) for k in range(begin[d], begin[d + 1])]
        assert ent == [e for e in want if e[0] == d], d
        assert all(idx[k + 1] == idx[k] + 1 for k in range(begin[d], begin[d + 1] - 1))   # hash_kernel walks them as consecutive pixels


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("groups,n,ctas", [(9, 2048, 296), (9, 1, 296), (9, 7, 296), (36, 64, 296), (1, 5, 296), (18, 18000, 296),
                                           (9, 808, 296), (3, 2, 4), (135, 33, 148)])
def test_unit_plan_covers_every_item_once(shim, mode, groups, n, ctas):
    cap = groups * n + 4 * ctas + 16
    units, begin, nu = np.zeros((cap, 4), np.int32), np.zeros(ctas + 2, np.int32), C.c_int()
    grid = shim.shim_unit_plan(groups, n, ctas, mode, units.ctypes.data, cap, begin.ctypes.data, len(begin), C.byref(nu))
    assert 1 <= grid <= min(ctas, groups * n)
    units = units[:nu.value]
    assert begin[0] == 0 and begin[grid] == nu.value and np.all(np.diff(begin[:grid + 1]) >= 0)
    cover = np.zeros((groups, n), np.int32)
    for rg, f0, f1, _ in units:
        assert 0 <= rg < groups and 0 <= f0 < f1 <= n
        cover[rg, f0:f1] += 1
    assert np.all(cover == 1)
    per_cta = [sum(int(u[2] - u[1]) for u in units[begin[g]:begin[g + 1]]) for g in range(grid)]
    if mode == 1:   # strips: equal shares, at most one halo re-read (a unit not starting at frame 0) more than row-group changes
        assert max(per_cta) - min(per_cta) <= 1


@pytest.mark.parametrize("src_w,dst_w,bytes_per_px,n_words,want_conflict_free", [
    (1920, 256, 3, 3, True),    # 1080p BGR24: 8 columns = 60 px = 180 B = 45 words (odd) apart
    (1280, 256, 3, 3, True),    # 720p: 4 columns = 20 px = 60 B = 15 words
    (3840, 256, 3, 3, True),    # 4K: 4 columns = 60 px = 180 B
    (1920, 256, 1, 2, True),    # NV12 luma: 8 columns = 60 B = 15 words
    (256, 256, 6, 3, True),     # gathered taps: 6 B per column, 2 columns = 3 words
    (854, 285, 3, 3, False),    # irregular lattice: whatever is best, never worse than stride 1
])
def test_lane_stride_choice_removes_bank_conflicts(shim, src_w, dst_w, bytes_per_px, n_words, want_conflict_free):
    """The consumer lanes of the fused kernel read their taps from the staged row; with neighbouring columns on neighbouring
    lanes 1080p costs two passes per load.  The host picks the lane stride whose loads touch each bank once."""
    if bytes_per_px == 6:
        off = (6 * np.arange(dst_w)).astype(np.uint32)
    else:
        o0, _, _, _ = cf.linear_axis_tables(src_w, dst_w)
        off = (bytes_per_px * o0).astype(np.uint32)
    best = C.c_int()
    cost = {ks: shim.shim_tap_wavefronts(off.ctypes.data, dst_w, n_words, ks, C.byref(best)) for ks in (1, 2, 4, 8)}
    ideal = sum(n_words for w in range(8) if 32 * w < dst_w or True)  # one pass per load and warp
    assert cost[best.value] == min(cost.values())
    if want_conflict_free:
        assert cost[best.value] == 8 * n_words == ideal, cost
        assert cost[1] > cost[best.value], cost   # the round-1 mapping did conflict
    # the mapping is a permutation of the columns
    for ks in (1, 2, 4, 8):
        cols = sorted(ks * l + (w & (ks - 1)) + 32 * ks * (w // ks) for w in range(8) for l in range(32))
        assert cols == list(range(256))
