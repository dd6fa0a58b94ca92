"""Seeded synthetic clip *schedule* (SURVEY.md section 8d): which scene(s), blend, pan
and noise stream every frame is made of.  The schedule is pure Python/integer, so it
is identical everywhere; pixels are produced from it either on the GPU
(``synthclip.fill`` -> ``libesd_synth.so``) or by the CPU twin that the
test-suite's oracle builds from the same ``synthclip/synth_core.h``.

Clip structure: scene lengths uniform in [45, 240] frames; every 5th transition is a
24-frame dissolve, every 11th a 30-frame fade through black, the rest hard cuts;
every 17th scene holds a 3-frame white flash (exercises the flash / min-scene-len
filter); a slow horizontal pan and +-2 per-byte noise inside scenes.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

SCENE_BLACK = -1
SCENE_WHITE = -2
DESC_FIELDS = 8  # int32 columns of syn_frame_desc (csrc/synth_core.h)

_M32 = 0xFFFFFFFF


def _hash(x: int) -> int:
    """syn_hash of csrc/synth_core.h on Python ints."""
    x &= _M32
    x ^= x >> 16
    x = (x * 0x7FEB352D) & _M32
    x ^= x >> 15
    x = (x * 0x846CA68B) & _M32
    x ^= x >> 16
    return x


def _hash3(seed: int, a: int, b: int) -> int:
    h = _hash(seed ^ 0x9E3779B9)
    h = _hash((h + a * 0x85EBCA6B + 0x165667B1) & _M32)
    h = _hash(h ^ ((b * 0xC2B2AE35 + 0x27D4EB2F) & _M32))
    return h


@dataclass
class ClipSchedule:
    seed: int
    n_frames: int
    descs: np.ndarray  # int32 [n_frames, 8]
    hard_cuts: List[int]  # first frame of the new scene, hard cuts only
    dissolves: List[Tuple[int, int]]  # [start, end) frame ranges
    fades: List[Tuple[int, int]]
    flashes: List[Tuple[int, int]]


def build_schedule(seed: int, n_frames: int, *, noise_amp: int = 2, pan_div: int = 4,
                   min_len: int = 45, max_len: int = 240) -> ClipSchedule:
    descs = np.zeros((n_frames, DESC_FIELDS), np.int32)
    hard_cuts: List[int] = []
    dissolves: List[Tuple[int, int]] = []
    fades: List[Tuple[int, int]] = []
    flashes: List[Tuple[int, int]] = []

    def put(t, a, b, num, den, pa, pb):
        if 0 <= t < n_frames:
            descs[t] = (a, b, num, den, pa, pb, t, noise_amp)

    t = 0
    k = 0
    while t < n_frames:
        length = min_len + _hash3(seed, 0x5CE7E, k) % (max_len - min_len + 1)
        start = t
        pan = lambda u, s=start: (u - s) // pan_div if pan_div > 0 else 0  # noqa: E731
        for u in range(start, min(start + length, n_frames)):
            put(u, k, k, 0, 1, pan(u), 0)
        if k % 17 == 16:  # 3-frame white flash in the middle of the scene
            f0 = start + length // 2
            for u in range(f0, f0 + 3):
                put(u, SCENE_WHITE, SCENE_WHITE, 0, 1, 0, 0)
            if f0 < n_frames:
                flashes.append((f0, min(f0 + 3, n_frames)))
        t = start + length
        kind = "fade" if k % 11 == 10 else ("dissolve" if k % 5 == 4 else "cut")
        if kind == "cut":
            if t < n_frames:
                hard_cuts.append(t)
        elif kind == "dissolve":
            n = 24
            for i in range(n):  # frames t..t+n-1 blend k -> k+1
                u = t + i
                put(u, k, k + 1, i + 1, n + 1, pan(u), 0)
            if t < n_frames:
                dissolves.append((t, min(t + n, n_frames)))
            t += n
        else:
            n = 30
            half = n // 2
            for i in range(half):
                put(t + i, k, SCENE_BLACK, i + 1, half, pan(t + i), 0)
            for i in range(half):
                put(t + half + i, SCENE_BLACK, k + 1, i + 1, half + 1, 0, 0)
            if t < n_frames:
                fades.append((t, min(t + n, n_frames)))
            t += n
        k += 1
    return ClipSchedule(seed, n_frames, descs, hard_cuts, dissolves, fades, flashes)


def bgr_to_test_nv12(frames):
    """Deterministic NV12 test content from BGR frames [N,H,W,3] (numpy or torch; pure indexing, so identical on both):
    Y = the G channel, U / V = the B / R channels at the even rows and columns.  Returns [N, H * 3 // 2, W] uint8 --
    the contiguous NV12 layout (Y plane, then the interleaved UV plane).  Not a colour-space conversion: it only makes
    NV12 frames whose scene structure (cuts, fades, flashes) follows the synthetic clip."""
    n, h, w, _ = frames.shape
    assert h % 2 == 0 and w % 2 == 0
    if isinstance(frames, np.ndarray):
        out = np.empty((n, h * 3 // 2, w), np.uint8)
    else:
        import torch

        out = torch.empty((n, h * 3 // 2, w), dtype=torch.uint8, device=frames.device)
    out[:, :h, :] = frames[..., 1]
    out[:, h:, 0::2] = frames[:, 0::2, 0::2, 0]
    out[:, h:, 1::2] = frames[:, 0::2, 0::2, 2]
    return out


def nv12_to_i420(nv12):
    """The same samples as planar I420: [N, H * 3 // 2, W] uint8 with the Y plane, then the U plane and the V plane (each
    H/2 rows of W/2 bytes, stored densely) -- the array cv2.cvtColor(COLOR_YUV2BGR_I420) takes.  numpy or torch."""
    n, rows, w = nv12.shape
    h = rows * 2 // 3
    out = nv12.copy() if isinstance(nv12, np.ndarray) else nv12.clone()
    q = (h // 2) * (w // 2)
    flat = out.reshape(n, rows * w)
    flat[:, h * w:h * w + q] = nv12[:, h:, 0::2].reshape(n, q)
    flat[:, h * w + q:] = nv12[:, h:, 1::2].reshape(n, q)
    return out
