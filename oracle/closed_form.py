"""CPU ORACLE (test infrastructure only) -- closed-form integer restatement.

This file is part of the *oracle*: a CPU restatement of the arithmetic that
PySceneDetect's detectors run through OpenCV.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product (``eioku_b200``) never does.

PARITY STATUS: "parity unpinned" against PySceneDetect itself -- the reference
repository (codihuston/eioku) neither vendors nor pins ``scenedetect`` and has
no golden vectors for this path (the only mention
of the producer is ``ml-service/src/models/responses.py:141-142``).  What *is*
pinned: every integer stage below is checked bit-exact against the real
``cv2`` 4.13.0 installed in this image (tests/test_oracle_vs_cv2.py) and
against committed golden vectors generated from it (tests/golden/).

Stages:
  A.1  compute_downscale_factor / target size  (scenedetect scene_manager.py)
  A.2  cv2.resize INTER_LINEAR, uint8, 11-bit fixed point (OpenCV resize.cpp)
  A.3  cv2.cvtColor BGR2HSV uint8, H in [0,180)        (OpenCV color_hsv)
  A.7  BGR -> Y of cv2.cvtColor BGR2YUV                 (OpenCV color_yuv)
  A.4  sum |cur - prev| per HSV plane                   (content_detector.py)
"""
from __future__ import annotations

import numpy as np

INTER_BITS = 11  # OpenCV INTER_RESIZE_COEF_BITS
INTER_SCALE = 1 << INTER_BITS
HSV_SHIFT = 12


# --------------------------------------------------------------------------- A.1
def compute_downscale_factor(frame_width: int, effective_width: int = 256, mode: str = "float"):
    """scenedetect.scene_manager.compute_downscale_factor.

    mode="float": PySceneDetect >= 0.6.2 (W / 256.0); mode="int": <= 0.6.1 (W // 256).
    """
    assert frame_width >= 1 and effective_width >= 1
    if frame_width < effective_width:
        return 1
    if mode == "int":
        return frame_width // effective_width
    return frame_width / float(effective_width)


def downscaled_size(width: int, height: int, factor) -> tuple[int, int]:
    """(dst_w, dst_h) as SceneManager passes to cv2.resize (Python round = half-even)."""
    if not factor > 1:
        return width, height
    return max(1, round(width / factor)), max(1, round(height / factor))


# --------------------------------------------------------------------------- A.2
def linear_axis_tables(src: int, dst: int):
    """Per-axis tap offsets and int16 coefficient pairs of cv2.resize(INTER_LINEAR).

    Follows OpenCV resize.cpp: scale is a double, the fractional position is
    narrowed to float32, coefficients are float32 products rounded half-even
    (cvRound) to the 11-bit fixed-point scale.
    Returns (ofs int32[dst], ofs1 int32[dst], c0 int16[dst], c1 int16[dst]).
    """
    scale = 1.0 / (dst / src)  # inv_scale = dst/src (double); scale = 1/inv_scale
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    lo = s < 0
    s[lo] = 0
    f[lo] = 0.0
    hi = s >= src - 1
    s[hi] = src - 1
    f[hi] = 0.0
    one = np.float32(1.0)
    sc = np.float32(INTER_SCALE)
    c0 = np.rint((one - f) * sc).astype(np.int16)
    c1 = np.rint(f * sc).astype(np.int16)
    s1 = np.minimum(s + 1, src - 1).astype(np.int32)
    return s, s1, c0, c1


def resize_linear_u8(img: np.ndarray, dst_w: int, dst_h: int) -> np.ndarray:
    """cv2.resize(img, (dst_w, dst_h), interpolation=INTER_LINEAR) for uint8 HxWxC."""
    h, w = img.shape[:2]
    if (dst_w, dst_h) == (w, h):
        return img.copy()
    xo, xo1, a0, a1 = linear_axis_tables(w, dst_w)
    yo, yo1, b0, b1 = linear_axis_tables(h, dst_h)
    src = img.astype(np.int32)
    # horizontal pass on the two source rows of every destination row (scale 2^11)
    r0 = src[yo][:, xo] * a0.astype(np.int32)[None, :, None] + src[yo][:, xo1] * a1.astype(np.int32)[None, :, None]
    r1 = src[yo1][:, xo] * a0.astype(np.int32)[None, :, None] + src[yo1][:, xo1] * a1.astype(np.int32)[None, :, None]
    B0 = b0.astype(np.int32)[:, None, None]
    B1 = b1.astype(np.int32)[:, None, None]
    # VResizeLinear<uchar,int,short,FixedPtCast<int,uchar,22>>
    out = (((B0 * (r0 >> 4)) >> 16) + ((B1 * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def touched_rows(src_h: int, dst_h: int) -> np.ndarray:
    """Sorted unique source rows the vertical 2-tap filter reads."""
    yo, yo1, _, _ = linear_axis_tables(src_h, dst_h)
    return np.unique(np.concatenate([yo, yo1]))


# --------------------------------------------------------------------------- A.3
def hsv_tables():
    """sdiv_table / hdiv_table180 of OpenCV's RGB2HSV_b (hsv_shift = 12)."""
    i = np.arange(1, 256, dtype=np.float64)
    sdiv = np.zeros(256, np.int32)
    hdiv = np.zeros(256, np.int32)
    sdiv[1:] = np.rint((255 << HSV_SHIFT) / (1.0 * i)).astype(np.int32)
    hdiv[1:] = np.rint((180 << HSV_SHIFT) / (6.0 * i)).astype(np.int32)
    return sdiv, hdiv


_SDIV, _HDIV = hsv_tables()


def bgr2hsv_u8(img: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(img, COLOR_BGR2HSV) for uint8, H range 180."""
    b = img[..., 0].astype(np.int32)
    g = img[..., 1].astype(np.int32)
    r = img[..., 2].astype(np.int32)
    v = np.maximum(np.maximum(b, g), r)
    vmin = np.minimum(np.minimum(b, g), r)
    diff = v - vmin
    s = (diff * _SDIV[v] + (1 << (HSV_SHIFT - 1))) >> HSV_SHIFT
    h0 = np.where(v == r, g - b, np.where(v == g, b - r + 2 * diff, r - g + 4 * diff))
    h = (h0 * _HDIV[diff] + (1 << (HSV_SHIFT - 1))) >> HSV_SHIFT
    h = np.where(h < 0, h + 180, h)
    return np.stack([h, s, v], axis=-1).astype(np.uint8)


# --------------------------------------------------------------------------- A.7 (integer part)
def bgr2y_u8(img: np.ndarray) -> np.ndarray:
    """Y plane of cv2.cvtColor(img, COLOR_BGR2YUV) for uint8 (yuv_shift = 14)."""
    b = img[..., 0].astype(np.int32)
    g = img[..., 1].astype(np.int32)
    r = img[..., 2].astype(np.int32)
    return ((4899 * r + 9617 * g + 1868 * b + 8192) >> 14).astype(np.uint8)


def y_histogram(y: np.ndarray, bins: int = 256) -> np.ndarray:
    """cv2.calcHist([y],[0],None,[bins],[0,256]) as integer counts (uint32)."""
    idx = (y.astype(np.int64) * bins) >> 8
    return np.bincount(idx.ravel(), minlength=bins).astype(np.uint32)


# --------------------------------------------------------------------------- A.4 (integer part)
def plane_abs_sums(cur_hsv: np.ndarray, prev_hsv: np.ndarray) -> np.ndarray:
    """(sum|dH|, sum|dS|, sum|dV|) as int64 -- numerator of _mean_pixel_distance."""
    d = np.abs(cur_hsv.astype(np.int32) - prev_hsv.astype(np.int32))
    return d.reshape(-1, 3).sum(axis=0, dtype=np.int64)


# --------------------------------------------------------------------------- a14: ContentDetector._detect_edges
CANNY_SHIFT = 15
TG22 = int(0.4142135623730950488016887242097 * (1 << CANNY_SHIFT) + 0.5)


def estimated_kernel_size(frame_width: int, frame_height: int) -> int:
    """scenedetect.detectors.content_detector._estimated_kernel_size."""
    import math

    size = 4 + round(math.sqrt(frame_width * frame_height) / 192)
    if size % 2 == 0:
        size += 1
    return size


def canny_thresholds(lum: np.ndarray):
    """low/high of ContentDetector._detect_edges: sigma = 1/3 around numpy.median(lum), truncated to int."""
    sigma = 1.0 / 3.0
    median = np.median(lum)
    low = int(max(0, (1.0 - sigma) * median))
    high = int(min(255, (1.0 + sigma) * median))
    return low, high


def canny_u8(img: np.ndarray, low: int, high: int) -> np.ndarray:
    """cv2.Canny(img, low, high) (aperture 3, L1 gradient) restated from OpenCV canny.cpp:
    Sobel with BORDER_REPLICATE, |dx|+|dy|, non-maximum suppression with the TG22 fixed-point sector test
    against a zero-padded magnitude map, hysteresis = weak pixels 8-connected to a strong one."""
    if low > high:
        low, high = high, low
    img = img.astype(np.int32)
    p = np.pad(img, 1, mode="edge")
    dx = (p[:-2, 2:] + 2 * p[1:-1, 2:] + p[2:, 2:]) - (p[:-2, :-2] + 2 * p[1:-1, :-2] + p[2:, :-2])
    dy = (p[2:, :-2] + 2 * p[2:, 1:-1] + p[2:, 2:]) - (p[:-2, :-2] + 2 * p[:-2, 1:-1] + p[:-2, 2:])
    mag = np.abs(dx) + np.abs(dy)
    mp = np.pad(mag, 1, mode="constant")
    m = mp[1:-1, 1:-1]
    x = np.abs(dx)
    y = np.abs(dy) << CANNY_SHIFT
    tg22x = x * TG22
    tg67x = tg22x + (x << (CANNY_SHIFT + 1))
    left, right, up, down = mp[1:-1, :-2], mp[1:-1, 2:], mp[:-2, 1:-1], mp[2:, 1:-1]
    s_pos = (dx ^ dy) >= 0
    prev_d = np.where(s_pos, mp[:-2, :-2], mp[:-2, 2:])
    next_d = np.where(s_pos, mp[2:, 2:], mp[2:, :-2])
    horiz = (y < tg22x) & (m > left) & (m >= right)
    vert = (y > tg67x) & (m > up) & (m >= down)
    diag = ~(y < tg22x) & ~(y > tg67x) & (m > prev_d) & (m > next_d)
    nms = (m > low) & (horiz | vert | diag)
    out = nms & (m > high)
    while True:
        q = np.pad(out, 1)
        nb = (q[:-2, :-2] | q[:-2, 1:-1] | q[:-2, 2:] | q[1:-1, :-2] | q[1:-1, 2:] | q[2:, :-2] | q[2:, 1:-1] | q[2:, 2:])
        new = out | (nms & nb)
        if (new == out).all():
            break
        out = new
    return (out * 255).astype(np.uint8)


def dilate_ones_u8(img: np.ndarray, k: int) -> np.ndarray:
    """cv2.dilate(img, numpy.ones((k, k), uint8)): max over the k x k window centred on the pixel (anchor k//2),
    pixels outside the image ignored."""
    r0 = k // 2
    r1 = k - 1 - r0
    h, w = img.shape
    p = np.pad(img, ((r0, r1), (r0, r1)), mode="constant")
    out = np.zeros_like(img)
    for dy in range(k):
        for dx in range(k):
            out = np.maximum(out, p[dy:dy + h, dx:dx + w])
    return out


def detect_edges(lum: np.ndarray, kernel_size: int) -> np.ndarray:
    low, high = canny_thresholds(lum)
    return dilate_ones_u8(canny_u8(lum, low, high), kernel_size)
