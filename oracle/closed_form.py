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


# --------------------------------------------------------------------------- N4: HashDetector.hash_frame stages
def bgr2gray_u8(img: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(img, COLOR_BGR2GRAY) for uint8: 15-bit coefficients (OpenCV 4.x color_rgb, RGB2Gray<uchar>:
    BY15 = 3735, GY15 = 19235, RY15 = 9798), NOT the 14-bit Y of BGR2YUV.  Verified over all 2^24 BGR values
    against cv2 4.13 (tests/test_oracle.py)."""
    b = img[..., 0].astype(np.int32)
    g = img[..., 1].astype(np.int32)
    r = img[..., 2].astype(np.int32)
    return ((9798 * r + 19235 * g + 3735 * b + 16384) >> 15).astype(np.uint8)


def area_axis_table(ssize: int, dsize: int):
    """OpenCV resize.cpp computeResizeAreaTab for one axis (cn = 1): list of (dst index, src index, float32 weight)
    in the order the C++ loop visits them.  `scale` is a double; the weights are narrowed to float32."""
    import math

    scale = ssize / dsize
    tab = []
    for d in range(dsize):
        fsx1 = d * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1 = math.ceil(fsx1)
        sx2 = math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((d, sx1 - 1, np.float32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            tab.append((d, sx, np.float32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab.append((d, sx2, np.float32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return tab


def resize_area_u8(img: np.ndarray, dst_w: int, dst_h: int) -> np.ndarray:
    """cv2.resize(img, (dst_w, dst_h), interpolation=INTER_AREA) for one-channel uint8 when shrinking on both axes.

    * both scale factors integral: ResizeAreaFast_ -- integer box sum, `saturate_cast<uchar>(sum * float32(1/area))`
      (the 2x2 case is the SIMD special `(sum + 2) >> 2`);
    * otherwise ResizeArea_: per source row a float32 row buffer `buf[dx] += S[sx] * alpha` in table order, rows
      combined as `sum = beta * buf` / `sum += beta * buf`, float32 throughout, no FMA, cvRound at the end.
    Verified bit-exact against cv2 4.13 (IPP on and off) for a range of geometries (tests/test_oracle.py)."""
    h, w = img.shape
    if dst_w > w or dst_h > h:
        raise ValueError("resize_area_u8 only restates the shrinking case")
    sx, sy = w / dst_w, h / dst_h
    eps = np.finfo(np.float64).eps
    if abs(sx - int(sx)) < eps and abs(sy - int(sy)) < eps:
        isx, isy = int(sx), int(sy)
        s = img.reshape(dst_h, isy, dst_w, isx).astype(np.int64).sum(axis=(1, 3))
        if isx == 2 and isy == 2:
            return ((s + 2) >> 2).astype(np.uint8)
        v = s.astype(np.float32) * np.float32(1.0 / (isx * isy))
        return np.clip(np.rint(v), 0, 255).astype(np.uint8)
    xt = area_axis_table(w, dst_w)
    yt = area_axis_table(h, dst_h)
    f = img.astype(np.float32)
    res = np.zeros((dst_h, dst_w), np.float32)
    xd = np.array([t[0] for t in xt])
    xs = np.array([t[1] for t in xt])
    xa = np.array([t[2] for t in xt], np.float32)
    # position of every x entry inside its destination column (0, 1, 2, ...): entries of equal rank are independent
    rank = np.zeros(len(xt), np.int64)
    for i in range(1, len(xt)):
        rank[i] = rank[i - 1] + 1 if xd[i] == xd[i - 1] else 0
    prev = -1
    for (dy, sy_, beta) in yt:
        buf = np.zeros(dst_w, np.float32)
        for k in range(int(rank.max()) + 1):
            sel = rank == k
            buf[xd[sel]] = buf[xd[sel]] + f[sy_, xs[sel]] * xa[sel]
        if dy != prev:
            res[dy] = beta * buf
            prev = dy
        else:
            res[dy] = res[dy] + beta * buf
    return np.clip(np.rint(res), 0, 255).astype(np.uint8)


def dct_matrix(n: int) -> np.ndarray:
    """Orthonormal DCT-II basis C[u, i] (float64): cv2.dct(x) == C @ x @ C.T up to float32 rounding."""
    i = np.arange(n)
    c = np.sqrt(2.0 / n) * np.cos(np.pi * (2 * i[None, :] + 1) * i[:, None] / (2 * n))
    c[0, :] = np.sqrt(1.0 / n)
    return c


def hash_frame(frame_bgr: np.ndarray, hash_size: int = 16, factor: int = 2):
    """HashDetector.hash_frame with the DCT evaluated in float64 and narrowed to float32.

    Returns (bits bool[hash_size, hash_size], margin float64[hash_size, hash_size] = |coef - median|).  cv2.dct's
    float32 result depends on the OpenCV build (IPP vs. plain differ in ~70 % of the coefficients by 1-2 ulp), so
    only bits whose margin exceeds that noise are comparable across implementations."""
    gray = bgr2gray_u8(frame_bgr)
    imsize = hash_size * factor
    small = resize_area_u8(gray, imsize, imsize)
    max_value = int(small.max())
    if max_value == 0:
        max_value = 1
    x = small.astype(np.float32) / np.float32(max_value)
    c = dct_matrix(imsize)[:hash_size]
    low = (c @ x.astype(np.float64) @ c.T).astype(np.float32)
    med = np.median(low)
    return low > med, np.abs(low.astype(np.float64) - float(med))


# --------------------------------------------------------------------------- N1: decoder surfaces
def nv12_to_bgr_u8(nv12: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(nv12, COLOR_YUV2BGR_NV12) for a contiguous NV12 frame [H*3/2, W] uint8 -> [H, W, 3].

    OpenCV color_yuv (ITU-R BT.601, 20-bit fixed point): yy = max(0, Y-16)*1220542; R = sat((yy + 2^19 + 1673527 v) >> 20),
    G = sat((yy + 2^19 - 852492 v - 409993 u) >> 20), B = sat((yy + 2^19 + 2116026 u) >> 20) with u = U-128, v = V-128 taken
    from the chroma pair of the pixel's 2x2 block.  Verified over all 2^24 (Y, U, V) against cv2 4.13 (tests/test_oracle.py)."""
    rows, w = nv12.shape
    h = rows * 2 // 3
    y = nv12[:h].astype(np.int64)
    uv = nv12[h:].astype(np.int64)
    u = np.repeat(np.repeat(uv[:, 0::2], 2, 0), 2, 1) - 128
    v = np.repeat(np.repeat(uv[:, 1::2], 2, 0), 2, 1) - 128
    yy = np.maximum(0, y - 16) * 1220542
    r = np.clip((yy + (1 << 19) + 1673527 * v) >> 20, 0, 255)
    g = np.clip((yy + (1 << 19) - 852492 * v - 409993 * u) >> 20, 0, 255)
    b = np.clip((yy + (1 << 19) + 2116026 * u) >> 20, 0, 255)
    return np.stack([b, g, r], -1).astype(np.uint8)
