"""CPU ORACLE (test infrastructure only) -- PySceneDetect detector logic restated.

"PySceneDetect with OpenCV, minus the package import": the ContentDetector /
AdaptiveDetector / HistogramDetector ``process_frame`` logic, the FlashFilter
state machine, SceneManager's auto-downscale and ``get_scenes_from_cuts``,
restated from PySceneDetect 0.6.4+ semantics
and executed over *real* ``cv2`` calls (backend="cv2") or over the closed-form
integer restatement in ``oracle/closed_form.py`` (backend="closed_form").

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may
import this module; the product never does.

PARITY STATUS: "parity unpinned" w.r.t. the ``scenedetect`` package: it is not
vendored, pinned or installable here.  The reference
call site this path replaces is ``ModelManager.detect_scenes``
(/root/reference/ml-service/src/services/model_manager.py:715-835); the
producer it names is "scenedetect" (ml-service/src/models/responses.py:141-142).
The cv2 backend pins the integer/float arithmetic to OpenCV 4.13.0.
"""
from __future__ import annotations

import math
from typing import Iterable, List, NamedTuple, Optional, Sequence

import numpy as np

from . import closed_form as cf

try:  # cv2 exists in this image; the closed-form backend works without it
    import cv2  # type: ignore
except Exception:  # pragma: no cover
    cv2 = None


# ----------------------------------------------------------------------------- primitives
def _resize(img: np.ndarray, dst_w: int, dst_h: int, backend: str) -> np.ndarray:
    if backend == "cv2":
        return cv2.resize(img, (dst_w, dst_h), interpolation=cv2.INTER_LINEAR)
    return cf.resize_linear_u8(img, dst_w, dst_h)


def _hsv_planes(img: np.ndarray, backend: str):
    if backend == "cv2":
        return cv2.split(cv2.cvtColor(img, cv2.COLOR_BGR2HSV))
    hsv = cf.bgr2hsv_u8(img)
    return hsv[..., 0], hsv[..., 1], hsv[..., 2]


def normalize_l2_f32(hist_counts: np.ndarray) -> np.ndarray:
    """cv2.normalize(hist, hist) (NORM_L2, alpha=1) on a float32 histogram.

    Norm accumulated in double (exact for integer counts), 1/norm in double,
    narrowed to float32, product in float32 (probe-verified).
    """
    h = hist_counts.astype(np.float32)
    ssq = float(np.sum(h.astype(np.float64) ** 2))
    nrm = math.sqrt(ssq)
    scale = (1.0 / nrm) if nrm > np.finfo(np.float64).eps else 0.0
    return (h * np.float32(scale)).astype(np.float32)


def compare_hist_correl(h1: np.ndarray, h2: np.ndarray) -> float:
    """cv2.compareHist(h1, h2, HISTCMP_CORREL) for float32 1-D histograms.

    Double accumulation in two interleaved lanes (even / odd index) added at
    the end -- the order of this cv2 build's SSE2-baseline histogram.cpp
    (SURVEY.md A.7).  No FMA.
    """
    a = h1.astype(np.float64).ravel()
    b = h2.astype(np.float64).ravel()
    n = a.size
    lanes = 2
    s1 = [0.0] * lanes
    s2 = [0.0] * lanes
    s11 = [0.0] * lanes
    s12 = [0.0] * lanes
    s22 = [0.0] * lanes
    nv = (n // 4) * 4  # the vector loop consumes 4 float32 per iteration (v_float32x4)
    for j in range(0, nv, lanes):
        for k in range(lanes):
            x = float(a[j + k])
            y = float(b[j + k])
            s12[k] += x * y
            s1[k] += x
            s11[k] += x * x
            s2[k] += y
            s22[k] += y * y
    S1 = s1[0] + s1[1]
    S2 = s2[0] + s2[1]
    S11 = s11[0] + s11[1]
    S12 = s12[0] + s12[1]
    S22 = s22[0] + s22[1]
    for j in range(nv, n):  # scalar tail
        x = float(a[j])
        y = float(b[j])
        S12 += x * y
        S1 += x
        S11 += x * x
        S2 += y
        S22 += y * y
    scale = 1.0 / n
    num = S12 - S1 * S2 * scale
    den2 = (S11 - S1 * S1 * scale) * (S22 - S2 * S2 * scale)
    return num / math.sqrt(den2) if abs(den2) > np.finfo(np.float64).eps else 1.0


def calculate_histogram(img: np.ndarray, bins: int, backend: str):
    """HistogramDetector.calculate_histogram -> (counts uint32[bins], normalized float32[bins])."""
    if backend == "cv2":
        y, _, _ = cv2.split(cv2.cvtColor(img, cv2.COLOR_BGR2YUV))
        raw = cv2.calcHist([y], [0], None, [bins], [0, 256])
        counts = raw.flatten().astype(np.uint32)
        hist = cv2.normalize(raw, raw.copy()).flatten()
        return counts, hist
    y = cf.bgr2y_u8(img)
    counts = cf.y_histogram(y, bins)
    return counts, normalize_l2_f32(counts)


def _compare(h1, h2, backend):
    if backend == "cv2":
        return cv2.compareHist(h1, h2, cv2.HISTCMP_CORREL)
    return compare_hist_correl(h1, h2)


# ----------------------------------------------------------------------------- FlashFilter (A.5)
FILTER_MERGE = 0
FILTER_SUPPRESS = 1


class FlashFilter:
    """scenedetect.scene_detector.FlashFilter (0.6.4+); SUPPRESS == legacy min_scene_len."""

    def __init__(self, mode: int, length: int):
        self._mode = mode
        self._filter_length = length
        self._last_above = None
        self._merge_enabled = False
        self._merge_triggered = False
        self._merge_start = None

    def filter(self, frame_num: int, above_threshold: bool) -> List[int]:
        if not self._filter_length > 0:
            return [frame_num] if above_threshold else []
        if self._last_above is None:
            self._last_above = frame_num
        if self._mode == FILTER_MERGE:
            return self._filter_merge(frame_num, above_threshold)
        if self._mode == FILTER_SUPPRESS:
            return self._filter_suppress(frame_num, above_threshold)
        raise RuntimeError("Unhandled FlashFilter mode.")

    def _filter_suppress(self, frame_num, above_threshold):
        min_length_met = (frame_num - self._last_above) >= self._filter_length
        if not (above_threshold and min_length_met):
            return []
        self._last_above = frame_num
        return [frame_num]

    def _filter_merge(self, frame_num, above_threshold):
        min_length_met = (frame_num - self._last_above) >= self._filter_length
        if above_threshold:
            self._last_above = frame_num
        if self._merge_triggered:
            num_merged_frames = self._last_above - self._merge_start
            if min_length_met and not above_threshold and num_merged_frames >= self._filter_length:
                self._merge_triggered = False
                return [self._last_above]
            return []
        if not above_threshold:
            return []
        if min_length_met:
            self._merge_enabled = True
            return [frame_num]
        if self._merge_enabled:
            self._merge_triggered = True
            self._merge_start = frame_num
        return []


# ----------------------------------------------------------------------------- detectors
class Components(NamedTuple):
    delta_hue: float = 1.0
    delta_sat: float = 1.0
    delta_lum: float = 1.0
    delta_edges: float = 0.0


DEFAULT_COMPONENT_WEIGHTS = Components()
LUMA_ONLY_WEIGHTS = Components(0.0, 0.0, 1.0, 0.0)


def _mean_pixel_distance(left: np.ndarray, right: np.ndarray):
    num_pixels = float(left.shape[0] * left.shape[1])
    return np.sum(np.abs(left.astype(np.int32) - right.astype(np.int32))) / num_pixels


class ContentDetector:
    """scenedetect.detectors.ContentDetector.process_frame (A.4, A.5)."""

    def __init__(self, threshold=27.0, min_scene_len=15, weights=DEFAULT_COMPONENT_WEIGHTS,
                 luma_only=False, kernel_size=None, filter_mode=FILTER_MERGE, backend="cv2"):
        self._threshold = threshold
        self._min_scene_len = min_scene_len
        self._weights = Components(*weights)
        if luma_only:
            self._weights = LUMA_ONLY_WEIGHTS
        if kernel_size is not None and (kernel_size < 3 or kernel_size % 2 == 0):
            raise ValueError("kernel_size must be odd integer >= 3")
        self._kernel_size = kernel_size
        self._last = None
        self.edge_sums: list = []  # number of differing edge pixels x 255 (0 for the first frame / when unused)
        self._frame_score = None
        self._flash_filter = FlashFilter(filter_mode, min_scene_len)
        self._backend = backend
        # recorded per processed frame, for parity tests
        self.sums: list = []  # int64[3] (0 for the first frame)
        self.scores: list = []  # content_val (float64)

    def _detect_edges(self, lum):
        """ContentDetector._detect_edges: Canny around the median + dilate."""
        if self._kernel_size is None:
            self._kernel_size = cf.estimated_kernel_size(lum.shape[1], lum.shape[0])
        if self._backend == "cv2":
            low, high = cf.canny_thresholds(lum)
            edges = cv2.Canny(lum, low, high)
            return cv2.dilate(edges, np.ones((self._kernel_size, self._kernel_size), np.uint8))
        return cf.detect_edges(np.ascontiguousarray(lum), self._kernel_size)

    def _calculate_frame_score(self, frame_num, frame_img):
        hue, sat, lum = _hsv_planes(frame_img, self._backend)
        calculate_edges = self._weights.delta_edges > 0.0
        edges = self._detect_edges(lum) if calculate_edges else None
        if self._last is None:
            self._last = (hue, sat, lum, edges)
            self.sums.append(np.zeros(3, np.int64))
            self.edge_sums.append(0)
            return 0.0
        comps = Components(
            delta_hue=_mean_pixel_distance(hue, self._last[0]),
            delta_sat=_mean_pixel_distance(sat, self._last[1]),
            delta_lum=_mean_pixel_distance(lum, self._last[2]),
            delta_edges=(0.0 if edges is None else _mean_pixel_distance(edges, self._last[3])),
        )
        self.edge_sums.append(0 if edges is None else int(np.sum(np.abs(edges.astype(np.int32) - self._last[3].astype(np.int32)))))
        n = hue.shape[0] * hue.shape[1]
        self.sums.append(np.array([int(round(float(c) * n)) for c in comps[:3]], np.int64))
        score = sum(c * w for (c, w) in zip(comps, self._weights)) / sum(abs(w) for w in self._weights)
        self._last = (hue, sat, lum, edges)
        return score

    def process_frame(self, frame_num, frame_img) -> List[int]:
        self._frame_score = self._calculate_frame_score(frame_num, frame_img)
        self.scores.append(float(self._frame_score))
        above = self._frame_score >= self._threshold
        return self._flash_filter.filter(frame_num=frame_num, above_threshold=bool(above))

    def post_process(self, frame_num) -> List[int]:
        return []

    @property
    def event_buffer_length(self):
        return 0


class AdaptiveDetector(ContentDetector):
    """scenedetect.detectors.AdaptiveDetector.process_frame (A.6)."""

    def __init__(self, adaptive_threshold=3.0, min_scene_len=15, window_width=2, min_content_val=15.0,
                 weights=DEFAULT_COMPONENT_WEIGHTS, luma_only=False, kernel_size=None, backend="cv2"):
        if window_width < 1:
            raise ValueError("window_width must be at least 1.")
        super().__init__(threshold=255.0, min_scene_len=0, weights=weights, luma_only=luma_only,
                         kernel_size=kernel_size, backend=backend)
        self.min_scene_len = min_scene_len
        self.adaptive_threshold = adaptive_threshold
        self.min_content_val = min_content_val
        self.window_width = window_width
        self._last_cut: Optional[int] = None
        self._buffer: list = []
        self.ratios: dict = {}  # target_frame -> adaptive_ratio

    @property
    def event_buffer_length(self):
        return self.window_width

    def process_frame(self, frame_num, frame_img) -> List[int]:
        super().process_frame(frame_num=frame_num, frame_img=frame_img)
        if self._last_cut is None:
            self._last_cut = frame_num
        required = 1 + 2 * self.window_width
        self._buffer.append((frame_num, self._frame_score))
        if not len(self._buffer) >= required:
            return []
        self._buffer = self._buffer[-required:]
        target_frame, target_score = self._buffer[self.window_width]
        avg = sum(s for i, (_f, s) in enumerate(self._buffer) if i != self.window_width) / (2.0 * self.window_width)
        average_is_zero = abs(avg) < 0.00001
        ratio = 0.0
        if not average_is_zero:
            ratio = min(target_score / avg, 255.0)
        elif average_is_zero and target_score >= self.min_content_val:
            ratio = 255.0
        self.ratios[target_frame] = float(ratio)
        threshold_met = ratio >= self.adaptive_threshold and target_score >= self.min_content_val
        min_length_met = (frame_num - self._last_cut) >= self.min_scene_len
        if threshold_met and min_length_met:
            self._last_cut = target_frame
            return [target_frame]
        return []


class HistogramDetector:
    """scenedetect.detectors.HistogramDetector.process_frame (A.7)."""

    def __init__(self, threshold=0.05, bins=256, min_scene_len=15, backend="cv2"):
        self._threshold = max(0.0, min(1.0, 1.0 - threshold))
        self._bins = bins
        self._min_scene_len = min_scene_len
        self._last_hist = None
        self._last_scene_cut = None
        self._backend = backend
        self.counts: list = []  # uint32[bins] per frame
        self.diffs: list = []  # hist_diff per frame (nan for the first)

    def process_frame(self, frame_num, frame_img) -> List[int]:
        cut_list = []
        if frame_img.dtype != np.uint8:
            raise ValueError("Image must be 8-bit rgb for HistogramDetector")
        if frame_img.shape[2] != 3:
            raise ValueError("Image must have three color channels for HistogramDetector")
        if not self._last_scene_cut:
            self._last_scene_cut = frame_num
        counts, hist = calculate_histogram(frame_img, self._bins, self._backend)
        self.counts.append(counts)
        if self._last_hist is not None:
            hist_diff = _compare(self._last_hist, hist, self._backend)
            self.diffs.append(float(hist_diff))
            if hist_diff <= self._threshold and (frame_num - self._last_scene_cut) >= self._min_scene_len:
                cut_list.append(frame_num)
                self._last_scene_cut = frame_num
        else:
            self.diffs.append(float("nan"))
        self._last_hist = hist
        return cut_list

    def post_process(self, frame_num) -> List[int]:
        return []

    @property
    def event_buffer_length(self):
        return 0


class ThresholdDetector:
    """scenedetect.detectors.ThresholdDetector.process_frame / post_process (0.6.4)
    [upstream-recall: restated from knowledge of the package, parity unpinned]."""

    FLOOR, CEILING = 0, 1

    def __init__(self, threshold=12, min_scene_len=15, fade_bias=0.0, add_final_scene=False, method=0, backend="cv2"):
        self.threshold = int(threshold)
        self.method = method
        self.fade_bias = fade_bias
        self.min_scene_len = min_scene_len
        self.processed_frame = False
        self.last_scene_cut = None
        self.add_final_scene = add_final_scene
        self.last_fade = {"frame": 0, "type": None}
        self.averages: list = []

    def process_frame(self, frame_num, frame_img) -> List[int]:
        if self.last_scene_cut is None:
            self.last_scene_cut = frame_num
        cut_list = []
        num_pixel_values = float(frame_img.shape[0] * frame_img.shape[1] * frame_img.shape[2])
        frame_avg = np.sum(frame_img[:, :, :]) / num_pixel_values
        self.averages.append(float(frame_avg))
        if self.processed_frame:
            if self.last_fade["type"] == "in" and (
                    (self.method == self.FLOOR and frame_avg < self.threshold)
                    or (self.method == self.CEILING and frame_avg >= self.threshold)):
                self.last_fade["type"] = "out"
                self.last_fade["frame"] = frame_num
            elif self.last_fade["type"] == "out" and (
                    (self.method == self.FLOOR and frame_avg >= self.threshold)
                    or (self.method == self.CEILING and frame_avg < self.threshold)):
                if (frame_num - self.last_scene_cut) >= self.min_scene_len:
                    f_out = self.last_fade["frame"]
                    f_split = int((frame_num + f_out + int(self.fade_bias * (frame_num - f_out))) / 2)
                    cut_list.append(f_split)
                    self.last_scene_cut = frame_num
                self.last_fade["type"] = "in"
                self.last_fade["frame"] = frame_num
        else:
            self.last_fade["frame"] = 0
            if frame_avg < self.threshold:
                self.last_fade["type"] = "out"
            else:
                self.last_fade["type"] = "in"
        self.processed_frame = True
        return cut_list

    def post_process(self, frame_num) -> List[int]:
        cut_times = []
        if (self.last_fade["type"] == "out" and self.add_final_scene and (
                (self.last_scene_cut is None and frame_num >= self.min_scene_len)
                or (frame_num - self.last_scene_cut) >= self.min_scene_len)):
            cut_times.append(self.last_fade["frame"])
        return cut_times

    @property
    def event_buffer_length(self):
        return 0


class HashDetector:
    """scenedetect.detectors.HashDetector.process_frame / hash_frame (0.6.4)
    [upstream-recall: restated from knowledge of the package, parity unpinned].

    Perceptual hash: gray -> INTER_AREA resize to (size*lowpass)^2 -> / max -> cv2.dct -> low-frequency
    size x size block -> bits = coefficient > median; the score is the Hamming distance to the previous frame's
    hash divided by size^2.  cv2.dct's float32 output is build-dependent (IPP), so besides the bits this class
    records each bit's margin |coef - median| for tolerance-aware comparisons."""

    def __init__(self, threshold=0.395, size=16, lowpass=2, min_scene_len=15, backend="cv2"):
        self._threshold = threshold
        self._min_scene_len = min_scene_len
        self._size = size
        self._size_sq = float(size * size)
        self._factor = lowpass
        self._last_frame = None
        self._last_scene_cut = None
        self._last_hash = np.array([])
        self._backend = backend
        self.hashes: list = []   # bool[size, size] per frame
        self.margins: list = []  # float64[size, size] |coef - median| per frame
        self.dists: list = []    # hash_dist_norm per frame (nan for the first)

    def get_metrics(self):
        return [f"hash_dist [size={self._size} lowpass={self._factor}]"]

    def hash_frame(self, frame_img):
        if self._backend != "cv2":
            return cf.hash_frame(frame_img, self._size, self._factor)
        gray_img = cv2.cvtColor(frame_img, cv2.COLOR_BGR2GRAY)
        imsize = self._size * self._factor
        resized_img = cv2.resize(gray_img, (imsize, imsize), interpolation=cv2.INTER_AREA)
        max_value = np.max(np.max(resized_img))
        if max_value == 0:
            max_value = 1
        resized_img = np.float32(resized_img) / max_value
        dct_complete = cv2.dct(resized_img)
        dct_low_freq = dct_complete[:self._size, :self._size]
        med = np.median(dct_low_freq)
        return dct_low_freq > med, np.abs(dct_low_freq.astype(np.float64) - float(med))

    def process_frame(self, frame_num, frame_img) -> List[int]:
        cut_list = []
        if self._last_scene_cut is None:
            self._last_scene_cut = frame_num
        # the reference hashes a frame when it is compared; hashing every frame as it arrives yields the same values
        curr_hash, margin = self.hash_frame(frame_img)
        self.hashes.append(curr_hash)
        self.margins.append(margin)
        if self._last_frame is not None:
            last_hash = self._last_hash
            hash_dist = np.count_nonzero(curr_hash.flatten() != last_hash.flatten())
            hash_dist_norm = hash_dist / self._size_sq
            self.dists.append(float(hash_dist_norm))
            if hash_dist_norm >= self._threshold and (frame_num - self._last_scene_cut) >= self._min_scene_len:
                cut_list.append(frame_num)
                self._last_scene_cut = frame_num
        else:
            self.dists.append(float("nan"))
        self._last_hash = curr_hash
        self._last_frame = True
        return cut_list

    def post_process(self, frame_num) -> List[int]:
        return []

    @property
    def event_buffer_length(self):
        return 0


# ----------------------------------------------------------------------------- SceneManager (A.1, A.8)
def get_scenes_from_cuts(cut_list: Sequence[int], start_pos: int, end_pos: int):
    """scenedetect.scene_manager.get_scenes_from_cuts on frame numbers."""
    if not cut_list:
        return [(start_pos, end_pos)]
    scenes = []
    last = start_pos
    for cut in cut_list:
        scenes.append((last, cut))
        last = cut
    scenes.append((last, end_pos))
    return scenes


def scenes_to_dicts(scenes: Sequence[tuple], fps: float) -> list:
    """ml-service scene dicts; keys/units per model_manager.py:775-781, int() truncation per :771."""
    out = []
    for i, (a, b) in enumerate(scenes):
        s = int(a / fps * 1000)
        e = int(b / fps * 1000)
        out.append({"scene_index": i, "start_ms": s, "end_ms": e, "duration_ms": e - s})
    return out


def detect(frames: Iterable[np.ndarray], detectors: Sequence, *, auto_downscale=True, downscale=None,
           downscale_mode="float", backend="cv2", start_frame=0):
    """SceneManager.detect_scenes: per-frame downscale + detector loop.

    Returns (cut_list sorted/deduped, n_frames).
    """
    cuts: list = []
    factor = None
    n = 0
    for k, frame in enumerate(frames):
        if factor is None:
            if auto_downscale:
                factor = cf.compute_downscale_factor(frame.shape[1], mode=downscale_mode)
            else:
                factor = downscale if downscale else 1
        if factor > 1:
            dw, dh = cf.downscaled_size(frame.shape[1], frame.shape[0], factor)
            frame = _resize(frame, dw, dh, backend)
        for det in detectors:
            cuts += det.process_frame(start_frame + k, frame)
        n += 1
    for det in detectors:
        cuts += det.post_process(start_frame + n - 1)
    return sorted(set(cuts)), n


def detect_scenes_dicts(frames, detectors, fps, start_in_scene=True, **kw):
    """detect() + get_scene_list + ml-service dict glue."""
    cuts, n = detect(frames, detectors, **kw)
    if not cuts and not start_in_scene:
        return {"scenes": []}
    return {"scenes": scenes_to_dicts(get_scenes_from_cuts(cuts, 0, n), fps)}
