"""SceneDetector plugin surface (PySceneDetect's ContentDetector / AdaptiveDetector /
HistogramDetector: same constructor keywords, ``process_frame`` / ``post_process`` /
``get_metrics`` / ``event_buffer_length``), executed by libesd.so on a B200.

The reference names this surface in /root/reference/README.md:56 and
.kiro/specs/semantic-video-search/design.md:994-1007 and its ml-service would call it from
``ModelManager.detect_scenes`` (ml-service/src/services/model_manager.py:715-835).

Two ways to run a detector:
  * stand-alone, frame by frame: ``det.process_frame(frame_num, frame_img)`` exactly like
    PySceneDetect (frames arrive already downscaled; numpy or CUDA torch uint8 BGR);
  * batched through :class:`eioku_b200.scene_manager.SceneManager`, which fuses the
    SceneManager downscale into the scoring kernel and feeds all detectors from one pass.
There is no CPU path: without libesd.so and a GPU these classes raise.
"""
from __future__ import annotations

from enum import Enum
from typing import Dict, List, NamedTuple, Optional

import numpy as np

from . import capi


class FlashFilter:
    """Namespace mirror of scenedetect.scene_detector.FlashFilter (only ``Mode`` is host-side;
    the state machine itself runs on the device, csrc/esd_kernels.cuh:decide_kernel)."""

    class Mode(Enum):
        MERGE = 0
        SUPPRESS = 1


class StatsManager:
    """Stand-in for scenedetect.stats_manager.StatsManager: per-frame metric store + the stats-file CSV writer
    (``Frame Number,Timecode,<metric keys sorted>``, 1-based frame numbers, one row per frame holding metrics)."""

    def __init__(self, fps: float = 30.0):
        self._frame_metrics: Dict[int, Dict[str, float]] = {}
        self._metric_keys: List[str] = []
        self.fps = float(fps)  # for the Timecode column; SceneManager.detect_scenes sets it from the video

    def register_metrics(self, keys):
        for k in keys:
            if k not in self._metric_keys:
                self._metric_keys.append(k)

    def set_metrics(self, frame_number: int, metric_kv_dict: Dict[str, float]):
        self._frame_metrics.setdefault(frame_number, {}).update(metric_kv_dict)
        self.register_metrics(metric_kv_dict.keys())

    def get_metrics(self, frame_number: int, metric_keys):
        return [self._frame_metrics.get(frame_number, {}).get(k) for k in metric_keys]

    def metrics_exist(self, frame_number: int, metric_keys) -> bool:
        d = self._frame_metrics.get(frame_number, {})
        return all(k in d for k in metric_keys)

    @property
    def metric_keys(self):
        return list(self._metric_keys)

    def save_to_csv(self, csv_file, fps: Optional[float] = None):
        """scenedetect.StatsManager.save_to_csv [upstream-recall]: header ``Frame Number,Timecode`` + the sorted metric
        keys; per frame the 1-based frame number, its HH:MM:SS.nnn timecode and ``str(metric)`` per key."""
        import csv

        from .service import frame_to_timecode

        rate = float(fps or self.fps)
        own = isinstance(csv_file, (str, bytes))
        f = open(csv_file, "w", newline="") if own else csv_file
        try:
            wr = csv.writer(f, lineterminator="\n")
            keys = sorted(self._metric_keys)
            wr.writerow(["Frame Number", "Timecode"] + keys)
            for fn in sorted(self._frame_metrics):
                row = self._frame_metrics[fn]
                wr.writerow([fn + 1, frame_to_timecode(fn, rate)] + [str(row.get(k)) for k in keys])
        finally:
            if own:
                f.close()


class SceneDetector:
    """Base of the plugin surface (scenedetect.scene_detector.SceneDetector)."""

    stats_manager: Optional[StatsManager] = None
    _DET_FLAG = 0

    def __init__(self):
        self.stats_manager = None
        self._ctx: Optional[capi.EsdContext] = None
        self._manager_ctx: Optional[capi.EsdContext] = None  # set by SceneManager while it drives this detector
        self._cuts_seen = 0
        self._device = 0
        # deferred mode (see `defer`): frames staged on the host and scored `_defer` at a time
        self._defer = 0
        self._stage = None
        self._stage_n = 0
        self._stage_first = 0

    def defer(self, frames: int) -> "SceneDetector":
        """Opt-in deferred scoring for frame-by-frame callers: ``process_frame`` only stages the (host) frame and every
        `frames` calls one batch is scored, so cuts are reported up to `frames` - 1 calls late (with their exact frame
        numbers; ``event_buffer_length`` grows accordingly and ``post_process`` flushes the rest) -- the contract
        PySceneDetect already gives detectors that look ahead, e.g. AdaptiveDetector.  0 / 1 = score every call."""
        if self._stage_n:
            raise RuntimeError("defer() must be called before frames are staged")
        self._defer = max(0, int(frames))
        return self

    # ---- PySceneDetect API
    def is_processing_required(self, frame_num: int) -> bool:
        return True

    def stats_manager_required(self) -> bool:
        return False

    def get_metrics(self) -> List[str]:
        return []

    def post_process(self, frame_num: int) -> List[int]:
        return self._flush_staged()

    @property
    def event_buffer_length(self) -> int:
        return self._own_buffer_length() + max(0, self._defer - 1)

    def _own_buffer_length(self) -> int:
        return 0

    # ---- config plumbing shared with SceneManager
    def _fill_config(self, cfg: capi.EsdConfig) -> None:
        raise NotImplementedError

    def _metrics_for(self, scores: dict, k: int) -> Dict[str, float]:
        return {}

    def _make_ctx(self, width: int, height: int, device: int) -> capi.EsdContext:
        cfg = capi.default_config()
        cfg.detectors = 0
        self._fill_config(cfg)
        cfg.src_width, cfg.src_height = width, height
        cfg.dst_width, cfg.dst_height = width, height  # stand-alone detectors never resize (SceneManager does)
        if width * height <= 512 * 512:
            # frame-by-frame callers: smaller row groups spread one small frame over more CTAs (-5 us per call)
            cfg.rows_per_group = 4
        return capi.EsdContext(cfg, device)

    def _validate_frame(self, frame_img):
        pass

    def _to_device(self, frame_img):
        import torch

        if isinstance(frame_img, np.ndarray):
            self._validate_frame(frame_img)
            return torch.from_numpy(np.ascontiguousarray(frame_img)).to(f"cuda:{self._device}", non_blocking=False)
        if not frame_img.is_cuda:
            frame_img = frame_img.to(f"cuda:{self._device}")
        self._validate_frame(frame_img)
        return frame_img

    def process_frames(self, first_frame_num: int, frames) -> List[int]:
        """Batched process_frame: frames [N,H,W,3] uint8 BGR (numpy or CUDA tensor), already at detector
        resolution.  Returns every cut the N frames produced, in emission order."""
        t = self._to_device(frames)
        if t.dim() == 3:
            t = t.unsqueeze(0)
        if self._ctx is None:
            self._device = t.device.index
            self._ctx = self._make_ctx(t.shape[2], t.shape[1], self._device)
        self._ctx.push_tensor(t, first_frame_num)
        cuts, total = self._ctx.get_cuts(self._DET_FLAG, self._cuts_seen)
        self._cuts_seen = total
        if self.stats_manager is not None:
            self._publish_metrics(first_frame_num, t.shape[0])
        return cuts

    def _publish_late_metrics(self, first_frame_num: int, n: int):
        pass

    def process_frame(self, frame_num: int, frame_img) -> List[int]:
        """PySceneDetect's per-frame entry point.  A numpy frame takes the library's single-call fast path
        (esd_process_frame_host: pinned staging, one stream, one synchronisation); CUDA tensors and batches go
        through ``process_frames``."""
        if not isinstance(frame_img, np.ndarray) or frame_img.ndim != 3:
            return self.process_frames(frame_num, frame_img)
        self._validate_frame(frame_img)
        if self._defer > 1:
            return self._stage_frame(frame_num, frame_img)
        if self._ctx is None:
            self._ctx = self._make_ctx(frame_img.shape[1], frame_img.shape[0], self._device)
        cuts, total = self._ctx.process_frame_host(frame_img, frame_num, self._DET_FLAG, self._cuts_seen)
        self._cuts_seen = total
        if self.stats_manager is not None:
            self._publish_metrics(frame_num, 1)
        return cuts

    def _publish_metrics(self, first_frame_num: int, n: int):
        sc = self._ctx.read_scores(first_frame_num, n)
        # the first frame of the video has no predecessor: PySceneDetect's _calculate_frame_score returns before it
        # stores any content metric for it
        first_video_frame = first_frame_num + n - self._ctx.frames_pushed
        for k in range(n):
            if first_frame_num + k == first_video_frame and isinstance(self, ContentDetector):
                continue
            m = self._metrics_for(sc, k)
            if m:
                self.stats_manager.set_metrics(first_frame_num + k, m)
        self._publish_late_metrics(first_frame_num, n)

    def _stage_frame(self, frame_num: int, frame_img: np.ndarray) -> List[int]:
        import torch

        if self._stage is None:
            h, w, _ = frame_img.shape
            self._stage = torch.empty((self._defer, h, w, 3), dtype=torch.uint8).pin_memory()
            self._stage_np = self._stage.numpy()
            self._stage_dev = torch.empty((self._defer, h, w, 3), dtype=torch.uint8, device=f"cuda:{self._device}")
        if self._stage_n == 0:
            self._stage_first = frame_num
        elif frame_num != self._stage_first + self._stage_n:
            raise ValueError("deferred process_frame needs sequential frame numbers")
        np.copyto(self._stage_np[self._stage_n], frame_img)
        self._stage_n += 1
        return self._flush_staged() if self._stage_n == self._defer else []

    def _flush_staged(self) -> List[int]:
        n = self._stage_n
        if n == 0:
            return []
        self._stage_n = 0
        dev = self._stage_dev[:n]
        dev.copy_(self._stage[:n], non_blocking=True)
        return self.process_frames(self._stage_first, dev)

    def close(self):
        if self._ctx is not None:
            self._ctx.close()
            self._ctx = None


class ContentDetector(SceneDetector):
    """scenedetect.detectors.ContentDetector: weighted mean |delta| of the H, S, V planes between
    consecutive frames, cut when ``content_val >= threshold`` subject to the flash filter."""

    class Components(NamedTuple):
        delta_hue: float = 1.0
        delta_sat: float = 1.0
        delta_lum: float = 1.0
        delta_edges: float = 0.0

    DEFAULT_COMPONENT_WEIGHTS = Components()
    LUMA_ONLY_WEIGHTS = Components(delta_hue=0.0, delta_sat=0.0, delta_lum=1.0, delta_edges=0.0)
    FRAME_SCORE_KEY = "content_val"
    METRIC_KEYS = [FRAME_SCORE_KEY, *Components._fields]
    _DET_FLAG = capi.ESD_DET_CONTENT

    def __init__(self, threshold: float = 27.0, min_scene_len: int = 15,
                 weights: "ContentDetector.Components" = DEFAULT_COMPONENT_WEIGHTS, luma_only: bool = False,
                 kernel_size: Optional[int] = None, filter_mode: FlashFilter.Mode = FlashFilter.Mode.MERGE):
        super().__init__()
        self._threshold = threshold
        self._min_scene_len = min_scene_len
        self._weights = ContentDetector.Components(*weights)
        if luma_only:
            self._weights = ContentDetector.LUMA_ONLY_WEIGHTS
        if kernel_size is not None and (kernel_size < 3 or kernel_size % 2 == 0):
            raise ValueError("kernel_size must be odd integer >= 3")
        self._kernel_size = kernel_size
        self._filter_mode = FlashFilter.Mode(filter_mode) if not isinstance(filter_mode, FlashFilter.Mode) else filter_mode

    def get_metrics(self):
        return ContentDetector.METRIC_KEYS

    def _weight_div(self) -> float:
        # exactly PySceneDetect's divisor expression, evaluated by the host interpreter
        return float(sum(abs(w) for w in self._weights))

    def _fill_edge_config(self, cfg):
        # delta_edges > 0 switches the Canny + dilate edge map on (ContentDetector._detect_edges); one kernel size per pass
        if self._weights.delta_edges > 0.0 and self._kernel_size is not None:
            if cfg.edge_kernel_size not in (0, self._kernel_size):
                raise ValueError("detectors fed from one pass must share kernel_size")
            cfg.edge_kernel_size = int(self._kernel_size)

    def _fill_config(self, cfg):
        self._fill_edge_config(cfg)
        cfg.detectors |= capi.ESD_DET_CONTENT
        cfg.content_threshold = float(self._threshold)
        for i, w in enumerate(self._weights):
            cfg.content_weights[i] = float(w)
        cfg.content_weight_div = self._weight_div()
        cfg.content_min_scene_len = int(self._min_scene_len)
        cfg.content_filter_mode = int(self._filter_mode.value)

    def _metrics_for(self, scores, k):
        npx = float(self._ctx.geometry.dst_width * self._ctx.geometry.dst_height) if self._ctx else 1.0
        return _content_metrics(scores, k, "content_val", npx)

    def _publish_late_metrics(self, first_frame_num, n):
        if self._weights.delta_edges > 0.0:
            npx = float(self._ctx.geometry.dst_width * self._ctx.geometry.dst_height)
            cnt = self._ctx.read_edge_counts(first_frame_num, n)
            for k in range(n):
                self.stats_manager.set_metrics(first_frame_num + k, {"delta_edges": float(np.int64(255 * int(cnt[k])) / npx)})


def _content_metrics(scores, k, val_key, npx):
    s = scores["sums3"][k]
    return {
        ContentDetector.FRAME_SCORE_KEY: float(scores[val_key][k]),
        "delta_hue": float(np.int64(s[0]) / npx),
        "delta_sat": float(np.int64(s[1]) / npx),
        "delta_lum": float(np.int64(s[2]) / npx),
    }


class AdaptiveDetector(ContentDetector):
    """scenedetect.detectors.AdaptiveDetector: rolling-window ratio of content_val."""

    ADAPTIVE_RATIO_KEY_TEMPLATE = "adaptive_ratio{luma_only} (w={window_width})"
    _DET_FLAG = capi.ESD_DET_ADAPTIVE

    def __init__(self, adaptive_threshold: float = 3.0, min_scene_len: int = 15, window_width: int = 2,
                 min_content_val: float = 15.0, weights: ContentDetector.Components = ContentDetector.DEFAULT_COMPONENT_WEIGHTS,
                 luma_only: bool = False, kernel_size: Optional[int] = None, video_manager=None,
                 min_delta_hsv: Optional[float] = None):
        if min_delta_hsv is not None:
            min_content_val = min_delta_hsv
        if window_width < 1:
            raise ValueError("window_width must be at least 1.")
        super().__init__(threshold=255.0, min_scene_len=0, weights=weights, luma_only=luma_only, kernel_size=kernel_size)
        self.min_scene_len = min_scene_len
        self.adaptive_threshold = adaptive_threshold
        self.min_content_val = min_content_val
        self.window_width = window_width
        self._adaptive_ratio_key = AdaptiveDetector.ADAPTIVE_RATIO_KEY_TEMPLATE.format(
            window_width=window_width, luma_only="" if not luma_only else "_lum")

    def _own_buffer_length(self) -> int:
        return self.window_width

    def get_metrics(self):
        return super().get_metrics() + [self._adaptive_ratio_key]

    def _fill_config(self, cfg):
        self._fill_edge_config(cfg)
        cfg.detectors |= capi.ESD_DET_ADAPTIVE
        cfg.adaptive_threshold = float(self.adaptive_threshold)
        cfg.adaptive_min_content_val = float(self.min_content_val)
        for i, w in enumerate(self._weights):
            cfg.adaptive_weights[i] = float(w)
        cfg.adaptive_weight_div = self._weight_div()
        cfg.adaptive_window_width = int(self.window_width)
        cfg.adaptive_min_scene_len = int(self.min_scene_len)

    def _metrics_for(self, scores, k):
        npx = float(self._ctx.geometry.dst_width * self._ctx.geometry.dst_height) if self._ctx else 1.0
        return _content_metrics(scores, k, "adaptive_val", npx)

    def _publish_late_metrics(self, first_frame_num, n):
        ContentDetector._publish_late_metrics(self, first_frame_num, n)
        # the ratio of frame t becomes known once frame t + window_width has been processed
        w = self.window_width
        first_video_frame = first_frame_num + n - self._ctx.frames_pushed
        lo = max(first_frame_num - w, first_video_frame)
        cnt = first_frame_num + n - lo
        sc = self._ctx.read_scores(lo, cnt, ["adaptive_ratio"])
        for k in range(cnt):
            r = sc["adaptive_ratio"][k]
            if r == r:
                self.stats_manager.set_metrics(lo + k, {self._adaptive_ratio_key: float(r)})


class HistogramDetector(SceneDetector):
    """scenedetect.detectors.HistogramDetector: correlation of consecutive normalised Y histograms."""

    METRIC_KEYS = ["hist_diff"]
    _DET_FLAG = capi.ESD_DET_HIST

    def __init__(self, threshold: float = 0.05, bins: int = 256, min_scene_len: int = 15):
        super().__init__()
        self._user_threshold = threshold
        # PySceneDetect keeps the correlation-space threshold; the library applies the same clamp
        self._threshold = max(0.0, min(1.0, 1.0 - threshold))
        self._bins = bins
        self._min_scene_len = min_scene_len
        self._metric_keys = [f"hist_diff [bins={self._bins}]"]

    def get_metrics(self):
        return self._metric_keys

    def is_processing_required(self, frame_num: int) -> bool:
        return True

    def _validate_frame(self, frame_img):
        dt = str(frame_img.dtype)
        if not dt.endswith("uint8"):
            raise ValueError("Image must be 8-bit rgb for HistogramDetector")
        if frame_img.shape[-1] != 3:
            raise ValueError("Image must have three color channels for HistogramDetector")

    def _fill_config(self, cfg):
        cfg.detectors |= capi.ESD_DET_HIST
        cfg.hist_threshold = float(self._user_threshold)
        cfg.hist_bins = int(self._bins)
        cfg.hist_min_scene_len = int(self._min_scene_len)

    def _metrics_for(self, scores, k):
        d = scores["hist_diff"][k]
        return {self._metric_keys[0]: float(d)} if d == d else {}


class HashDetector(SceneDetector):
    """scenedetect.detectors.HashDetector: perceptual (DCT) hash of each frame; cut when the Hamming distance to the
    previous frame's hash, divided by size^2, reaches ``threshold`` and ``min_scene_len`` frames have passed
    (SURVEY.md section 8f, row N4).  gray -> INTER_AREA thumbnail -> DCT -> bits run on the device (hash_kernel)."""

    _DET_FLAG = capi.ESD_DET_HASH

    def __init__(self, threshold: float = 0.395, size: int = 16, lowpass: int = 2, min_scene_len: int = 15):
        super().__init__()
        self._threshold = threshold
        self._min_scene_len = min_scene_len
        self._size = size
        self._size_sq = float(size * size)
        self._factor = lowpass
        self._metric_keys = [f"hash_dist [size={self._size} lowpass={self._factor}]"]

    def get_metrics(self) -> List[str]:
        return self._metric_keys

    def is_processing_required(self, frame_num: int) -> bool:
        return True

    def _fill_config(self, cfg):
        cfg.detectors |= capi.ESD_DET_HASH
        cfg.hash_threshold = float(self._threshold)
        cfg.hash_size = int(self._size)
        cfg.hash_lowpass = int(self._factor)
        cfg.hash_min_scene_len = int(self._min_scene_len)

    def _publish_late_metrics(self, first_frame_num: int, n: int):
        _bits, dist = self._ctx.read_hash(first_frame_num, n)
        for k in range(n):
            if dist[k] == dist[k]:
                self.stats_manager.set_metrics(first_frame_num + k, {self._metric_keys[0]: float(dist[k])})


class ThresholdDetector(SceneDetector):
    """scenedetect.detectors.ThresholdDetector: fade out / fade in detection on the average B,G,R value of the
    frame.  Cuts are placed between a fade-out and the next fade-in."""

    class Method(Enum):
        FLOOR = 0    # fade out when the frame average falls below the threshold
        CEILING = 1  # fade out when it rises to / above the threshold

    THRESHOLD_VALUE_KEY = "average_rgb"
    _DET_FLAG = capi.ESD_DET_THRESHOLD

    def __init__(self, threshold: float = 12, min_scene_len: int = 15, fade_bias: float = 0.0,
                 add_final_scene: bool = False, method: "ThresholdDetector.Method" = Method.FLOOR, block_size=None):
        super().__init__()
        self.threshold = int(threshold)
        self.method = ThresholdDetector.Method(method)
        self.fade_bias = fade_bias
        self.min_scene_len = min_scene_len
        self.add_final_scene = add_final_scene
        self._metric_keys = [ThresholdDetector.THRESHOLD_VALUE_KEY]

    def get_metrics(self) -> List[str]:
        return self._metric_keys

    def _fill_config(self, cfg):
        cfg.detectors |= capi.ESD_DET_THRESHOLD
        cfg.thresh_threshold = float(self.threshold)
        cfg.thresh_fade_bias = float(self.fade_bias)
        cfg.thresh_min_scene_len = int(self.min_scene_len)
        cfg.thresh_add_final_scene = 1 if self.add_final_scene else 0
        cfg.thresh_method = int(self.method.value)

    def post_process(self, frame_num: int) -> List[int]:
        cuts = self._flush_staged()
        ctx = self._manager_ctx or self._ctx  # while a SceneManager drives this detector its context is the live one
        if ctx is None:
            return cuts
        return cuts + ctx.post_process(capi.ESD_DET_THRESHOLD, frame_num)

    def _publish_late_metrics(self, first_frame_num: int, n: int):
        avg = self._ctx.read_average_rgb(first_frame_num, n)
        for k in range(n):
            self.stats_manager.set_metrics(first_frame_num + k, {self._metric_keys[0]: float(avg[k])})
