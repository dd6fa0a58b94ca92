"""SceneManager-compatible driver: auto-downscale, detector loop, cut list -> scene list.

Mirrors scenedetect.scene_manager.SceneManager (compute_downscale_factor, detect_scenes,
get_cut_list, get_scene_list, get_scenes_from_cuts).  The per-frame
Python loop of the original (resize -> for det in detectors: det.process_frame) becomes one
fused kernel launch per batch of frames; every registered detector is fed from that pass.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import capi
from .detectors import (AdaptiveDetector, ContentDetector, HashDetector, HistogramDetector, SceneDetector, StatsManager,
                        ThresholdDetector)

DEFAULT_MIN_WIDTH: int = 256


def compute_downscale_factor(frame_width: int, effective_width: int = DEFAULT_MIN_WIDTH, mode: str = "float"):
    """scenedetect.scene_manager.compute_downscale_factor (>= 0.6.2 float; mode="int" for <= 0.6.1)."""
    assert not (frame_width < 1 or effective_width < 1)
    if frame_width < effective_width:
        return 1
    if mode == "int":
        return frame_width // effective_width
    return frame_width / float(effective_width)


def get_scenes_from_cuts(cut_list: Sequence[int], start_pos: int, end_pos: int) -> List[Tuple[int, int]]:
    """scenedetect.scene_manager.get_scenes_from_cuts on frame numbers (end_pos = one past the last frame)."""
    scene_list: List[Tuple[int, int]] = []
    if not cut_list:
        scene_list.append((start_pos, end_pos))
        return scene_list
    last_cut = start_pos
    for cut in cut_list:
        scene_list.append((last_cut, cut))
        last_cut = cut
    scene_list.append((last_cut, end_pos))
    return scene_list


# YUV 4:2:0 source formats: frames are [N, H*3/2, W] arrays (Y plane, then the chroma plane(s))
YUV420_FORMATS = {"nv12": capi.ESD_FMT_NV12, "i420": capi.ESD_FMT_I420}


class TensorVideo:
    """A decoded clip held as one array: numpy uint8 [N,H,W,3] (host) or a CUDA torch tensor.
    Stands where PySceneDetect's VideoStream stands; decode itself is out of scope."""

    def __init__(self, frames, fps: float = 30.0, start_frame: int = 0, pixel_format: str = "bgr24"):
        self.pixel_format = pixel_format
        if pixel_format in YUV420_FORMATS:
            if frames.ndim != 3 or frames.shape[1] % 3:
                raise ValueError("NV12 / I420 frames must be [N, H*3/2, W] uint8 (Y plane, then the interleaved UV plane or the U and V planes)")
        elif pixel_format != "bgr24":
            raise ValueError("pixel_format must be 'bgr24', 'nv12' or 'i420'")
        elif frames.ndim != 4 or frames.shape[3] != 3:
            raise ValueError("frames must be [N,H,W,3] uint8 BGR")
        self.frames = frames
        self.frame_rate = float(fps)
        self.start_frame = int(start_frame)
        self._pos = 0
        # batches are views of one persistent array: the ingest ring may keep reading them after read_batch returns
        self.frames_stable = True

    @property
    def frame_size(self) -> Tuple[int, int]:
        if self.pixel_format in YUV420_FORMATS:
            return int(self.frames.shape[2]), int(self.frames.shape[1]) * 2 // 3
        return int(self.frames.shape[2]), int(self.frames.shape[1])

    @property
    def is_cuda(self) -> bool:
        return not isinstance(self.frames, np.ndarray) and bool(self.frames.is_cuda)

    def read_batch(self, n: int):
        if self._pos >= self.frames.shape[0]:
            return None
        out = self.frames[self._pos:self._pos + n]
        self._pos += out.shape[0]
        return out


class BatchVideo:
    """A clip delivered as an iterable of batches ([k,H,W,3] numpy or CUDA tensors), e.g. a decoder
    or a synthetic generator that cannot hold the whole clip."""

    def __init__(self, batches: Iterable, frame_size: Tuple[int, int], fps: float = 30.0, start_frame: int = 0,
                 pixel_format: str = "bgr24"):
        self.pixel_format = pixel_format
        self._it = iter(batches)
        self.frame_size = (int(frame_size[0]), int(frame_size[1]))
        self.frame_rate = float(fps)
        self.start_frame = int(start_frame)

    def read_batch(self, n: int):
        return next(self._it, None)


class SceneManager:
    def __init__(self, stats_manager: Optional[StatsManager] = None, device: int = 0, batch_frames: int = 512,
                 downscale_mode: str = "float", tuning: Optional[dict] = None, ingest_threads: int = 0):
        self._detector_list: List[SceneDetector] = []
        self.stats_manager = stats_manager
        self._device = device
        self._batch_frames = int(batch_frames)
        self._auto_downscale = True
        self._downscale = 1
        self._downscale_mode = downscale_mode
        self._tuning = dict(tuning or {})
        # host frames: > 0 lets that many host threads gather only the bytes the kernel reads before the H2D copy
        self._ingest_threads = int(ingest_threads)
        self._cutting_list: List[int] = []
        self._cuts_by_detector: dict = {}
        self._start_pos: Optional[int] = None
        self._last_pos: Optional[int] = None
        self._frame_rate = 30.0
        self._ctx: Optional[capi.EsdContext] = None
        self._ctx_key = None
        self.scores: dict = {}

    # ---- PySceneDetect-compatible knobs
    @property
    def auto_downscale(self) -> bool:
        return self._auto_downscale

    @auto_downscale.setter
    def auto_downscale(self, value: bool):
        self._auto_downscale = bool(value)

    @property
    def downscale(self) -> int:
        return self._downscale

    @downscale.setter
    def downscale(self, value: int):
        if value < 1:
            raise ValueError("Downscale factor must be a positive integer >= 1!")
        if self._auto_downscale:
            self._auto_downscale = False
        self._downscale = value

    def add_detector(self, detector: SceneDetector) -> None:
        kinds = [type(d)._DET_FLAG for d in self._detector_list]
        if type(detector)._DET_FLAG in kinds:
            raise ValueError("one detector of each kind (content / adaptive / histogram / threshold / hash) per SceneManager")
        if self.stats_manager is not None:
            detector.stats_manager = self.stats_manager
            self.stats_manager.register_metrics(detector.get_metrics())
        self._detector_list.append(detector)

    def get_num_detectors(self) -> int:
        return len(self._detector_list)

    def clear(self) -> None:
        self._cutting_list = []
        self._cuts_by_detector = {}
        self._start_pos = None
        self._last_pos = None
        self.scores = {}
        self.close()

    def clear_detectors(self) -> None:
        self._detector_list.clear()

    def close(self):
        if self._ctx is not None:
            self._ctx.close()
            self._ctx = None

    # ---- geometry
    def _target_size(self, width: int, height: int) -> Tuple[int, int]:
        factor = compute_downscale_factor(width, mode=self._downscale_mode) if self._auto_downscale else self._downscale
        if factor > 1:
            return max(1, round(width / factor)), max(1, round(height / factor))
        return width, height

    def make_context(self, width: int, height: int, device: Optional[int] = None, pixel_format: str = "bgr24") -> capi.EsdContext:
        """The esd_ctx this manager's detectors need for frames of the given size."""
        if not self._detector_list:
            raise RuntimeError("No detectors registered; call add_detector() first.")
        cfg = capi.default_config()
        cfg.detectors = 0
        for det in self._detector_list:
            det._fill_config(cfg)
        cfg.src_width, cfg.src_height = width, height
        cfg.dst_width, cfg.dst_height = self._target_size(width, height)
        cfg.src_format = YUV420_FORMATS.get(pixel_format, capi.ESD_FMT_BGR24)
        for k, v in self._tuning.items():
            setattr(cfg, k, v)
        return capi.EsdContext(cfg, self._device if device is None else device)

    # ---- detection
    def detect_scenes(self, video=None, frames=None, fps: Optional[float] = None, collect_scores: bool = False,
                      reuse_context: bool = False) -> int:
        """Process every frame of `video` (TensorVideo / BatchVideo) or of `frames` ([N,H,W,3]).
        Returns the number of frames processed.  reuse_context: keep the device context of the previous call when the
        geometry and detectors are unchanged (library batches: a reset instead of a rebuild per video)."""
        if video is None:
            if frames is None:
                raise ValueError("detect_scenes needs video= or frames=")
            video = TensorVideo(frames, fps or 30.0)
        if fps is not None:
            self._frame_rate = float(fps)
        else:
            self._frame_rate = float(getattr(video, "frame_rate", 30.0))
        if self.stats_manager is not None:
            self.stats_manager.fps = self._frame_rate
        width, height = video.frame_size
        pixel_format = getattr(video, "pixel_format", "bgr24")
        nv12 = pixel_format in YUV420_FORMATS   # [N, H*3/2, W] batches: NV12, or planar I420
        key = (width, height, pixel_format, tuple(id(d) for d in self._detector_list), self._auto_downscale, self._downscale)
        if reuse_context and self._ctx is not None and self._ctx_key == key:
            ctx = self._ctx
            ctx.reset()
            self._cuts_by_detector = {}
        else:
            self.close()
            self._ctx = ctx = self.make_context(width, height, pixel_format=pixel_format)
            self._ctx_key = key
        start = int(getattr(video, "start_frame", 0))
        self._start_pos = start
        pos = start
        host_ring_open = False
        while True:
            batch = video.read_batch(self._batch_frames)
            if batch is None or batch.shape[0] == 0:
                break
            n = int(batch.shape[0])
            if isinstance(batch, np.ndarray):
                if not host_ring_open:
                    ctx.ingest_open(3, min(self._batch_frames, 64))
                    if self._ingest_threads > 0 and ctx.dst_size != (width, height):
                        ctx.ingest_set_gather(self._ingest_threads)
                    host_ring_open = True
                if nv12:
                    ctx.ingest_push_nv12_numpy(batch, pos)
                else:
                    ctx.ingest_push_numpy(batch, pos)
                if not getattr(video, "frames_stable", False):
                    ctx.ingest_wait_copied()  # the producer may recycle `batch` as soon as we return to read_batch (H2D done; scoring continues)
            elif nv12:
                ctx.push_nv12_tensor(batch, pos)
            else:
                ctx.push_tensor(batch, pos)
            pos += n
        total = pos - start
        self._last_pos = pos - 1 if total > 0 else None
        if total == 0:
            return 0
        self._cutting_list = []
        for det in self._detector_list:
            cuts, _ = ctx.get_cuts(type(det)._DET_FLAG, 0)
            det._manager_ctx = ctx
            cuts = cuts + det.post_process(pos - 1)
            det._manager_ctx = None
            self._cuts_by_detector[type(det).__name__] = cuts
            self._cutting_list += cuts
        if collect_scores or self.stats_manager is not None:
            self.scores = ctx.read_scores(start, total)
            if any(isinstance(d, ThresholdDetector) for d in self._detector_list):
                self.scores["average_rgb"] = ctx.read_average_rgb(start, total)
            if any(isinstance(d, HashDetector) for d in self._detector_list):
                self.scores["hash_bits"], self.scores["hash_dist"] = ctx.read_hash(start, total)
            if any(isinstance(d, ContentDetector) and d._weights.delta_edges > 0.0 for d in self._detector_list):
                self.scores["edge_counts"] = ctx.read_edge_counts(start, total)
            if self.stats_manager is not None:
                self._publish_stats(start, total)
        if host_ring_open:
            ctx.ingest_close()
        return total

    def _publish_stats(self, start: int, total: int):
        npx = float(self._ctx.geometry.dst_width * self._ctx.geometry.dst_height)
        sc = self.scores
        for det in self._detector_list:
            for k in range(total):
                fn = start + k
                if isinstance(det, AdaptiveDetector):
                    if k > 0:
                        m = {"content_val": float(sc["adaptive_val"][k])}
                        m.update(_deltas(sc, k, npx))
                        self.stats_manager.set_metrics(fn, m)
                    r = sc["adaptive_ratio"][k]
                    if r == r:
                        self.stats_manager.set_metrics(fn, {det._adaptive_ratio_key: float(r)})
                elif isinstance(det, ContentDetector):
                    if k > 0:
                        m = {"content_val": float(sc["content_val"][k])}
                        m.update(_deltas(sc, k, npx))
                        self.stats_manager.set_metrics(fn, m)
                elif isinstance(det, HistogramDetector):
                    d = sc["hist_diff"][k]
                    if d == d:
                        self.stats_manager.set_metrics(fn, {det.get_metrics()[0]: float(d)})
            if isinstance(det, HashDetector):
                for k in range(total):
                    d = sc["hash_dist"][k]
                    if d == d:
                        self.stats_manager.set_metrics(start + k, {det.get_metrics()[0]: float(d)})
            if isinstance(det, ThresholdDetector):
                avg = self._ctx.read_average_rgb(start, total)
                for k in range(total):
                    self.stats_manager.set_metrics(start + k, {det.get_metrics()[0]: float(avg[k])})

    def get_cut_list(self) -> List[int]:
        return sorted(set(self._cutting_list))

    def cuts_of(self, detector: SceneDetector) -> List[int]:
        return list(self._cuts_by_detector.get(type(detector).__name__, []))

    def get_scene_list(self, start_in_scene: bool = False) -> List[Tuple[int, int]]:
        if self._last_pos is None:
            return []
        cut_list = self.get_cut_list()
        scene_list = get_scenes_from_cuts(cut_list, self._start_pos, self._last_pos + 1)
        if not cut_list and not start_in_scene:
            scene_list = []
        return scene_list

    @property
    def frame_rate(self) -> float:
        return self._frame_rate


def _deltas(sc, k, npx):
    s = sc["sums3"][k]
    out = {"delta_hue": float(np.int64(s[0]) / npx), "delta_sat": float(np.int64(s[1]) / npx),
           "delta_lum": float(np.int64(s[2]) / npx)}
    if "edge_counts" in sc:
        out["delta_edges"] = float(np.int64(255 * int(sc["edge_counts"][k])) / npx)
    return out
