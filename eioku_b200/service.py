"""ml-service scene-task surface: ``detect_scenes(video, config) -> {"scenes": [...]}``.

Drop-in for ``ModelManager.detect_scenes`` of the reference
(/root/reference/ml-service/src/services/model_manager.py:715-835): same coroutine signature,
same result keys (``scene_index, start_ms, end_ms, duration_ms`` -- :775-781, :809-824), same
"raise on failure" behaviour (the task handler marks the task failed,
ml-service/src/workers/task_handler.py:452-469).  Every scene carries start_ms and end_ms with
start <= end as the handler requires (task_handler.py:277-308) and satisfies SceneV1
(backend/src/domain/schemas/scene_v1.py:13-16: all >= 0, duration_ms > 0).

Behaviour changes versus the ffmpeg implementation, on purpose:
the first scene [0, first_cut) is emitted, scene_index is dense, and ``threshold`` is a
PySceneDetect threshold only when ``detector`` is named -- the legacy 0-1 ffmpeg value and
``min_scene_length`` (seconds) of existing configs are ignored, not reinterpreted.
"""
from __future__ import annotations

import logging
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from .detectors import AdaptiveDetector, ContentDetector, FlashFilter, HashDetector, HistogramDetector, ThresholdDetector
from .scene_manager import BatchVideo, SceneManager, TensorVideo

logger = logging.getLogger(__name__)

PRODUCER = "scenedetect"  # ml-service/src/models/responses.py:141
PRODUCER_VERSION = "0.6.4-b200"


def scenes_to_dicts(scenes: Sequence[Tuple[int, int]], fps: float) -> List[dict]:
    """[(start_frame, end_frame)] -> ml-service scene dicts (int() truncation as model_manager.py:771)."""
    out = []
    for a, b in scenes:
        start_ms = int(a / fps * 1000)
        end_ms = int(b / fps * 1000)
        if end_ms - start_ms <= 0:
            continue  # SceneV1 requires duration_ms > 0 (only possible for fps > 1000)
        out.append({"scene_index": len(out), "start_ms": start_ms, "end_ms": end_ms, "duration_ms": end_ms - start_ms})
    return out


_METHOD_OF = {"ContentDetector": ("content", "content_val"), "AdaptiveDetector": ("adaptive", "adaptive_ratio"),
              "HistogramDetector": ("hist", "hist_diff"), "ThresholdDetector": ("threshold", "average_rgb"),
              "HashDetector": ("hash", "hash_dist")}


def scene_artifact_payloads(sm: SceneManager) -> List[dict]:
    """Scene artifact payloads of the reference's artifact-envelope design -- ``SceneV1 {scene_index, method, score,
    frame_number}`` (/root/reference/.kiro/specs/artifact-envelope-architecture/design.md:159-167, requirements.md:173-183;
    specified, never implemented there).  One record per scene: the first scene starts at the first frame (method
    "start", score 0.0); every other scene starts at a cut and carries the detector that emitted it and that detector's
    metric at the cut frame.  Needs ``detect_scenes(..., collect_scores=True)``."""
    if sm._start_pos is None or sm._last_pos is None:
        return []
    start = sm._start_pos
    by_frame = {}
    for det in sm._detector_list:
        method, key = _METHOD_OF[type(det).__name__]
        vals = sm.scores.get(key)
        for cut in sm.cuts_of(det):
            score = float(vals[cut - start]) if vals is not None and 0 <= cut - start < len(vals) else float("nan")
            by_frame.setdefault(cut, (method, score))  # first registered detector wins when two agree on a frame
    out = [{"scene_index": 0, "method": "start", "score": 0.0, "frame_number": start}]
    for cut in sorted(by_frame):
        method, score = by_frame[cut]
        out.append({"scene_index": len(out), "method": method, "score": score, "frame_number": int(cut)})
    return out


def frame_to_timecode(frame: int, fps: float, precision: int = 3) -> str:
    """HH:MM:SS.nnn of a frame number, formatted like scenedetect.FrameTimecode.get_timecode (rounded to
    `precision` decimals, with the 60-second carry guard) [upstream-recall]."""
    secs = frame / fps
    hrs = int(secs / 3600.0)
    secs -= hrs * 3600.0
    mins = int(secs / 60.0)
    secs = max(0.0, secs - mins * 60.0)
    secs = min(60.0, round(secs, precision))
    if int(secs) == 60:
        secs = 0.0
        mins += 1
        if mins >= 60:
            mins = 0
            hrs += 1
    msec = format(secs, ".%df" % (precision + 1)) if precision else ""
    msec_str = msec[-(2 + precision):-1]
    return "%02d:%02d:%02d%s" % (hrs, mins, int(secs), msec_str)


def scenes_to_boundaries(scenes: Sequence[Tuple[int, int]], fps: float) -> List[dict]:
    """SceneBoundary records of the reference's design document
    (/root/reference/.kiro/specs/semantic-video-search/design.md:994-1007): {scene, start, end} with
    HH:MM:SS.mmm timecodes."""
    return [{"scene": i, "start": frame_to_timecode(a, fps), "end": frame_to_timecode(b, fps)}
            for i, (a, b) in enumerate(scenes)]


def build_detectors(config: dict) -> list:
    """Detector objects for a scene-task config dict (surface B1)."""
    cfg = dict(config or {})
    names = cfg.get("detector")
    if names is None:
        if "threshold" in cfg or "min_scene_length" in cfg:
            logger.info("scene config has no 'detector': legacy ffmpeg keys threshold=%r min_scene_length=%r ignored; "
                        "running ContentDetector(27.0, 15)", cfg.get("threshold"), cfg.get("min_scene_length"))
        return [ContentDetector()]
    if isinstance(names, str):
        names = [n.strip() for n in names.split("+")]
    weights = cfg.get("weights")
    if weights is not None:
        weights = ContentDetector.Components(*weights)
    common = {}
    if "min_scene_len" in cfg:
        common["min_scene_len"] = int(cfg["min_scene_len"])
    dets = []
    for name in names:
        if name in ("content", "detect-content"):
            kw = dict(common)
            if "threshold" in cfg:
                kw["threshold"] = float(cfg["threshold"])
            if weights is not None:
                kw["weights"] = weights
            if "luma_only" in cfg:
                kw["luma_only"] = bool(cfg["luma_only"])
            if "filter_mode" in cfg:
                fm = cfg["filter_mode"]
                kw["filter_mode"] = FlashFilter.Mode[fm.upper()] if isinstance(fm, str) else FlashFilter.Mode(fm)
            dets.append(ContentDetector(**kw))
        elif name in ("adaptive", "detect-adaptive"):
            kw = dict(common)
            for k in ("adaptive_threshold", "min_content_val"):
                if k in cfg:
                    kw[k] = float(cfg[k])
            if "window_width" in cfg:
                kw["window_width"] = int(cfg["window_width"])
            if weights is not None:
                kw["weights"] = weights
            if "luma_only" in cfg:
                kw["luma_only"] = bool(cfg["luma_only"])
            dets.append(AdaptiveDetector(**kw))
        elif name in ("hist", "histogram", "detect-hist"):
            kw = dict(common)
            if "hist_threshold" in cfg:
                kw["threshold"] = float(cfg["hist_threshold"])
            elif "threshold" in cfg and len(names) == 1:
                kw["threshold"] = float(cfg["threshold"])
            if "bins" in cfg:
                kw["bins"] = int(cfg["bins"])
            dets.append(HistogramDetector(**kw))
        elif name in ("threshold", "detect-threshold"):
            kw = dict(common)
            if "fade_threshold" in cfg:
                kw["threshold"] = float(cfg["fade_threshold"])
            elif "threshold" in cfg and len(names) == 1:
                kw["threshold"] = float(cfg["threshold"])
            for k in ("fade_bias",):
                if k in cfg:
                    kw[k] = float(cfg[k])
            if "add_final_scene" in cfg:
                kw["add_final_scene"] = bool(cfg["add_final_scene"])
            dets.append(ThresholdDetector(**kw))
        elif name in ("hash", "detect-hash"):
            kw = dict(common)
            if "hash_threshold" in cfg:
                kw["threshold"] = float(cfg["hash_threshold"])
            elif "threshold" in cfg and len(names) == 1:
                kw["threshold"] = float(cfg["threshold"])
            for k in ("size", "lowpass"):
                if k in cfg:
                    kw[k] = int(cfg[k])
            dets.append(HashDetector(**kw))
        else:
            raise ValueError(f"unknown scene detector {name!r} (content | adaptive | hist | threshold | hash)")
    return dets


def detect_scenes_frames(video, config: Optional[dict] = None, fps: Optional[float] = None, device: int = 0,
                         batch_frames: int = 512, devices: Optional[Sequence[int]] = None) -> dict:
    """Synchronous core: `video` is a TensorVideo/BatchVideo or an [N,H,W,3] uint8 array/tensor.

    devices: more than one GPU id fans ONE video out over the box by frame ranges with the window_width + 1 halo and a
    single global decision pass (eioku_b200.multi.detect_sharded); `video` must then be a host array (or a TensorVideo
    over one) so every device can pull its own range, or a list of per-device CUDA tensors (multi.plan_shards)."""
    config = config or {}
    if devices is not None and len(devices) > 1:
        return _detect_scenes_sharded(video, config, fps, list(devices), batch_frames)
    if devices:
        device = int(devices[0])
    if not hasattr(video, "read_batch"):
        video = TensorVideo(video, fps or float(config.get("fps", 30.0)))
    sm = SceneManager(device=device, batch_frames=batch_frames,
                      downscale_mode=str(config.get("downscale_mode", "float")),
                      ingest_threads=int(config.get("ingest_threads", 0)))
    if "downscale" in config:
        sm.downscale = int(config["downscale"])
    if config.get("auto_downscale") is False:
        sm.auto_downscale = False
    for det in build_detectors(config):
        sm.add_detector(det)
    want_artifacts = bool(config.get("artifact_payloads"))
    try:
        n = sm.detect_scenes(video, collect_scores=want_artifacts)
        rate = fps or video.frame_rate
        if n == 0:
            return {"scenes": [], "artifact_payloads": []} if want_artifacts else {"scenes": []}
        # the reference always reports at least one scene for a readable video (model_manager.py:816-828)
        scenes = sm.get_scene_list(start_in_scene=True)
        out = {"scenes": scenes_to_dicts(scenes, rate)}
        if want_artifacts:  # opt-in: the artifact-envelope SceneV1 payloads next to the task schema
            out["artifact_payloads"] = scene_artifact_payloads(sm)
        return out
    finally:
        sm.close()


def _detect_scenes_sharded(video, config: dict, fps: Optional[float], devices: List[int], batch_frames: int) -> dict:
    from . import multi

    pixel_format, start = "bgr24", 0
    if isinstance(video, TensorVideo):
        fps = fps or video.frame_rate
        pixel_format, start = video.pixel_format, video.start_frame
        video = video.frames
    if hasattr(video, "read_batch"):
        raise ValueError("frame-range sharding needs random access to the frames: pass an array, a TensorVideo over a host "
                         "array, or one CUDA tensor per device")
    if not isinstance(video, (list, tuple)) and not isinstance(video, np.ndarray):
        raise ValueError("frame-range sharding over several devices takes host frames (numpy) or per-device CUDA tensors")
    rate = fps or float(config.get("fps", 30.0))
    if "downscale" in config or config.get("auto_downscale") is False:
        raise ValueError("devices=[...] supports the default auto-downscale only")
    res = multi.detect_sharded(video, build_detectors(config), devices, fps=rate, batch_frames=max(batch_frames, 256),
                               downscale_mode=str(config.get("downscale_mode", "float")),
                               ingest_threads=int(config.get("ingest_threads", 0)), pixel_format=pixel_format, start_frame=start)
    if res.n_frames == 0:
        return {"scenes": []}
    return {"scenes": scenes_to_dicts(res.scene_list(start_in_scene=True), rate)}


def _detect_scenes_mjpeg_sharded(video_path: str, probe, config: dict, devices: List[int]) -> dict:
    """One Motion-JPEG file over several GPUs: MJPEG is intra-only, so every device decodes its own frame range (+ halo)."""
    from . import decode, multi

    n, size, rate = probe.n_frames, probe.frame_size, probe.frame_rate
    probe.close()
    batch = int(config.get("decode_batch", 64))
    res = multi.detect_sharded(
        lambda sh, dev: decode.MjpegVideo(video_path, device=dev, batch_frames=batch, first_frame=sh.load_start, end_frame=sh.load_end),
        build_detectors(config), devices, fps=rate, batch_frames=batch, downscale_mode=str(config.get("downscale_mode", "float")),
        n_frames=n, frame_size=size)
    if res.n_frames == 0:
        return {"scenes": []}
    return {"scenes": scenes_to_dicts(res.scene_list(start_in_scene=True), rate)}


def default_decode_workers(n_frames: int) -> int:
    """How many captures decode one file at once when the caller does not say (`decode_workers`): half the cores this process
    may use, at most 8, and none for clips too short to pay for the seeks."""
    import os

    if n_frames < 2048:
        return 1
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        cores = os.cpu_count() or 2
    return max(1, min(8, cores // 2))   # 16 cores: 1 / 2 / 4 / 8 / 12 / 16 captures -> 606 / 770 / 1 014 / 1 149 / 1 089 / 532 frames/s


def _detect_scenes_capture_sharded(video_path: str, config: dict, device: int, workers: int) -> Optional[dict]:
    """One file, `workers` host captures decoding frame ranges (+ halo) at once, all scored on `device`, ONE decision pass.
    Returns None when a range could not be trusted (an inexact seek, a short read, a wrong frame count): the caller then decodes
    the file sequentially -- the results are never built on a capture that may have landed on the wrong frame."""
    from . import decode, multi
    from .sharding import frame_range_shards

    with decode.CaptureRangeVideo(video_path, 0, None, 1) as probe:
        n, size, rate = probe.n_frames, probe.frame_size, probe.frame_rate
    if n <= 0:
        return None
    dets = build_detectors(config)
    batch = int(config.get("decode_batch", 64))
    shards = frame_range_shards(n, workers, multi.halo_width(dets))
    starts = [sh.load_start for sh in shards if sh.own_end > sh.own_start and sh.load_start > 0]
    try:
        import os

        cores = len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        cores = 4
    threads = max(1, cores // workers)   # FFmpeg frame threads per capture: together they fill the cores once, not `workers` times
    sources = {}

    def make(sh, dev):
        src = decode.CaptureRangeVideo(video_path, sh.load_start, sh.load_end, batch, watch=starts, threads=threads, frame_count=n)
        sources[sh.rank] = src
        return src

    try:
        res = multi.detect_sharded(make, dets, [device] * workers, fps=rate, batch_frames=batch,
                                   downscale_mode=str(config.get("downscale_mode", "float")), n_frames=n, frame_size=size)
    except (decode.SeekError, RuntimeError) as e:
        logger.warning("segment-parallel decode of %s abandoned (%s); decoding sequentially", video_path, e)
        return None
    # every range's first frame must be the frame its owner decoded at that index
    for sh in shards:
        if sh.rank not in sources or sh.load_start == 0:
            continue
        mine = sources[sh.rank].digests.get(sh.load_start)
        owner = next((s for s in shards if s.own_start <= sh.load_start < s.own_end and s.rank in sources), None)
        theirs = sources[owner.rank].digests.get(sh.load_start) if owner is not None else None
        if mine is None or theirs is None or mine != theirs:
            logger.warning("segment-parallel decode of %s: frame %d differs between two captures (inexact seek); decoding sequentially",
                           video_path, sh.load_start)
            return None
    # ... and the file must end where the container said it does: a capture that still had frames behind the claimed count means
    # the sequential decode (which reads until the decoder stops) would see a longer video
    if any(src.trailing_frames for src in sources.values()):
        logger.warning("segment-parallel decode of %s: the container under-reports its length; decoding sequentially", video_path)
        return None
    if res.n_frames == 0:
        return {"scenes": []}
    return {"scenes": scenes_to_dicts(res.scene_list(start_in_scene=True), rate)}


def detect(video, detector, stats_file_path: Optional[str] = None, start_in_scene: bool = False, fps: Optional[float] = None,
           device: int = 0) -> List[Tuple[int, int]]:
    """Counterpart of ``scenedetect.detect(video_path, detector, ...)``: run one detector over a video and return the
    scene list as (start_frame, end_frame) pairs.  `video` is a path (decoded on the host, see `_default_decoder`),
    a TensorVideo / BatchVideo, or an [N,H,W,3] uint8 array / CUDA tensor."""
    if isinstance(video, str):
        video = _default_decoder(video, {"fps": fps or 30.0})
    elif not hasattr(video, "read_batch"):
        video = TensorVideo(video, fps or 30.0)
    from .detectors import StatsManager

    stats = StatsManager() if stats_file_path else None
    sm = SceneManager(stats_manager=stats, device=device)
    sm.add_detector(detector)
    try:
        sm.detect_scenes(video)
        if stats is not None:
            stats.save_to_csv(stats_file_path)
        return sm.get_scene_list(start_in_scene=start_in_scene)
    finally:
        sm.close()


def _wants_capture_workers(video_path: str, config: dict) -> bool:
    """True for files that take the cv2.VideoCapture route (not .npy dumps, not Motion-JPEG AVI with GPU decode on) unless the
    caller asked for one worker or for options the sharded path does not carry."""
    if video_path.endswith(".npy") or int(config.get("decode_workers", 0)) == 1:
        return False
    if "downscale" in config or config.get("auto_downscale") is False or config.get("artifact_payloads"):
        return False
    if config.get("gpu_decode", True):
        from . import decode

        if decode.is_mjpeg_avi(video_path):
            return False
    try:
        build = build_detectors(config)
        from . import multi

        multi._check_shardable(build)
    except Exception:
        return False
    return True


def _default_decoder(video_path: str, config: dict):
    """Frames for a path: .npy frame dumps; Motion-JPEG AVI files decoded ON THE GPU (eioku_b200.decode, SURVEY.md 8f N1 --
    the decoded frames never visit host memory); any other container through cv2.VideoCapture in host batches, the
    reference's own decode loop (model_manager.py:237-263), feeding the ingest ring."""
    if video_path.endswith(".npy"):
        arr = np.load(video_path, mmap_mode="r")
        return TensorVideo(arr, float(config.get("fps", 30.0)))
    if config.get("gpu_decode", True):
        from . import decode

        if decode.is_mjpeg_avi(video_path):
            return decode.MjpegVideo(video_path, device=int(config.get("device", 0)), batch_frames=int(config.get("decode_batch", 64)))
    try:
        import cv2  # noqa: F401, WPS433 (decode helper only; never used for scoring)
    except Exception as e:  # pragma: no cover
        raise RuntimeError("no decoder available for " + video_path) from e
    from . import decode

    # frames are decoded straight into reused batch buffers by a thread of the source's own (no per-batch np.stack, decode
    # overlapped with the push): 295 -> ~680 frames/s on a 1080p MPEG-4 file, the rate of cv2's decode itself
    src = decode.CaptureRangeVideo(video_path, 0, None, batch_frames=int(config.get("decode_batch", 64)), until_eof=True)
    if "fps" in config and not src.frame_rate:
        src.frame_rate = float(config["fps"])
    return src


def provenance_hashes(video_path: str, config: dict) -> Tuple[str, str]:
    """(config_hash, input_hash) with the semantics of the reference's provenance helpers
    (/root/reference/ml-service/src/utils/hashing.py:12-54): xxh64 of the key-sorted JSON config and xxh64 of the video
    file's bytes (of the path string when the file does not exist), first 16 hex digits.  The task handler leaves both
    empty today (ml-service/src/workers/task_handler.py:149-150); SURVEY.md 8f N2 asks for real ones."""
    import json
    import os

    try:
        import xxhash

        def new():
            return xxhash.xxh64()
    except Exception:  # pragma: no cover - xxhash is a reference dependency; keep the task alive without it
        import hashlib

        def new():
            return hashlib.blake2b(digest_size=8)
    h = new()
    h.update(json.dumps(config or {}, sort_keys=True).encode())
    config_hash = h.hexdigest()[:16]
    h = new()
    if os.path.exists(video_path):
        with open(video_path, "rb") as f:
            for block in iter(lambda: f.read(1 << 20), b""):
                h.update(block)
    else:
        h.update(video_path.encode())
    return config_hash, h.hexdigest()[:16]


def scene_detection_response(video_path: str, config: dict, scenes: Sequence[dict], run_id: Optional[str] = None) -> dict:
    """Fields of the reference's SceneDetectionResponse (ml-service/src/models/responses.py:135-143) around a
    detect_scenes result."""
    import uuid

    config_hash, input_hash = provenance_hashes(video_path, config)
    return {"run_id": run_id or str(uuid.uuid4()), "config_hash": config_hash, "input_hash": input_hash,
            "producer": PRODUCER, "producer_version": PRODUCER_VERSION,
            "scenes": [{"scene_index": s["scene_index"], "start_ms": s["start_ms"], "end_ms": s["end_ms"]} for s in scenes]}


def scene_artifact_envelopes(result: dict, video_id: str, run_id: Optional[str] = None, created_at=None,
                             task_type: str = "scene_detection") -> List[dict]:
    """The ArtifactEnvelope fields the reference's task handler builds per scene from a scene-task result
    (/root/reference/ml-service/src/workers/task_handler.py:145-153 provenance defaults, :257-331 per-item loop;
    dataclass at ml-service/src/domain/artifacts.py:7-73), as a pure function: one dict per scene, keyed like the
    dataclass, so ``ArtifactEnvelope(**d)`` validates.  `result` is what ``detect_scenes`` returns, optionally with the
    provenance keys of ``scene_detection_response`` merged in; like the handler, a missing producer falls back to
    "ml-service"/"1.0.0", missing hashes to "", and items without start_ms/end_ms or with start > end are dropped."""
    import json
    import uuid
    from datetime import datetime, timezone

    run_id = run_id or result.get("run_id") or str(uuid.uuid4())
    created_at = created_at or datetime.now(timezone.utc).replace(tzinfo=None)
    out = []
    for idx, scene in enumerate(result.get("scenes", [])):
        if "start_ms" not in scene or "end_ms" not in scene:
            continue  # task_handler.py:277-293: no time information -> dropped
        a, b = int(scene["start_ms"]), int(scene["end_ms"])
        if a < 0 or b < 0 or a > b:
            continue  # task_handler.py:295-308
        out.append({
            "artifact_id": f"{video_id}_{task_type}_{run_id}_{idx}",
            "asset_id": video_id,
            "artifact_type": "scene",  # task_to_artifact_type, task_handler.py:160-168
            "schema_version": 1,
            "span_start_ms": a,
            "span_end_ms": b,
            "payload_json": json.dumps(scene),
            "producer": result.get("producer", "ml-service"),
            "producer_version": result.get("producer_version", "1.0.0"),
            "model_profile": result.get("model_profile", "balanced"),
            "config_hash": result.get("config_hash", ""),
            "input_hash": result.get("input_hash", ""),
            "run_id": run_id,
            "created_at": created_at,
        })
    return out


class ModelManager:
    """The scene-detection slice of the reference's ModelManager (model_manager.py:715)."""

    def __init__(self, decoder: Optional[Callable[[str, dict], object]] = None, device: int = 0,
                 devices: Optional[Sequence[int]] = None):
        """devices: GPUs one job may fan out over (frame-range sharding; the reference's worker runs one job at a time,
        ml-service/src/main_worker.py:124, so a job is the unit that has to use the whole box)."""
        self._decoder = decoder or _default_decoder
        self._device = device
        self._devices = list(devices) if devices else None

    async def detect_scenes(self, video_path: str, config: dict) -> dict:
        try:
            logger.info("Scene detection: %s", video_path)
            dev0 = self._devices[0] if self._devices else self._device
            if self._decoder is _default_decoder and _wants_capture_workers(video_path, config or {}):
                # a codec only the host can decode: several captures decode frame ranges of the file at once
                workers = int((config or {}).get("decode_workers", 0))
                if workers <= 0:
                    from . import decode as _decode

                    with _decode.CaptureRangeVideo(video_path, 0, None, 1) as probe:
                        workers = default_decode_workers(probe.n_frames)
                if workers > 1:
                    result = _detect_scenes_capture_sharded(video_path, config or {}, dev0, workers)
                    if result is not None:
                        logger.info("Scene detection complete: %d scenes (%d decode workers)", len(result["scenes"]), workers)
                        return result
            video = self._decoder(video_path, {**(config or {}), "device": dev0})
            if self._devices and len(self._devices) > 1 and type(video).__name__ == "MjpegVideo":
                result = _detect_scenes_mjpeg_sharded(video_path, video, config or {}, self._devices)
            elif self._devices and len(self._devices) > 1 and isinstance(getattr(video, "frames", None), np.ndarray):
                result = detect_scenes_frames(video, config, devices=self._devices)
            else:
                result = detect_scenes_frames(video, config, device=self._devices[0] if self._devices else self._device)
            logger.info("Scene detection complete: %d scenes", len(result["scenes"]))
            return result
        except Exception as e:
            logger.error("Scene detection failed: %s", e, exc_info=True)
            raise
