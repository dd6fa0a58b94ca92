"""One scene-detection job fanned out over the GPUs of a box.

The reference runs one job at a time per worker (/root/reference/ml-service/src/main_worker.py:124, WORKER_MAX_JOBS=1)
and its scene task is one call, ``ModelManager.detect_scenes`` (ml-service/src/services/model_manager.py:715-723); so one
call must be able to use every GPU of the node.  Two partitionings, both without a data-path collective:

* **frame ranges of one long video** (`detect_sharded`): device g scores frames [s_g - w - 1, e_g + w) (halo: 1 frame for the
  previous-frame delta + the adaptive window) through its own ``esd_ctx`` and stream; the owned float64 score slices are
  copied device-to-device (NVLink peer copy, ``esd_copy_scores_device``) into one array on the deciding GPU and ONE global
  decision pass (``esd_decide_device``) walks the min-scene-len / flash-filter state machines, which are sequential and
  therefore cannot be merged from per-shard cut lists.  Scores never visit the host; no pickle, no NCCL.
* **whole videos of a library** (`detect_library`): longest-processing-time greedy over the devices; every device runs the
  complete chain (scoring + decision) for its videos.

The multi-process flavour (one rank per GPU under torchrun, as bench.py runs it) shares the same pieces:
``all_gather_scores`` moves the owned slices as plain float64 tensors (NCCL on device tensors, gloo on CPU tensors in the
tests) and rank 0 runs the same single decision pass.
"""
from __future__ import annotations

import threading
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import capi
from .detectors import AdaptiveDetector, SceneDetector, ThresholdDetector
from .scene_manager import SceneManager, TensorVideo, get_scenes_from_cuts
from .sharding import FrameShard, all_gather_scores, frame_range_shards, partition_videos

_SCORE_NAME = {capi.ESD_DET_CONTENT: "content_val", capi.ESD_DET_ADAPTIVE: "adaptive_val", capi.ESD_DET_HIST: "hist_diff",
               capi.ESD_DET_THRESHOLD: "average_rgb", capi.ESD_DET_HASH: "hash_dist"}


@dataclass
class ShardedResult:
    n_frames: int
    start_frame: int
    cuts_by_detector: Dict[str, List[int]]
    scores: Dict[str, np.ndarray] = field(default_factory=dict)
    shards: List[FrameShard] = field(default_factory=list)

    @property
    def cut_list(self) -> List[int]:
        out = set()
        for c in self.cuts_by_detector.values():
            out.update(c)
        return sorted(out)

    def scene_list(self, start_in_scene: bool = True) -> List[Tuple[int, int]]:
        if self.n_frames == 0:
            return []
        cuts = self.cut_list
        if not cuts and not start_in_scene:
            return []
        return get_scenes_from_cuts(cuts, self.start_frame, self.start_frame + self.n_frames)


def halo_width(detectors: Sequence[SceneDetector]) -> int:
    """window_width of the widest AdaptiveDetector (0 without one): shards load w + 1 frames before and w after."""
    return max([d.window_width for d in detectors if isinstance(d, AdaptiveDetector)] + [0])


def _check_shardable(detectors: Sequence[SceneDetector]):
    for d in detectors:
        if isinstance(d, ThresholdDetector) and d.add_final_scene:
            raise ValueError("ThresholdDetector(add_final_scene=True) keeps its fade state in the scoring context; "
                             "run it on one device (its post_process is not part of the global decision pass)")


def _manager(detectors, device, batch_frames, downscale_mode, tuning, ingest_threads=0) -> SceneManager:
    sm = SceneManager(device=device, batch_frames=batch_frames, downscale_mode=downscale_mode, tuning=tuning,
                      ingest_threads=ingest_threads)
    for d in detectors:
        sm.add_detector(d)
    return sm


class ShardScorer:
    """One device's half of a frame-range job: an esd_ctx, a stream, and the peer copy of its owned score slices."""

    def __init__(self, detectors: Sequence[SceneDetector], width: int, height: int, device: int, shard: FrameShard,
                 batch_frames: int = 1024, downscale_mode: str = "float", tuning: Optional[dict] = None,
                 pixel_format: str = "bgr24", ingest_threads: int = 0):
        import torch

        self.device = int(device)
        self.shard = shard
        self.flags = [type(d)._DET_FLAG for d in detectors]
        tuning = dict(tuning or {})
        tuning.setdefault("initial_capacity", shard.load_end - shard.load_start + 16)
        self._sm = _manager(detectors, self.device, batch_frames, downscale_mode, tuning)
        self.ctx = self._sm.make_context(width, height, device=self.device, pixel_format=pixel_format)
        self._batch = int(batch_frames)
        self._ingest_threads = int(ingest_threads)
        self._ring_open = False
        self._nv12 = pixel_format in ("nv12", "i420")
        with torch.cuda.device(self.device):
            self.stream = torch.cuda.Stream(device=self.device)
            self.done = torch.cuda.Event()
        self.pos = shard.load_start

    # -- feeding: frames of [load_start, load_end) in order
    def push_device(self, frames):
        """CUDA tensor holding the next frames of this shard's load range (on this device)."""
        n = int(frames.shape[0])
        if self._nv12:
            self.ctx.push_nv12_tensor(frames, self.pos, self.stream.cuda_stream)
        else:
            self.ctx.push_tensor(frames, self.pos, self.stream.cuda_stream)
        self.pos += n

    def push_host(self, frames: np.ndarray):
        """Host frames (numpy; pinned or pageable) through this device's ingest ring, in batches."""
        if not self._ring_open:
            self.ctx.ingest_open(3, min(self._batch, 128))
            if self._ingest_threads > 0 and self.ctx.dst_size != (self.ctx.cfg.src_width, self.ctx.cfg.src_height):
                self.ctx.ingest_set_gather(self._ingest_threads)
            self._ring_open = True
        for a in range(0, frames.shape[0], self._batch):
            part = frames[a:a + self._batch]
            if self._nv12:
                self.ctx.ingest_push_nv12_numpy(part, self.pos)
            else:
                self.ctx.ingest_push_numpy(part, self.pos)
            self.pos += int(part.shape[0])

    def copy_owned(self, merged, dst_device: int):
        """Enqueue the device-to-device copies of the owned score slices into `merged` ([kinds, N] float64 on
        `dst_device`, row k = self.flags[k]) and record `self.done` behind them."""
        sh = self.shard
        n_own = sh.own_end - sh.own_start
        if self.pos != sh.load_end:
            raise RuntimeError(f"shard {sh.rank}: fed frames up to {self.pos}, expected {sh.load_end}")
        # ingest pushes run on the library's compute stream; device pushes on ours -- copy_scores_device orders the copy
        # behind the scoring tails, and an ingest ring is drained first
        if self._ring_open:
            self.ctx.synchronize()
        for k, flag in enumerate(self.flags):
            if n_own > 0:
                dst = merged[k, sh.own_start:].data_ptr()
                self.ctx.copy_scores_device(capi.DECISION_SCORE_KIND[flag], sh.own_start, n_own, dst, dst_device, self.stream.cuda_stream)
        self.done.record(self.stream)

    def close(self):
        if self._ring_open:
            self.ctx.ingest_close()
            self._ring_open = False
        self.ctx.close()


def decide_merged(ctx: capi.EsdContext, flags: Sequence[int], merged, n_frames: int, first_frame: int, stream: int,
                  names: Optional[Sequence[str]] = None) -> Dict[str, List[int]]:
    """ONE global decision pass per detector over the merged device-resident score rows (row k = flags[k])."""
    out = {}
    for k, flag in enumerate(flags):
        cuts = ctx.decide_device(flag, first_frame, merged[k].data_ptr(), n_frames, stream)
        out[names[k] if names else str(flag)] = cuts
    return out


def detect_sharded(frames, detectors: Sequence[SceneDetector], devices: Sequence[int], fps: float = 30.0,
                   batch_frames: int = 1024, downscale_mode: str = "float", tuning: Optional[dict] = None,
                   collect_scores: bool = False, ingest_threads: int = 0, pixel_format: str = "bgr24",
                   start_frame: int = 0, n_frames: Optional[int] = None, frame_size: Optional[Tuple[int, int]] = None) -> ShardedResult:
    """Frame-range sharding of ONE video over `devices` (one process, one esd_ctx + stream per device).

    `frames`: a host array [N,H,W,3] uint8 (numpy, pinned or pageable -- each device's ingest ring pulls its own range),
    or a list with one entry per device holding that device's load range [load_start, load_end) as a CUDA tensor on that
    device (frames born on the devices; use ``plan_shards`` to get the ranges), or a callable ``(shard, device) -> source``
    returning a video source (``read_batch``) for the shard's load range on that device -- e.g. one GPU decoder per device
    (``decode.MjpegVideo(path, device, first_frame=shard.load_start, end_frame=shard.load_end)``); the callable flavour
    needs `n_frames` and `frame_size` (width, height).
    """
    import torch

    devices = [int(d) for d in devices]
    if not devices:
        raise ValueError("detect_sharded needs at least one device")
    _check_shardable(detectors)
    w = halo_width(detectors)
    pre_sharded = isinstance(frames, (list, tuple))
    per_device_source = callable(frames)
    if per_device_source:
        if n_frames is None or frame_size is None:
            raise ValueError("a per-device source factory needs n_frames= and frame_size=")
        n_total = int(n_frames)
    elif pre_sharded:
        if len(frames) != len(devices):
            raise ValueError("pre-sharded frames: one tensor per device")
        n_total = None
    else:
        n_total = int(frames.shape[0])
    if pre_sharded:
        # lengths of the load ranges determine N: sum of owned sizes; recover N from the halo layout
        n_total = _total_from_loads([int(t.shape[0]) for t in frames], w)
    shards = frame_range_shards(n_total, len(devices), w)
    if n_total == 0:
        return ShardedResult(0, start_frame, {type(d).__name__: [] for d in detectors}, shards=shards)
    sample = frames[0] if pre_sharded else frames
    if per_device_source:
        width, height = int(frame_size[0]), int(frame_size[1])
    elif pixel_format in ("nv12", "i420"):
        width, height = int(sample.shape[2]), int(sample.shape[1]) * 2 // 3
    else:
        width, height = int(sample.shape[2]), int(sample.shape[1])
    flags = [type(d)._DET_FLAG for d in detectors]
    names = [type(d).__name__ for d in detectors]
    scorers: List[Optional[ShardScorer]] = [None] * len(devices)
    errors: List[BaseException] = []
    dst_dev = devices[next(g for g, sh in enumerate(shards) if sh.own_end > sh.own_start)]  # deciding GPU: first non-empty shard
    merged = torch.empty((len(flags), n_total), dtype=torch.float64, device=f"cuda:{dst_dev}")

    def run(g: int):
        try:
            sh = shards[g]
            if sh.own_end <= sh.own_start:
                return
            # the contexts number frames from 0; `start_frame` is applied when the cuts are reported
            sc = ShardScorer(detectors, width, height, devices[g], sh, batch_frames, downscale_mode, tuning, pixel_format,
                             ingest_threads)
            scorers[g] = sc
            if per_device_source:
                import torch as _torch

                with _torch.cuda.device(devices[g]), _torch.cuda.stream(sc.stream):
                    src = frames(sh, devices[g])
                    try:
                        while True:
                            b = src.read_batch(batch_frames)
                            if b is None or b.shape[0] == 0:
                                break
                            if isinstance(b, np.ndarray):
                                sc.push_host(b)
                            else:
                                sc.push_device(b)
                        # the source's buffers must outlive the scoring that reads them
                        sc.stream.synchronize()
                    finally:
                        if hasattr(src, "close"):
                            src.close()
            elif pre_sharded:
                t = frames[g]
                if int(t.shape[0]) != sh.load_end - sh.load_start:
                    raise ValueError(f"device {devices[g]}: tensor holds {int(t.shape[0])} frames, its load range has {sh.load_end - sh.load_start}")
                for a in range(0, int(t.shape[0]), batch_frames):
                    sc.push_device(t[a:a + batch_frames])
            else:
                sc.push_host(frames[sh.load_start:sh.load_end])
            sc.copy_owned(merged, dst_dev)
        except BaseException as e:  # noqa: BLE001 - re-raised on the calling thread
            errors.append(e)

    try:
        if len(devices) == 1:
            run(0)
        else:
            # one feeding thread per device: ctypes releases the GIL inside libesd, so host-side gathers and waits overlap
            threads = [threading.Thread(target=run, args=(g,), name=f"esd-shard-{g}") for g in range(len(devices))]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
        if errors:
            raise errors[0]
        with torch.cuda.device(dst_dev):
            st0 = torch.cuda.current_stream(dst_dev)
            for sc in scorers:
                if sc is not None:
                    st0.wait_event(sc.done)
            decider = next(sc for sc in scorers if sc is not None and sc.device == dst_dev)
            cuts = decide_merged(decider.ctx, flags, merged, n_total, 0, st0.cuda_stream, names)
            scores = {}
            if collect_scores:
                host = merged.cpu().numpy()
                scores = {_SCORE_NAME[f]: host[k].copy() for k, f in enumerate(flags)}
        if start_frame:
            cuts = {k: [c + start_frame for c in v] for k, v in cuts.items()}
        return ShardedResult(n_total, start_frame, cuts, scores, shards)
    finally:
        for sc in scorers:
            if sc is not None:
                sc.close()


def _total_from_loads(loads: Sequence[int], w: int) -> int:
    """N such that frame_range_shards(N, G, w) has exactly these load-range lengths."""
    g = len(loads)
    lo = max(0, sum(loads) - g * (2 * w + 1))
    for n in range(lo, sum(loads) + 1):
        if [s.load_end - s.load_start for s in frame_range_shards(n, g, w)] == list(loads):
            return n
    raise ValueError("the per-device tensors do not match any frame-range sharding (see plan_shards)")


def plan_shards(n_frames: int, n_devices: int, detectors: Sequence[SceneDetector]) -> List[FrameShard]:
    """The load / own ranges `detect_sharded` expects for pre-sharded device tensors."""
    return frame_range_shards(n_frames, n_devices, halo_width(detectors))


def detect_library(videos: Sequence, detectors: Sequence[SceneDetector], devices: Sequence[int], fps: float = 30.0,
                   batch_frames: int = 512, downscale_mode: str = "float", tuning: Optional[dict] = None,
                   ingest_threads: int = 0) -> List[dict]:
    """Whole-video partition of a library (BASELINE config 5): video j -> a device by longest-processing-time greedy on
    frame counts; every device runs the complete chain for its videos, one after another, reusing its context.
    `videos`: arrays / TensorVideo / BatchVideo objects (anything SceneManager.detect_scenes takes; CUDA tensors must
    live on the device LPT assigns -- use `sharding.partition_videos(lengths, len(devices))` to place them).
    Returns, per video, {"device", "n_frames", "cuts", "scenes"} in input order."""
    devices = [int(d) for d in devices]
    vids = [v if hasattr(v, "read_batch") else TensorVideo(v, fps) for v in videos]
    lengths = [int(getattr(getattr(v, "frames", None), "shape", [0])[0]) or 1 for v in vids]
    assign = partition_videos(lengths, len(devices))
    out: List[Optional[dict]] = [None] * len(vids)
    errors: List[BaseException] = []

    def run(g: int):
        try:
            sm = _manager(detectors, devices[g], batch_frames, downscale_mode, tuning, ingest_threads)
            try:
                for j in assign[g]:
                    n = sm.detect_scenes(vids[j], reuse_context=True)
                    out[j] = {"device": devices[g], "n_frames": n, "cuts": sm.get_cut_list(),
                              "scenes": sm.get_scene_list(start_in_scene=True)}
            finally:
                sm.close()
        except BaseException as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=run, args=(g,), name=f"esd-lib-{g}") for g in range(len(devices))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return out  # type: ignore[return-value]


# ------------------------------------------------------------------------------------ multi-process flavour (torchrun)
class DistributedShard:
    """This rank's part of a frame-range job under torch.distributed (one rank per GPU): score the load range through
    `ctx`, then `finish()` = copy the owned slices into the gather buffer, all-gather, and (rank 0) ONE decision pass.
    Everything is enqueued on `stream`; nothing visits the host until the cuts come back."""

    def __init__(self, ctx: capi.EsdContext, flags: Sequence[int], n_frames: int, window_width: int, rank: int, world: int,
                 device: int, group=None):
        import torch

        self.ctx, self.flags, self.n_frames = ctx, list(flags), int(n_frames)
        self.rank, self.world, self.group = rank, world, group
        self.shards = frame_range_shards(n_frames, world, window_width)
        self.shard = self.shards[rank]
        self.own_lens = [s.own_end - s.own_start for s in self.shards]
        self.local = torch.zeros((len(self.flags), max(1, max(self.own_lens))), dtype=torch.float64, device=f"cuda:{device}")
        self.device = device

    def finish(self, stream: int, want_scores: bool = False):
        """-> (cuts per flag on rank 0 else None, merged scores tensor [kinds, N] if want_scores)."""
        sh = self.shard
        n_own = sh.own_end - sh.own_start
        for k, flag in enumerate(self.flags):
            if n_own > 0:
                self.ctx.copy_scores_device(capi.DECISION_SCORE_KIND[flag], sh.own_start, n_own, self.local[k].data_ptr(), -1, stream)
        merged = all_gather_scores(self.local, self.own_lens, self.group)
        cuts = None
        if self.rank == 0:
            merged = merged.contiguous()
            cuts = {flag: self.ctx.decide_device(flag, 0, merged[k].data_ptr(), self.n_frames, stream) for k, flag in enumerate(self.flags)}
        return cuts, (merged if want_scores else None)
