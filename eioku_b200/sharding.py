"""Multi-GPU partitioning: independent shards, no data-path collective.

* whole videos (library batch): video j -> a GPU, longest-processing-time greedy;
* frame ranges of one long video: shard g owns frames [s_g, e_g) and additionally loads a left
  halo of ``window_width + 1`` frames (1 for the previous-frame delta, w for the adaptive window)
  and a right halo of ``window_width`` frames.  Shards return *scores*; the min-scene-len / flash
  filter state machines are sequential, so ONE decision pass runs over the concatenated arrays
  (``EsdContext.decide_arrays``) -- per-shard cut lists cannot be merged exactly.
Only small host objects (score arrays, ~40 B/frame) move between ranks, through
``torch.distributed.all_gather_object`` (any backend; gloo in the CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np


@dataclass(frozen=True)
class FrameShard:
    rank: int
    own_start: int  # first owned frame
    own_end: int    # one past the last owned frame
    load_start: int  # first frame that must be resident (own_start - left halo, clamped)
    load_end: int    # one past the last resident frame (own_end + right halo, clamped)


def frame_range_shards(n_frames: int, world_size: int, window_width: int = 0) -> List[FrameShard]:
    """Even split of [0, n_frames) with the halo the detectors need."""
    if n_frames < 0 or world_size < 1:
        raise ValueError("bad shard request")
    left = window_width + 1
    right = window_width
    out = []
    for g in range(world_size):
        s = n_frames * g // world_size
        e = n_frames * (g + 1) // world_size
        out.append(FrameShard(g, s, e, max(0, s - left) if e > s else s, min(n_frames, e + right) if e > s else e))
    return out


def partition_videos(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Longest-processing-time greedy: video indices per rank, balanced by frame count."""
    loads = [0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for j in sorted(range(len(lengths)), key=lambda i: (-lengths[i], i)):
        g = min(range(world_size), key=lambda r: (loads[r], r))
        out[g].append(j)
        loads[g] += lengths[j]
    for lst in out:
        lst.sort()
    return out


def owned_slice(scores: Dict[str, np.ndarray], shard: FrameShard) -> Dict[str, np.ndarray]:
    """Cut a shard's per-loaded-frame arrays down to its owned range."""
    a = shard.own_start - shard.load_start
    b = a + (shard.own_end - shard.own_start)
    return {k: np.asarray(v)[a:b] for k, v in scores.items()}


def merge_owned(parts: Sequence[Dict[str, np.ndarray]], shards: Sequence[FrameShard]) -> Dict[str, np.ndarray]:
    """Concatenate owned-range arrays in frame order, checking that the shards tile the video."""
    order = sorted(range(len(shards)), key=lambda i: shards[i].own_start)
    pos = shards[order[0]].own_start if order else 0
    for i in order:
        if shards[i].own_start != pos:
            raise ValueError("shards do not tile the frame range")
        pos = shards[i].own_end
    keys = parts[order[0]].keys() if order else []
    return {k: np.concatenate([np.asarray(parts[i][k]) for i in order]) for k in keys}


def fix_video_start(merged: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """Frame 0 of the *video* has no predecessor: score 0.0, no histogram difference."""
    for k in ("content_val", "adaptive_val"):
        if k in merged and merged[k].size:
            merged[k][0] = 0.0
    if "sums3" in merged and merged["sums3"].size:
        merged["sums3"][0] = 0
    for k in ("hist_diff", "hash_dist"):
        if k in merged and merged[k].size:
            merged[k][0] = np.nan
    return merged


def sharded_detect(score_shard: Callable[[FrameShard], Dict[str, np.ndarray]],
                   decide: Callable[[Dict[str, np.ndarray]], Dict[str, List[int]]],
                   n_frames: int, window_width: int = 0, group=None) -> Optional[Dict[str, List[int]]]:
    """Run one frame-range shard per rank of the default (or given) process group.

    score_shard(shard) -> per-loaded-frame score arrays for this rank's shard (frames
    [load_start, load_end)); decide(merged) -> cuts per detector.  Rank 0 returns the cuts,
    other ranks return None.  Without an initialised process group this is a 1-rank job.
    """
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    shards = frame_range_shards(n_frames, world, window_width)
    mine = shards[rank]
    part = owned_slice(score_shard(mine), mine) if mine.own_end > mine.own_start else {}
    if world > 1:
        gathered: List[Optional[dict]] = [None] * world
        dist.all_gather_object(gathered, part, group=group)
    else:
        gathered = [part]
    if rank != 0:
        return None
    live = [i for i in range(world) if shards[i].own_end > shards[i].own_start]
    merged = fix_video_start(merge_owned([gathered[i] for i in live], [shards[i] for i in live]))
    return decide(merged)
