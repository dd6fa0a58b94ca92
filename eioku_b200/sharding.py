"""Multi-GPU partitioning: independent shards, no data-path collective.

* whole videos (library batch): video j -> a GPU, longest-processing-time greedy;
* frame ranges of one long video: shard g owns frames [s_g, e_g) and additionally loads a left
  halo of ``window_width + 1`` frames (1 for the previous-frame delta, w for the adaptive window)
  and a right halo of ``window_width`` frames.  Shards return *scores*; the min-scene-len / flash
  filter state machines are sequential, so ONE decision pass runs over the concatenated arrays
  (``EsdContext.decide_arrays``) -- per-shard cut lists cannot be merged exactly.
Only score arrays (~40 B/frame) move between ranks, as plain float64 tensors through
``torch.distributed.all_gather_into_tensor`` (NCCL on device tensors, gloo on CPU tensors in the tests) -- no pickled
objects on this path.  The single-process, several-devices flavour lives in ``eioku_b200.multi``.
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


def all_gather_scores(local, own_lens: Sequence[int], group=None):
    """Owned score slices of every rank -> one [kinds, N] tensor on every rank, as plain float64 tensors (no pickle).

    `local`: [kinds, max(own_lens)] float64 tensor (this rank's slice in the first own_lens[rank] columns; CUDA under
    NCCL, CPU under gloo).  Shards own consecutive frame ranges in rank order."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    kinds, width = int(local.shape[0]), int(local.shape[1])
    if world == 1:
        return local[:, :own_lens[0]]
    flat = torch.empty((world * kinds, width), dtype=local.dtype, device=local.device)  # concatenation along dim 0 (gloo and NCCL)
    dist.all_gather_into_tensor(flat, local.contiguous(), group=group)
    buf = flat.view(world, kinds, width)
    if all(n == width for n in own_lens):
        return buf.permute(1, 0, 2).reshape(kinds, world * width)
    return torch.cat([buf[g, :, :own_lens[g]] for g in range(world)], dim=1)


def _pack(part: Dict[str, np.ndarray], schema, n_own: int, width: int):
    """dict of per-frame arrays -> [columns, width] float64 matrix (integers up to 2^53 are exact in float64)."""
    import torch

    cols = sum(c for _, c, _ in schema)
    m = np.zeros((cols, width), np.float64)
    r = 0
    for key, c, _dt in schema:
        if n_own:
            a = np.asarray(part[key]).reshape(n_own, c)
            if a.dtype.kind in "iu" and a.size and int(a.max()) >= (1 << 53):
                raise ValueError(f"{key}: integers >= 2^53 do not survive the float64 transport")
            m[r:r + c, :n_own] = a.T
        r += c
    return torch.from_numpy(m)


def _unpack(mat, schema) -> Dict[str, np.ndarray]:
    out, r = {}, 0
    a = mat.cpu().numpy()
    for key, c, dt in schema:
        blk = a[r:r + c].T
        blk = blk[:, 0] if c == 1 else blk
        out[key] = np.ascontiguousarray(blk).astype(dt) if np.dtype(dt) != np.float64 else np.ascontiguousarray(blk)
        r += c
    return out


def sharded_detect(score_shard: Callable[[FrameShard], Dict[str, np.ndarray]],
                   decide: Callable[[Dict[str, np.ndarray]], Dict[str, List[int]]],
                   n_frames: int, window_width: int = 0, group=None, schema=None) -> Optional[Dict[str, List[int]]]:
    """Run one frame-range shard per rank of the default (or given) process group.

    score_shard(shard) -> per-loaded-frame score arrays for this rank's shard (frames
    [load_start, load_end)); decide(merged) -> cuts per detector.  Rank 0 returns the cuts,
    other ranks return None.  Without an initialised process group this is a 1-rank job.
    The owned slices travel as ONE float64 matrix per rank (``all_gather_scores``).  schema: [(key, columns, dtype)] of the
    score arrays; inferred from this rank's arrays when every rank owns frames (required when a shard is empty).
    """
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    shards = frame_range_shards(n_frames, world, window_width)
    mine = shards[rank]
    own_lens = [s.own_end - s.own_start for s in shards]
    n_own = own_lens[rank]
    part = owned_slice(score_shard(mine), mine) if n_own > 0 else {}
    if schema is None:
        if min(own_lens) == 0:
            raise ValueError("a shard owns no frames: pass schema=[(key, columns, dtype), ...] so every rank packs the same matrix")
        schema = [(k, int(np.asarray(part[k]).reshape(n_own, -1).shape[1]), np.asarray(part[k]).dtype) for k in sorted(part)]
    merged_mat = all_gather_scores(_pack(part, schema, n_own, max(1, max(own_lens))), own_lens, group)
    if rank != 0:
        return None
    return decide(fix_video_start(_unpack(merged_mat, schema)))
