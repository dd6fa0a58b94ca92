#!/usr/bin/env python
"""A/B of fused-kernel variants on the non-headline geometries: sustained frames/s and roofline fraction
of fused_score_kernel for 1080p / 720p BGR24, 4K (histogram + luma_only from one pass) and 1080p NV12, with the consumer
lane stride forced to 1 (round-1 mapping) and chosen by the library (bank-conflict-free).  One JSON line per case."""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import synthclip as synth  # noqa: E402
from eioku_b200 import capi  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def run_case(name, w, h, n, detectors, fmt, tune, seconds, luma=False):
    dev = "cuda:0"
    sch = synth.build_schedule(1002, n)
    clip = torch.empty((n, h, w, 3), dtype=torch.uint8, device=dev)
    synth.fill(clip, 1002, sch.descs, chunk=128)
    if fmt == capi.ESD_FMT_NV12:
        nv = torch.empty((n, h * 3 // 2, w), dtype=torch.uint8, device=dev)
        for a in range(0, n, 256):
            nv[a:a + 256] = synth.bgr_to_test_nv12(clip[a:a + 256])
        del clip
        clip = nv
    cfg = capi.default_config()
    cfg.detectors = detectors
    cfg.src_width, cfg.src_height, cfg.src_format = w, h, fmt
    if luma:
        cfg.content_weights[0] = cfg.content_weights[1] = 0.0
        cfg.content_weight_div = 1.0
    for k, v in tune.items():
        setattr(cfg, k, v)
    stream = torch.cuda.current_stream().cuda_stream
    push = (lambda c, p: c.push_nv12_tensor(clip, p, stream)) if fmt == capi.ESD_FMT_NV12 else (lambda c, p: c.push_tensor(clip, p, stream))
    cfg.initial_capacity = 64 * n
    with capi.EsdContext(cfg, 0) as ctx:
        pos = 0
        for _ in range(3):
            push(ctx, pos); pos += n
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            push(ctx, pos); pos += n
        ctx.join(stream)
        e1.record()
        torch.cuda.synchronize()
        est = e0.elapsed_time(e1) / 4
        alg, fetched = ctx.alg_bytes_per_frame, int(ctx.geometry.compact_frame_bytes)
        lane_stride = int(ctx.geometry.lane_stride)
    steps = max(20, int(seconds * 1000 / est))
    cfg.initial_capacity = (steps + 16) * n
    cfg.max_cuts = max(65536, 64 * steps)
    with capi.EsdContext(cfg, 0) as ctx:
        pos = 0
        for _ in range(5):
            push(ctx, pos); pos += n
        ctx.synchronize()
        ctx.set_timing(True)
        e0.record()
        for _ in range(steps):
            push(ctx, pos); pos += n
        ctx.join(stream)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        kms, kn = ctx.kernel_time()
    out = {"case": name, "tune": tune, "lane_stride": lane_stride, "frames_per_s": steps * n / (ms / 1000), "fused_ms": kms / kn, "steps": steps, "seconds": ms / 1000,
           "alg_bytes_per_frame": alg, "fetched_bytes_per_frame": fetched,
           "frac_alg": n * alg / (kms / kn / 1000) / 1e9 / peak(), "frac_fetched": n * fetched / (kms / kn / 1000) / 1e9 / peak()}
    print(json.dumps(out), flush=True)
    del clip
    torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="1080p,720p,4k,nv12")
    ap.add_argument("--strides", default="1,0")
    ap.add_argument("--seconds", type=float, default=1.5)
    ap.add_argument("--tune", action="append", default=[], help="extra esd_config field=value applied to every case")
    args = ap.parse_args()
    extra = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in args.tune}
    C, H = capi.ESD_DET_CONTENT, capi.ESD_DET_HIST
    for case in args.cases.split(","):
        for ks in [int(x) for x in args.strides.split(",")]:
            tune = {"reserved1": ks, **extra}
            if case == "1080p":
                run_case(case, 1920, 1080, 2048, C, capi.ESD_FMT_BGR24, tune, args.seconds)
            elif case == "720p":
                run_case(case, 1280, 720, 2048, C, capi.ESD_FMT_BGR24, tune, args.seconds)
            elif case == "4k":
                run_case(case, 3840, 2160, 512, C | H, capi.ESD_FMT_BGR24, tune, args.seconds, luma=True)
            elif case == "4k_content":
                run_case(case, 3840, 2160, 512, C, capi.ESD_FMT_BGR24, tune, args.seconds)
            elif case == "nv12":
                run_case(case, 1920, 1080, 2048, C, capi.ESD_FMT_NV12, tune, args.seconds)


if __name__ == "__main__":
    main()
