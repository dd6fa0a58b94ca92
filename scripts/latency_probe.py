"""Per-call latency of esd_process_frame_host (one 256x144 host frame per call) for a few kernel shapes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from eioku_b200 import capi

rng = np.random.default_rng(0)
frames = rng.integers(0, 256, (400, 144, 256, 3), dtype=np.uint8)
for det in (capi.ESD_DET_CONTENT, capi.ESD_DET_HIST):
    for R, RS, ST in ((0, 0, 0), (16, 4, 2), (8, 4, 2), (4, 4, 2), (4, 2, 2), (2, 2, 2), (2, 1, 2), (1, 1, 2), (4, 1, 4)):
        cfg = capi.default_config()
        cfg.detectors = det
        cfg.src_width, cfg.src_height, cfg.dst_width, cfg.dst_height = 256, 144, 256, 144
        cfg.rows_per_group, cfg.rows_per_stage, cfg.pipeline_stages = R, RS, ST
        ctx = capi.EsdContext(cfg, 0)
        seen = 0
        for k in range(100):
            _, seen = ctx.process_frame_host(frames[k], k, det, seen)
        t0 = time.perf_counter()
        for k in range(100, 400):
            _, seen = ctx.process_frame_host(frames[k], k, det, seen)
        us = (time.perf_counter() - t0) / 300 * 1e6
        print(f"det={det} rows_per_group={R} rows_per_stage={RS} stages={ST}: {us:.1f} us/frame", flush=True)
        ctx.close()
