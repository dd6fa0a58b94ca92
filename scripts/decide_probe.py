#!/usr/bin/env python
"""Host-timed cost of the stand-alone decision pass (esd_decide_device) over n frames of the config-2 golden scores, AdaptiveDetector
and ContentDetector; run under `ncu --metrics gpu__time_duration.sum -k regex:decide_kernel` for the kernel's own share."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from eioku_b200 import capi  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "clip_c2_1080p_full.npz"))
scores = np.ascontiguousarray(g["content_val"])
cfg = capi.default_config()
cfg.detectors = capi.ESD_DET_CONTENT | capi.ESD_DET_ADAPTIVE
cfg.src_width, cfg.src_height = 1920, 1080
cfg.initial_capacity = 40000
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
out = {}
with capi.EsdContext(cfg, 0) as ctx:
    d = torch.from_numpy(scores).cuda()
    st = torch.cuda.current_stream().cuda_stream
    for det, name in ((capi.ESD_DET_ADAPTIVE, "adaptive"), (capi.ESD_DET_CONTENT, "content")):
        for n in (2048, 8192, 18000):
            for _ in range(5):
                ctx.decide_device(det, 0, d.data_ptr(), n, st)
            ts = []
            for _ in range(reps):
                t0 = time.perf_counter()
                cuts = ctx.decide_device(det, 0, d.data_ptr(), n, st)
                ts.append(time.perf_counter() - t0)
            out[f"{name}_{n}"] = {"median_us": round(float(np.median(ts)) * 1e6, 1), "min_us": round(min(ts) * 1e6, 1), "cuts": len(cuts)}
print(json.dumps(out))
