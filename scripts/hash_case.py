"""Small content+hash workload for ncu: three 1024-frame 1080p pushes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eioku_b200 import capi
import synthclip as synth
W, H, n = 1920, 1080, 1024
sch = synth.build_schedule(1002, n)
clip = torch.empty((n, H, W, 3), dtype=torch.uint8, device="cuda:0")
synth.fill(clip, 1002, sch.descs)
cfg = capi.default_config()
cfg.detectors = capi.ESD_DET_CONTENT | capi.ESD_DET_HASH
cfg.src_width, cfg.src_height = W, H
with capi.EsdContext(cfg, 0) as ctx:
    for i in range(3):
        ctx.push_tensor(clip, i * n)
    ctx.synchronize()
    print(ctx.read_hash(5, 3)[1], ctx.get_cuts(capi.ESD_DET_HASH)[1])
