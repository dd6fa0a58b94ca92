"""Small NV12 workload for ncu: four 1024-frame 1080p NV12 pushes of ContentDetector."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eioku_b200 import capi
import synthclip as synth
W, H, n = 1920, 1080, 1024
sch = synth.build_schedule(1002, n)
nv12 = torch.empty((n, H * 3 // 2, W), dtype=torch.uint8, device="cuda:0")
for a in range(0, n, 256):
    bgr = torch.empty((256, H, W, 3), dtype=torch.uint8, device="cuda:0")
    synth.fill(bgr, 1002, sch.descs[a:a + 256])
    nv12[a:a + 256] = synth.bgr_to_test_nv12(bgr)
cfg = capi.default_config()
cfg.src_width, cfg.src_height, cfg.src_format = W, H, capi.ESD_FMT_NV12
with capi.EsdContext(cfg, 0) as ctx:
    for i in range(4):
        ctx.push_nv12_tensor(nv12, i * n)
    ctx.synchronize()
    print(ctx.get_cuts(capi.ESD_DET_CONTENT)[1])
