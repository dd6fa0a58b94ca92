"""Small content+edges run (ncu target for edges_kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eioku_b200 import capi
import synthclip as synth
from eioku_b200.detectors import ContentDetector
from eioku_b200.scene_manager import SceneManager
W, H, n = 1920, 1080, 296
sch = synth.build_schedule(1002, n)
clip = torch.empty((n, H, W, 3), dtype=torch.uint8, device="cuda:0")
synth.fill(clip, 1002, sch.descs)
sm = SceneManager()
sm.add_detector(ContentDetector(weights=ContentDetector.Components(1.0, 1.0, 1.0, 1.0)))
ctx = sm.make_context(W, H)
for i in range(3):
    ctx.push_tensor(clip, i * n)
ctx.synchronize()
print("done", ctx.get_cuts(capi.ESD_DET_CONTENT)[1])
