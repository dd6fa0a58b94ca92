"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): every kernel of libesd.so once, tiny sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from eioku_b200 import capi
import synthclip as synth
from eioku_b200.detectors import AdaptiveDetector, ContentDetector, HistogramDetector, ThresholdDetector
from eioku_b200.scene_manager import SceneManager, TensorVideo

dev = "cuda:0"
for (w, h, n) in ((640, 360, 24), (333, 77, 9), (96, 54, 7)):
    sch = synth.build_schedule(7, n, min_len=5, max_len=9)
    clip = torch.empty((n, h, w, 3), dtype=torch.uint8, device=dev)
    synth.fill(clip, 7, sch.descs)
    sm = SceneManager(batch_frames=10)
    for d in (ContentDetector(weights=ContentDetector.Components(1, 1, 1, 1)), AdaptiveDetector(weights=ContentDetector.Components(1, 1, 1, 1)),
              HistogramDetector(bins=100), ThresholdDetector(add_final_scene=True)):
        sm.add_detector(d)
    print(w, h, sm.detect_scenes(TensorVideo(clip, 30.0), collect_scores=True), sm.get_cut_list())
    sm.close()
    host = clip.cpu().numpy()
    sm = SceneManager(batch_frames=8)
    sm.add_detector(ContentDetector())
    sm.detect_scenes(TensorVideo(host, 30.0))   # ingest ring, pageable source
    sm.close()
print("sanitize case done")
