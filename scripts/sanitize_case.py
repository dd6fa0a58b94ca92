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

# round 2: YUV source formats, the hash detector, the stand-alone decision pass and the JPEG decoder
import cv2  # noqa: E402
from eioku_b200 import decode  # noqa: E402
from eioku_b200.detectors import HashDetector  # noqa: E402

w, h, n = 320, 180, 12
sch = synth.build_schedule(9, n, min_len=4, max_len=6)
bgr = torch.empty((n, h, w, 3), dtype=torch.uint8, device=dev)
synth.fill(bgr, 9, sch.descs)
nv12 = synth.bgr_to_test_nv12(bgr)
i420 = synth.nv12_to_i420(nv12)
for frames, fmt in ((nv12, "nv12"), (i420, "i420"), (nv12.cpu().numpy(), "nv12"), (i420.cpu().numpy(), "i420")):
    for threads in ((0,) if not isinstance(frames, np.ndarray) else (0, 2)):
        sm = SceneManager(batch_frames=5)
        sm._downscale, sm._auto_downscale = 2, False
        sm.add_detector(ContentDetector()); sm.add_detector(HistogramDetector())
        sm._ingest_threads = threads
        print(fmt, type(frames).__name__, threads, sm.detect_scenes(TensorVideo(frames, 30.0, pixel_format=fmt)), sm.get_cut_list())
        sm.close()
sm = SceneManager(batch_frames=7)
sm.add_detector(HashDetector()); sm.add_detector(ContentDetector())
print("hash", sm.detect_scenes(TensorVideo(bgr, 30.0), collect_scores=True), sm.get_cut_list())
ctx = sm._ctx
d = torch.from_numpy(np.ascontiguousarray(sm.scores["content_val"])).to(dev)
print("decide_device", ctx.decide_device(capi.ESD_DET_CONTENT, 0, d.data_ptr(), d.numel(), torch.cuda.current_stream().cuda_stream))
sm.close()
path = f"/tmp/esd_sanitize_{os.getpid()}.avi"
wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 25.0, (w, h))
for f in bgr.cpu().numpy():
    wr.write(f)
wr.release()
for lanes in ("2", "1"):
    os.environ["ESD_DEC_LANES"] = lanes
    with decode.MjpegVideo(path, batch_frames=5, backend=decode.ESD_JPEG_NATIVE) as v:
        sm = SceneManager(batch_frames=5)
        sm.add_detector(ContentDetector())
        print("decode lanes", lanes, sm.detect_scenes(v), sm.get_cut_list())
        sm.close()
os.remove(path)
print("sanitize case (round 2) done")
