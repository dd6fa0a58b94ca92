"""One small native-decoder run for ncu: a 64-frame 1080p MJPEG file, two batches of 32."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eioku_b200 import decode
from scripts.decode_probe import make_file
path = "/dev/shm/esd_case.avi"
if not os.path.exists(path):
    make_file(path, 64)
with decode.MjpegVideo(path, batch_frames=32, backend=decode.ESD_JPEG_NATIVE) as v:
    n = 0
    while True:
        b = v.read_batch(0)
        if b is None:
            break
        n += b.shape[0]
    torch.cuda.synchronize()
print("decoded", n)
