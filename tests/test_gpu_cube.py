"""GPU: the per-pixel colour stages over ALL 2^24 BGR values, against checksums of real cv2 (tests/golden/cube.json, written
by tests/golden/make_golden.py).  VERDICT r1 weak #1 / task 5c: the exhaustive checks used to run on the CPU oracle only.

  * BGR2HSV: one 4096 x 4096 frame holding every colour once, scored without a resize; the device's packed HSV of that
    frame (esd_debug_read_prev) must hash to cv2.cvtColor(COLOR_BGR2HSV)'s bytes.
  * BGR2YUV's Y (HistogramDetector): the cube as 4096 frames of one 4096-pixel row; the 256-bin histogram of every row must
    equal numpy.bincount of cv2's Y plane row by row (a wrong Y moves a count between two bins of its row).
  * BGR2GRAY (HashDetector): the cube as 4096 frames of 64 x 64 with a 64 x 64 hash thumbnail (INTER_AREA at scale 1 is the
    identity), so the thumbnail bytes are the kernel's gray plane; they must hash to cv2.cvtColor(COLOR_BGR2GRAY)'s bytes.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from eioku_b200 import capi  # noqa: E402


@pytest.fixture(scope="module")
def cube():
    ref = json.load(open(os.path.join(GOLDEN, "cube.json")))
    x = torch.arange(1 << 24, dtype=torch.int32, device="cuda:0")
    img = torch.stack([x & 255, (x >> 8) & 255, (x >> 16) & 255], -1).to(torch.uint8).contiguous()  # index = b | g << 8 | r << 16
    return ref, img


def test_bgr2hsv_all_colours_on_the_device(cube):
    ref, img = cube
    cfg = capi.default_config()
    cfg.detectors = capi.ESD_DET_CONTENT
    cfg.src_width = cfg.src_height = cfg.dst_width = cfg.dst_height = 4096
    with capi.EsdContext(cfg, 0) as ctx:
        assert ctx.dst_size == (4096, 4096)
        ctx.push_tensor(img.view(1, 4096, 4096, 3), 0)
        hsv = ctx.debug_last_hsv()
    assert hsv.shape == (4096, 4096, 3) and int(hsv[..., 0].max()) == 179
    assert hashlib.sha256(hsv.tobytes()).hexdigest() == ref["hsv_sha256"]


def test_bgr2y_all_colours_on_the_device(cube):
    ref, img = cube
    cfg = capi.default_config()
    cfg.detectors = capi.ESD_DET_HIST
    cfg.hist_bins = 256
    cfg.src_width, cfg.src_height, cfg.dst_width, cfg.dst_height = 4096, 1, 4096, 1
    cfg.initial_capacity = 4096
    with capi.EsdContext(cfg, 0) as ctx:
        ctx.push_tensor(img.view(4096, 1, 4096, 3), 0)
        hist = ctx.read_scores(0, 4096, ["hist"])["hist"]
    assert hist.shape == (4096, 256) and hist.dtype == np.uint32 and np.all(hist.sum(1) == 4096)
    assert hashlib.sha256(np.ascontiguousarray(hist).tobytes()).hexdigest() == ref["y_rowhist_sha256"]


def test_bgr2gray_all_colours_on_the_device(cube):
    ref, img = cube
    cfg = capi.default_config()
    cfg.detectors = capi.ESD_DET_HASH
    cfg.hash_size, cfg.hash_lowpass = 32, 2   # 64 x 64 thumbnail of a 64 x 64 frame: INTER_AREA is the identity
    cfg.src_width = cfg.src_height = cfg.dst_width = cfg.dst_height = 64
    cfg.initial_capacity = 4096
    with capi.EsdContext(cfg, 0) as ctx:
        ctx.push_tensor(img.view(4096, 64, 64, 3), 0)
        gray = np.stack([ctx.debug_hash_input(f) for f in range(4096)])
    assert gray.shape == (4096, 64, 64)
    assert hashlib.sha256(gray.tobytes()).hexdigest() == ref["gray_sha256"]
