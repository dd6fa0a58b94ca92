"""Per-frame plugin surface on the GPU: SceneDetector.process_frame(frame_num, numpy frame) through the library's
single-call fast path (esd_process_frame_host), the opt-in deferred mode, and their equality with the batched path."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from eioku_b200 import capi  # noqa: E402
from eioku_b200.detectors import (AdaptiveDetector, ContentDetector, HashDetector, HistogramDetector,  # noqa: E402
                                  StatsManager, ThresholdDetector)
from oracle import psd_cv2 as P  # noqa: E402


def small_clip(n=150, w=256, h=144, seed=3):
    """Scenes of flat-ish colour with noise, a fade through black and a flash."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, h, w, 3), np.uint8)
    base = rng.integers(30, 226, (3,))
    for k in range(n):
        if k % 37 == 0:
            base = rng.integers(30, 226, (3,))
        f = np.clip(base[None, None, :] + rng.integers(-8, 9, (h, w, 3)), 0, 255)
        if 60 <= k < 70:
            f = f * (70 - k) // 10
        if 70 <= k < 80:
            f = f * (k - 70) // 10
        out[k] = f.astype(np.uint8)
    if n > 100:
        out[100] = 255
    return out


DETS = [
    (lambda: ContentDetector(threshold=20.0, min_scene_len=5), lambda: P.ContentDetector(threshold=20.0, min_scene_len=5)),
    (lambda: AdaptiveDetector(adaptive_threshold=2.5, min_scene_len=5, window_width=3),
     lambda: P.AdaptiveDetector(adaptive_threshold=2.5, min_scene_len=5, window_width=3)),
    (lambda: HistogramDetector(threshold=0.1, bins=64, min_scene_len=5), lambda: P.HistogramDetector(threshold=0.1, bins=64, min_scene_len=5)),
    (lambda: ThresholdDetector(threshold=20, min_scene_len=5, add_final_scene=True),
     lambda: P.ThresholdDetector(threshold=20, min_scene_len=5, add_final_scene=True)),
    (lambda: HashDetector(threshold=0.3, min_scene_len=5), lambda: P.HashDetector(threshold=0.3, min_scene_len=5)),
]


@pytest.mark.parametrize("which", range(len(DETS)))
def test_process_frame_fast_path_equals_oracle_frame_by_frame(which):
    """Every call returns exactly what PySceneDetect's process_frame returns for that frame (cv2-backed oracle)."""
    clip = small_clip()
    det, ref = DETS[which][0](), DETS[which][1]()
    if isinstance(det, HashDetector):
        # flat frames make every AC coefficient rounding noise, so hash bits are implementation-defined there (cv2.dct's
        # IPP and plain paths disagree too); cv2 parity lives in test_gpu_hash.py -- here the reference is the batched path
        ref = _Batched(DETS[which][0](), clip)
    for k in range(len(clip)):
        frame = clip[k] if k % 3 else np.ascontiguousarray(clip[k])
        assert det.process_frame(k, frame) == ref.process_frame(k, clip[k]), f"frame {k}"
    assert det.post_process(len(clip) - 1) == ref.post_process(len(clip) - 1)
    det.close()


class _Batched:
    """Per-frame view of one batched run of a detector (cuts attributed to the frame that emits them: for detectors
    without look-ahead that is the cut frame itself)."""

    def __init__(self, det, clip):
        self.cuts = set(det.process_frames(0, clip))
        det.close()

    def process_frame(self, k, _frame):
        return [k] if k in self.cuts else []

    def post_process(self, _k):
        return []


def test_fast_path_pitched_rows_and_mixed_entry_points():
    clip = small_clip(90)
    padded = np.zeros((90, 144, 300, 3), np.uint8)
    padded[:, :, 20:276] = clip
    a, b = ContentDetector(threshold=20.0, min_scene_len=5), ContentDetector(threshold=20.0, min_scene_len=5)
    sm_a, sm_b = StatsManager(), StatsManager()
    a.stats_manager, b.stats_manager = sm_a, sm_b
    cuts_a, cuts_b = [], []
    for k in range(40):
        cuts_a += a.process_frame(k, padded[k, :, 20:276])           # strided view: row pitch 900 bytes
    cuts_a += a.process_frames(40, clip[40:70])                      # batched entry point mid-stream
    for k in range(70, 90):
        cuts_a += a.process_frame(k, clip[k])
    cuts_b += b.process_frames(0, clip)
    assert cuts_a == cuts_b and len(cuts_b) >= 2
    for k in (1, 39, 40, 69, 70, 89):
        assert sm_a.get_metrics(k, ContentDetector.METRIC_KEYS[:4]) == sm_b.get_metrics(k, ContentDetector.METRIC_KEYS[:4])
    with pytest.raises(capi.EsdError):
        a.process_frame(95, clip[0])   # frame numbers must stay sequential
    with pytest.raises(ValueError):
        a.process_frame(90, clip[0][:100])
    a.close(); b.close()


@pytest.mark.parametrize("which", range(len(DETS)))
@pytest.mark.parametrize("defer", [2, 16, 64])
def test_deferred_mode_reports_the_same_cuts_late(which, defer):
    clip = small_clip()
    det, ref = DETS[which][0]().defer(defer), DETS[which][1]()
    if isinstance(det, HashDetector):
        ref = _Batched(DETS[which][0](), clip)
    own = getattr(det, "window_width", 0)
    assert det.event_buffer_length == own + defer - 1
    got, want = [], []
    for k in range(len(clip)):
        got += det.process_frame(k, clip[k])
        want += ref.process_frame(k, clip[k])
        assert got == want[:len(got)]
    got += det.post_process(len(clip) - 1)
    want += ref.post_process(len(clip) - 1)
    assert got == want
    det.close()


def test_scene_artifact_payloads_of_the_artifact_envelope_spec():
    """config["artifact_payloads"]: SceneV1 {scene_index, method, score, frame_number} records (reference spec
    .kiro/specs/artifact-envelope-architecture/design.md:159-167) next to the unchanged task schema."""
    from eioku_b200.service import detect_scenes_frames

    clip = small_clip(150)
    plain = detect_scenes_frames(clip, {"detector": "content", "threshold": 20.0, "min_scene_len": 5, "fps": 30.0})
    rich = detect_scenes_frames(clip, {"detector": "content+hist", "threshold": 20.0, "min_scene_len": 5, "fps": 30.0,
                                       "hist_threshold": 0.1, "bins": 64, "artifact_payloads": True})
    assert set(plain) == {"scenes"} and len(plain["scenes"]) >= 3
    recs = rich["artifact_payloads"]
    assert len(recs) == len(rich["scenes"]) and recs[0] == {"scene_index": 0, "method": "start", "score": 0.0, "frame_number": 0}
    ref = P.ContentDetector(threshold=20.0, min_scene_len=5)
    cuts = []
    for k in range(len(clip)):
        cuts += ref.process_frame(k, clip[k])
    by_frame = {r["frame_number"]: r for r in recs[1:]}
    for c in cuts:
        assert by_frame[c]["method"] == "content" and by_frame[c]["score"] == ref.scores[c] >= 20.0
    for i, r in enumerate(recs):
        assert r["scene_index"] == i and set(r) == {"scene_index", "method", "score", "frame_number"}
        assert int(r["frame_number"] / 30.0 * 1000) == rich["scenes"][i]["start_ms"]
        assert r["method"] in ("start", "content", "hist")
