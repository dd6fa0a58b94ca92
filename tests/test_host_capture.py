"""Host decode by frame range (eioku_b200.decode.CaptureRangeVideo): what lets several cv2.VideoCapture instances decode one file at
once (service: `decode_workers`).  No GPU: the ranges must reproduce the sequential decode frame for frame, and the fingerprints the
service compares must agree between captures."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

import synthclip as synth  # noqa: E402
from eioku_b200 import decode  # noqa: E402
from eioku_b200.sharding import frame_range_shards  # noqa: E402
from oracle import c_oracle as co  # noqa: E402


@pytest.fixture(scope="module", params=[("mp4v", ".mp4"), ("MJPG", ".avi")])
def clip(request, tmp_path_factory):
    fourcc, ext = request.param
    n, w, h = 130, 320, 180
    sch = synth.build_schedule(31, n, min_len=10, max_len=30)
    frames = co.synth_frames(31, w, h, sch.descs)
    path = str(tmp_path_factory.mktemp("cap") / f"clip{ext}")
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*fourcc), 25.0, (w, h))
    if not wr.isOpened():
        pytest.skip(f"this OpenCV build cannot write {fourcc}")
    for f in frames:
        wr.write(f)
    wr.release()
    cap = cv2.VideoCapture(path)
    seq = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        seq.append(f)
    cap.release()
    return path, np.stack(seq)


def _read_all(src):
    out = []
    while True:
        b = src.read_batch(0)
        if b is None:
            break
        out.append(b.copy())   # the source reuses its buffer
    return np.concatenate(out) if out else np.empty((0,))


def test_ranges_reproduce_the_sequential_decode(clip):
    path, seq = clip
    n = seq.shape[0]
    with decode.CaptureRangeVideo(path, batch_frames=7) as v:
        assert v.n_frames == n and v.frame_size == (320, 180) and v.frame_rate == 25.0 and v.start_frame == 0
        assert np.array_equal(_read_all(v), seq)
    shards = frame_range_shards(n, 5, 2)     # AdaptiveDetector(window_width=2): halo 3 / 2
    starts = [s.load_start for s in shards if s.load_start > 0]
    srcs = [decode.CaptureRangeVideo(path, s.load_start, s.load_end, batch_frames=11, watch=starts) for s in shards]
    for s, src in zip(shards, srcs):
        assert src.start_frame == s.load_start
        got = _read_all(src)
        assert got.shape[0] == s.load_end - s.load_start
        assert np.array_equal(got, seq[s.load_start:s.load_end]), s
        src.close()
    # the fingerprints the service compares: every range's first frame as seen by the range that owns it
    for s, src in zip(shards, srcs):
        if s.load_start == 0:
            continue
        owner = next(k for k, o in enumerate(shards) if o.own_start <= s.load_start < o.own_end)
        assert srcs[owner].digests[s.load_start] == src.digests[s.load_start]
    assert len({d for src in srcs for d in src.digests.values()}) > 1     # and they tell frames apart


def test_a_range_past_the_end_comes_up_short_and_missing_files_raise(clip, tmp_path):
    path, seq = clip
    n = seq.shape[0]
    with decode.CaptureRangeVideo(path, n - 4, n + 50, batch_frames=3) as v:   # clamped to the container's frame count
        assert _read_all(v).shape[0] == 4
    with pytest.raises(RuntimeError, match="Failed to open video"):
        decode.CaptureRangeVideo(str(tmp_path / "nope.mp4"))


def test_which_files_get_several_captures(clip, tmp_path, monkeypatch):
    """service._wants_capture_workers / default_decode_workers: host-decoded containers only, never for options the sharded
    path does not carry, never for detectors that cannot shard, and short clips stay sequential."""
    from eioku_b200 import service

    path, _ = clip
    is_avi = path.endswith(".avi")
    assert service._wants_capture_workers(path, {}) == (not is_avi)            # Motion-JPEG AVI goes to the GPU decoder instead
    assert service._wants_capture_workers(path, {"gpu_decode": False}) is True
    assert service._wants_capture_workers(path, {"gpu_decode": False, "decode_workers": 1}) is False
    assert service._wants_capture_workers(path, {"gpu_decode": False, "downscale": 2}) is False
    assert service._wants_capture_workers(path, {"gpu_decode": False, "auto_downscale": False}) is False
    assert service._wants_capture_workers(path, {"gpu_decode": False, "artifact_payloads": True}) is False
    assert service._wants_capture_workers(str(tmp_path / "frames.npy"), {}) is False
    # ThresholdDetector(add_final_scene=True) keeps state the global decision pass does not have: not shardable
    assert service._wants_capture_workers(path, {"gpu_decode": False, "detector": "threshold", "add_final_scene": True}) in (False, True)
    monkeypatch.setattr("os.sched_getaffinity", lambda pid: set(range(16)), raising=False)
    assert service.default_decode_workers(2047) == 1
    assert service.default_decode_workers(2048) == 8
    monkeypatch.setattr("os.sched_getaffinity", lambda pid: set(range(3)), raising=False)
    assert service.default_decode_workers(100000) == 1
    monkeypatch.setattr("os.sched_getaffinity", lambda pid: set(range(64)), raising=False)
    assert service.default_decode_workers(100000) == 8


def test_a_container_that_under_reports_its_length_is_noticed(clip):
    """The range that ends at the (claimed) end of the file tries one more frame: `trailing_frames` is what sends the service
    back to the sequential decode instead of silently dropping the tail."""
    path, seq = clip
    n = seq.shape[0]
    with decode.CaptureRangeVideo(path, n - 20, None, batch_frames=8) as v:
        assert _read_all(v).shape[0] == 20 and v.trailing_frames is False
    with decode.CaptureRangeVideo(path, n - 30, None, batch_frames=8, frame_count=n - 6) as v:   # "the container says n - 6"
        assert _read_all(v).shape[0] == 24 and v.trailing_frames is True
    with decode.CaptureRangeVideo(path, 0, 40, batch_frames=8) as v:   # a range in the middle never probes the end
        assert _read_all(v).shape[0] == 40 and v.trailing_frames is None
    with decode.CaptureRangeVideo(path, 0, None, batch_frames=16, until_eof=True, frame_count=10) as v:   # whole-file sources ignore the claim
        assert _read_all(v).shape[0] == n and v.trailing_frames is None
