"""GPU: compressed video -> device frames -> scores.  Parity is defined on the DECODED surface: what the
GPU decoder produced is downloaded, the oracle (PySceneDetect logic on the closed forms / cv2) runs on exactly those frames,
and sums / scores / cuts must be bit-exact.  That the decoder decodes the right pictures is checked separately against
cv2.VideoCapture's decode of the same file (PSNR: two JPEG decoders differ in IDCT and chroma-upsampling rounding)."""
import asyncio
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
cv2 = pytest.importorskip("cv2")

import synthclip as synth  # noqa: E402
from eioku_b200 import decode, service  # noqa: E402
from eioku_b200.detectors import AdaptiveDetector, ContentDetector, HistogramDetector  # noqa: E402
from eioku_b200.scene_manager import SceneManager  # noqa: E402
from oracle import c_oracle as co  # noqa: E402
from oracle import psd_cv2 as P  # noqa: E402


def _write_avi(path, frames, fourcc="MJPG", fps=25.0):
    h, w = frames.shape[1:3]
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*fourcc), fps, (w, h))
    assert wr.isOpened()
    for f in frames:
        wr.write(f)
    wr.release()


@pytest.fixture(scope="module")
def mjpeg_clip(tmp_path_factory):
    d = tmp_path_factory.mktemp("mjpeg")
    n, w, h, seed = 150, 640, 360, 77
    sch = synth.build_schedule(seed, n, min_len=20, max_len=50)
    frames = co.synth_frames(seed, w, h, sch.descs)
    path = str(d / "clip.avi")
    _write_avi(path, frames)
    return path, frames


def _decode_all(path, **kw):
    with decode.MjpegVideo(path, **kw) as v:
        out = []
        while True:
            b = v.read_batch(0)
            if b is None:
                break
            out.append(b.cpu().numpy())  # copies: the tensor aliases the decoder's ring
        return np.concatenate(out), v.frame_rate, v.backend, v.n_frames


def _avi_pictures(path):
    """The JPEG pictures of an AVI written by cv2.VideoWriter (movi list of 00dc chunks)."""
    import struct

    b = open(path, "rb").read()
    i = b.find(b"movi") + 4
    out = []
    while i + 8 <= len(b) and b[i:i + 4] in (b"00dc", b"00db"):
        sz = struct.unpack("<I", b[i + 4:i + 8])[0]
        out.append(b[i + 8:i + 8 + sz])
        i += 8 + sz + (sz & 1)
    return out


def test_native_decoder_is_bit_identical_to_cv2_imdecode(mjpeg_clip, tmp_path):
    """The library's own kernels (Huffman decode, libjpeg's ISLOW IDCT, fancy upsampling, JFIF colour conversion) against the
    real reference decoder, picture by picture, every byte: the decoder itself is pinned, not only the scoring behind it."""
    path, src = mjpeg_clip
    got, fps, backend, n = _decode_all(path, batch_frames=40, backend=decode.ESD_JPEG_NATIVE)
    assert backend == "native" and got.shape == src.shape and fps == 25.0
    pics = _avi_pictures(path)
    assert len(pics) == n
    for k, jpg in enumerate(pics):
        want = cv2.imdecode(np.frombuffer(jpg, np.uint8), cv2.IMREAD_COLOR)
        assert np.array_equal(got[k], want), (k, int(np.abs(got[k].astype(int) - want.astype(int)).max()))
    # AUTO picks the native decoder for such a file; odd sizes (not multiples of 16 / 2 / 4) and a low quality go through it too
    with decode.MjpegVideo(path) as v:
        assert v.backend == "native"
    rng = np.random.default_rng(3)
    for (w, h, q) in ((322, 182, 95), (53, 37, 60), (1280, 720, 30)):
        frames = [cv2.resize(rng.integers(0, 256, (5, 7, 3), dtype=np.uint8), (w, h), interpolation=cv2.INTER_CUBIC) for _ in range(5)]
        frames = [np.clip(f.astype(int) + rng.integers(-8, 9, f.shape), 0, 255).astype(np.uint8) for f in frames]
        p2 = str(tmp_path / f"odd_{w}x{h}.avi")
        wr = cv2.VideoWriter(p2, cv2.VideoWriter_fourcc(*"MJPG"), 30.0, (w, h))
        wr.set(cv2.VIDEOWRITER_PROP_QUALITY, q)
        for f in frames:
            wr.write(f)
        wr.release()
        dec, *_ = _decode_all(p2, batch_frames=3, backend=decode.ESD_JPEG_NATIVE)
        for k, jpg in enumerate(_avi_pictures(p2)):
            assert np.array_equal(dec[k], cv2.imdecode(np.frombuffer(jpg, np.uint8), cv2.IMREAD_COLOR)), (w, h, k)


@pytest.mark.parametrize("lanes", ["2", "1"])
def test_decode_lanes_respect_a_slow_asynchronous_consumer(mjpeg_clip, monkeypatch, lanes):
    """Consecutive batches decode on two streams of the library's own (ESD_DEC_LANES=1: on the caller's).  The ring slot of batch k
    is rewritten by batch k + 2: that write must wait for whatever the caller enqueued to consume batch k -- here an asynchronous
    copy queued behind a deliberately slow kernel, no host synchronisation anywhere in the loop -- and the caller's stream must
    see finished pictures.  30 small batches; every byte against cv2.imdecode."""
    monkeypatch.setenv("ESD_DEC_LANES", lanes)
    path, src = mjpeg_clip
    n = src.shape[0]
    keep = torch.empty((n,) + src.shape[1:], dtype=torch.uint8, device="cuda:0")
    ballast = torch.randn((2048, 2048), device="cuda:0")
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        with decode.MjpegVideo(path, batch_frames=5, backend=decode.ESD_JPEG_NATIVE) as v:
            pos = 0
            while True:
                b = v.read_batch(0)
                if b is None:
                    break
                for _ in range(3):
                    ballast = (ballast @ ballast).clamp_(-1, 1)     # the consumer is late: ~1 ms of work ahead of the copy
                keep[pos:pos + b.shape[0]].copy_(b, non_blocking=True)
                pos += int(b.shape[0])
            st.synchronize()
    assert pos == n
    got = keep.cpu().numpy()
    for k, jpg in enumerate(_avi_pictures(path)):
        assert np.array_equal(got[k], cv2.imdecode(np.frombuffer(jpg, np.uint8), cv2.IMREAD_COLOR)), k


def test_decoder_decodes_the_right_pictures(mjpeg_clip):
    path, src = mjpeg_clip
    got, fps, backend, n = _decode_all(path, batch_frames=32, backend=decode.ESD_JPEG_GPU_HYBRID)
    assert got.shape == src.shape and n == src.shape[0] and fps == 25.0 and backend in ("hardware", "gpu_hybrid", "default")
    cap = cv2.VideoCapture(path)
    ref = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        ref.append(f)
    ref = np.stack(ref)
    assert ref.shape == got.shape
    mse = np.mean((got.astype(np.float64) - ref.astype(np.float64)) ** 2, axis=(1, 2, 3))
    psnr = 10 * np.log10(255.0 ** 2 / np.maximum(mse, 1e-9))
    assert psnr.min() > 30.0, psnr.min()  # same pictures (garbage would sit below 15 dB); decoder rounding / chroma upsampling apart
    print(f"nvJPEG backend {backend}: PSNR vs cv2/ffmpeg decode min {psnr.min():.1f} dB")


class _Tee:
    """Video source wrapper that keeps a device-side copy of every batch the SceneManager scores: the decoded surface."""

    def __init__(self, src):
        self.src, self.kept = src, []
        self.frame_size, self.frame_rate, self.start_frame, self.pixel_format = src.frame_size, src.frame_rate, src.start_frame, src.pixel_format

    def read_batch(self, n):
        b = self.src.read_batch(n)
        if b is not None:
            self.kept.append(b.clone())  # same stream as the decode and the scoring
        return b

    def frames(self):
        return torch.cat(self.kept).cpu().numpy()


@pytest.mark.parametrize("backend", [decode.ESD_JPEG_NATIVE, decode.ESD_JPEG_GPU_HYBRID])
def test_parity_on_the_decoded_surface(mjpeg_clip, backend):
    """The frames the scoring kernels read are downloaded (a device-side copy made between decode and scoring) and the
    oracle runs on exactly those: integer sums, float64 scores, histograms' differences and cut lists bit-exact."""
    path, _ = mjpeg_clip
    sm = SceneManager(batch_frames=64)
    dets = [ContentDetector(min_scene_len=10), AdaptiveDetector(window_width=2, min_scene_len=10), HistogramDetector(min_scene_len=10)]
    for d in dets:
        sm.add_detector(d)
    with decode.MjpegVideo(path, batch_frames=48, backend=backend) as v:
        tee = _Tee(v)
        n = sm.detect_scenes(tee, collect_scores=True)
        decoded = tee.frames()
    assert n == v.n_frames == decoded.shape[0]
    want_sums, _, _ = co.score_frames(decoded, 256, 144)
    assert np.array_equal(sm.scores["sums3"].astype(np.int64), want_sums)
    oc = P.ContentDetector(min_scene_len=10, backend="closed_form")
    oa = P.AdaptiveDetector(window_width=2, min_scene_len=10, backend="closed_form")
    oh = P.HistogramDetector(min_scene_len=10, backend="closed_form")
    for o, d in ((oc, dets[0]), (oa, dets[1]), (oh, dets[2])):
        cuts, _ = P.detect(decoded, [o], backend="closed_form")
        assert sm.cuts_of(d) == cuts, type(d).__name__
    assert np.array_equal(np.array(oc.scores).view(np.uint64), sm.scores["content_val"].view(np.uint64))
    assert np.array_equal(np.array(oh.diffs)[1:].view(np.uint64), sm.scores["hist_diff"][1:].view(np.uint64))
    assert len(sm.cuts_of(dets[0])) >= 2
    sm.close()
    # how repeatable is the decoder itself?  (reported, not asserted: nvJPEG does not promise bit-identical output across
    # batch compositions; parity above never depends on it)
    again, *_ = _decode_all(path, batch_frames=48, backend=backend)
    other, *_ = _decode_all(path, batch_frames=64, backend=backend)
    if backend == decode.ESD_JPEG_NATIVE:  # the library's own decoder is deterministic whatever the batching
        assert np.array_equal(again, decoded) and np.array_equal(other, decoded)
    print(f"backend {backend} repeatability: same batching differs in {int((again != decoded).sum())} bytes, "
          f"other batching in {int((other != decoded).sum())} of {decoded.size} bytes (max |delta| {int(np.abs(other.astype(int) - decoded.astype(int)).max())})")


def test_ranges_are_independent(mjpeg_clip):
    path, _ = mjpeg_clip
    full, *_ = _decode_all(path, batch_frames=64)
    part, *_ = _decode_all(path, batch_frames=7, first_frame=50, end_frame=93)
    close = lambda a, b: a.shape == b.shape and int(np.abs(a.astype(int) - b.astype(int)).max()) <= 2  # noqa: E731 (decoder rounding)
    assert close(part, full[50:93])
    with decode.MjpegVideo(path, batch_frames=16) as v:
        v.seek(140)
        b = v.read_batch(64)
        assert b.shape[0] == 10 and close(b.cpu().numpy(), full[140:150])
        assert v.read_batch(64) is None
        with pytest.raises(decode.DecodeError):
            v.seek(10 ** 6)


def test_task_surface_decodes_on_the_gpu_and_shards_by_frame_range(mjpeg_clip):
    path, _ = mjpeg_clip
    decoded, *_ = _decode_all(path)
    cfg = {"detector": "content+adaptive", "min_scene_len": 10, "window_width": 2}
    # a separate decode may differ from the scored one by decoder rounding (+-1 in a few bytes), far from any threshold here
    want = P.detect_scenes_dicts(decoded, [P.ContentDetector(min_scene_len=10, backend="closed_form"),
                                           P.AdaptiveDetector(window_width=2, min_scene_len=10, backend="closed_form")], 25.0, backend="closed_form")
    got = asyncio.run(service.ModelManager().detect_scenes(path, cfg))
    assert got == want and len(got["scenes"]) >= 3
    k = torch.cuda.device_count()
    many = asyncio.run(service.ModelManager(devices=[g % k for g in range(3)]).detect_scenes(path, cfg))
    assert many == want
    # gpu_decode=False keeps the reference's host decode loop (cv2.VideoCapture) -- other pixels, same schema
    host = asyncio.run(service.ModelManager().detect_scenes(path, {**cfg, "gpu_decode": False}))
    assert [s["scene_index"] for s in host["scenes"]] == list(range(len(host["scenes"])))


def test_unsupported_and_broken_files(tmp_path, mjpeg_clip):
    _, frames = mjpeg_clip
    other = str(tmp_path / "mpeg4.avi")
    _write_avi(other, frames[:20], fourcc="mp4v")
    assert not decode.is_mjpeg_avi(other)
    with pytest.raises(decode.DecodeError) as e:
        decode.MjpegVideo(other)
    assert e.value.status == -6 and "not Motion-JPEG" in str(e.value)
    got = asyncio.run(service.ModelManager().detect_scenes(other, {}))  # falls back to the host decoder
    assert len(got["scenes"]) >= 1
    junk = str(tmp_path / "junk.avi")
    open(junk, "wb").write(b"RIFF" + b"\x00" * 64)
    with pytest.raises(decode.DecodeError):
        decode.MjpegVideo(junk)
    with pytest.raises(decode.DecodeError):
        decode.MjpegVideo(str(tmp_path / "missing.avi"))
    with pytest.raises(decode.DecodeError):
        decode.MjpegVideo(mjpeg_clip[0], batch_frames=0)
