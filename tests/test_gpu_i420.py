"""Planar I420 input (VERDICT r1 "a third src_format: what software decoders emit"): the library interleaves
the chroma rows the taps touch (i420_interleave_kernel) and the fused kernel's YUV 4:2:0 variant converts exactly like
cv2.cvtColor(COLOR_YUV2BGR_I420).  Reference = real cv2 conversion followed by the BGR oracle; bit-exact sums, histograms,
previous-frame HSV, float64 scores and cut lists -- and the numbers of the NV12 form of the same samples."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
cv2 = pytest.importorskip("cv2")

from eioku_b200 import capi  # noqa: E402
import synthclip as synth  # noqa: E402
from eioku_b200.detectors import AdaptiveDetector, ContentDetector, HistogramDetector, ThresholdDetector  # noqa: E402
from eioku_b200.scene_manager import SceneManager, TensorVideo  # noqa: E402
from oracle import c_oracle as co  # noqa: E402

DEV = "cuda:0"
ALL = capi.ESD_DET_CONTENT | capi.ESD_DET_ADAPTIVE | capi.ESD_DET_HIST


def yuv_ctx(fmt, w, h, dst=None, detectors=ALL, **kw):
    cfg = capi.default_config()
    cfg.detectors = detectors
    cfg.src_width, cfg.src_height = w, h
    cfg.src_format = fmt
    if dst is not None:
        cfg.dst_width, cfg.dst_height = dst
    for k, v in kw.items():
        setattr(cfg, k, v)
    return capi.EsdContext(cfg, 0)


def same(got, want):
    for k in want:
        assert np.array_equal(np.nan_to_num(got[k], nan=-1), np.nan_to_num(want[k], nan=-1)), k


@pytest.mark.parametrize("w,h,dst", [(1920, 1080, None), (1280, 720, None), (3840, 2160, None), (642, 362, (256, 144)),
                                     (1920, 1080, (274, 154)), (300, 200, (256, 171)), (34, 18, (17, 9)), (2050, 40, (1024, 20))])
def test_i420_random_frames_vs_cv2_then_oracle(w, h, dst):
    rng = np.random.default_rng(w * 17 + h)
    n = 5
    nv12 = rng.integers(0, 256, (n, h * 3 // 2, w), dtype=np.uint8)
    nv12[1, :h] = rng.integers(0, 32, (h, w))
    nv12[2, h:] = rng.integers(0, 2, (h // 2, w)) * 255
    i420 = synth.nv12_to_i420(nv12)
    with yuv_ctx(capi.ESD_FMT_I420, w, h, dst) as ctx:
        dw, dh = ctx.dst_size
        ctx.push_nv12_tensor(torch.from_numpy(i420).to(DEV), 0)   # [N, H*3/2, W]: the contiguous planar frame on an I420 context
        sc = ctx.read_scores(0, n)
        hsv = ctx.debug_last_hsv()
        launches = ctx.kernel_launches
    assert launches >= 2   # the chroma repack and the fused kernel, at least
    bgr = np.stack([cv2.cvtColor(f, cv2.COLOR_YUV2BGR_I420) for f in i420])
    sums, hist, last_hsv = co.score_frames(bgr, dw, dh, bins=256)
    assert np.array_equal(sc["sums3"].astype(np.int64), sums)
    assert np.array_equal(sc["hist"], hist)
    assert np.array_equal(hsv, last_hsv)
    with yuv_ctx(capi.ESD_FMT_NV12, w, h, dst) as ctx:
        ctx.push_nv12_tensor(torch.from_numpy(nv12).to(DEV), 0)
        same(sc, ctx.read_scores(0, n))


def test_i420_separate_planes_pitches_misaligned_bases_and_batches():
    """An AVFrame-style hand-off: three plane pointers, their own pitches, bases at odd addresses, several pushes (the
    previous-frame state crosses pushes, the repack buffer is reused and grown)."""
    w, h, n = 1280, 720, 11
    rng = np.random.default_rng(11)
    nv12 = rng.integers(0, 256, (n, h * 3 // 2, w), dtype=np.uint8)
    with yuv_ctx(capi.ESD_FMT_NV12, w, h) as ref:
        ref.push_nv12_tensor(torch.from_numpy(nv12).to(DEV), 0)
        want = ref.read_scores(0, n)
    y = torch.from_numpy(nv12[:, :h]).to(DEV)
    u = torch.from_numpy(np.ascontiguousarray(nv12[:, h:, 0::2])).to(DEV)
    v = torch.from_numpy(np.ascontiguousarray(nv12[:, h:, 1::2])).to(DEV)
    for oy, ou, ov, py, puv in ((0, 0, 0, 1280, 640), (3, 5, 9, 1283, 647), (16, 32, 48, 1536, 768), (1, 2, 15, 1281, 641)):
        fs = ((oy + h * py + ou + (h // 2) * puv + ov + (h // 2) * puv + 63) // 64) * 64 + 17
        raw = torch.zeros(n * fs + 64, dtype=torch.uint8, device=DEV)
        fr = raw[:n * fs].view(n, fs)
        y0, u0 = oy, oy + h * py + ou
        v0 = u0 + (h // 2) * puv + ov
        fr[:, y0:y0 + h * py].view(n, h, py)[:, :, :w] = y
        fr[:, u0:u0 + (h // 2) * puv].view(n, h // 2, puv)[:, :, :w // 2] = u
        fr[:, v0:v0 + (h // 2) * puv].view(n, h // 2, puv)[:, :, :w // 2] = v
        base = raw.data_ptr()
        st = torch.cuda.current_stream().cuda_stream
        with yuv_ctx(capi.ESD_FMT_I420, w, h) as ctx:
            for a, b in ((0, 2), (2, 9), (9, 11)):
                ctx.push_i420_device(base + a * fs + y0, base + a * fs + u0, base + a * fs + v0, b - a, fs, py, puv, a, st)
            same(ctx.read_scores(0, n), want)


def test_i420_host_frames_through_the_ingest_ring():
    """Host frames (what ffmpeg / PyAV hand over as yuv420p): pinned rows by DMA, pageable rows by memcpy, and the host tap
    gather -- every path lands on the numbers of the NV12 form and moves only the touched rows / taps."""
    w, h, n = 1280, 720, 12
    rng = np.random.default_rng(5)
    nv12 = rng.integers(0, 256, (n, h * 3 // 2, w), dtype=np.uint8)
    i420 = synth.nv12_to_i420(nv12)
    with yuv_ctx(capi.ESD_FMT_NV12, w, h) as ref:
        ref.push_nv12_tensor(torch.from_numpy(nv12).to(DEV), 0)
        want = ref.read_scores(0, n)
    for pinned in (False, True):
        host = torch.from_numpy(i420.copy())
        if pinned:
            host = host.pin_memory()
        with yuv_ctx(capi.ESD_FMT_I420, w, h) as ctx:
            ctx.ingest_open(3, 5)
            ctx.ingest_push_nv12_numpy(host.numpy()[:7], 0)
            ctx.ingest_push_nv12_numpy(host.numpy()[7:], 7)
            got = ctx.read_scores(0, n)
            h2d, _ = ctx.ingest_stats()
            assert h2d == n * ctx.geometry.compact_frame_bytes < n * w * h * 3 // 2
            same(got, want)
            ctx.ingest_close()
        for threads in (1, 5):
            with yuv_ctx(capi.ESD_FMT_I420, w, h) as ctx:
                ctx.ingest_open(3, 5)
                ctx.ingest_set_gather(threads)
                ctx.ingest_push_nv12_numpy(host.numpy()[:7], 0)
                ctx.ingest_push_nv12_numpy(host.numpy()[7:], 7)
                got = ctx.read_scores(0, n)
                h2d, _ = ctx.ingest_stats()
                assert h2d == n * len(ctx.touched_rows()) * 1024 < n * ctx.geometry.compact_frame_bytes
                same(got, want)
                ctx.ingest_close()


def test_i420_clip_through_scene_manager_all_detectors():
    """The synthetic 1080p clip as I420 through SceneManager, device batches and host batches: every detector's cut list and
    scores equal those of the NV12 form (itself pinned to cv2 + the oracle in test_gpu_nv12.py)."""
    n = 200
    sch = synth.build_schedule(1002, n, min_len=20, max_len=70)
    bgr = torch.empty((n, 1080, 1920, 3), dtype=torch.uint8, device=DEV)
    synth.fill(bgr, 1002, sch.descs)
    nv12 = synth.bgr_to_test_nv12(bgr)
    i420 = synth.nv12_to_i420(nv12)
    del bgr

    def run(frames, fmt, threads=0):
        dets = [ContentDetector(), AdaptiveDetector(), HistogramDetector(), ThresholdDetector(threshold=12, add_final_scene=True)]
        sm = SceneManager(batch_frames=64)
        for det in dets:
            sm.add_detector(det)
        sm._ingest_threads = threads
        assert sm.detect_scenes(TensorVideo(frames, 30.0, pixel_format=fmt), collect_scores=True) == n
        cuts = [sm.cuts_of(d) for d in dets]
        scores = {k: np.nan_to_num(np.asarray(v), nan=-1) for k, v in sm.scores.items()}
        sm.close()
        return cuts, scores

    want_cuts, want_scores = run(nv12, "nv12")
    for frames, threads in ((i420, 0), (i420.cpu().numpy(), 0), (i420.cpu().numpy(), 3)):
        cuts, scores = run(frames, "i420", threads)
        assert cuts == want_cuts and len(cuts[0]) >= 2
        for k in want_scores:
            assert np.array_equal(scores[k], want_scores[k]), k


def test_i420_errors():
    with yuv_ctx(capi.ESD_FMT_I420, 640, 360) as ctx:
        t = torch.zeros((2, 540, 640), dtype=torch.uint8, device=DEV)
        with pytest.raises(capi.EsdError):   # the NV12 entry point on a planar context
            ctx.push_nv12_device(t.data_ptr(), t.data_ptr() + 360 * 640, 2, 540 * 640, 640, 0)
        with pytest.raises(capi.EsdError):   # contiguous planar frames need an even pitch
            ctx.push_device(t.data_ptr(), 1, 540 * 641, 641, 0)
        with pytest.raises(capi.EsdError):   # the V plane is not device memory the library can reach
            ctx.push_i420_device(t.data_ptr(), t.data_ptr() + 360 * 640, t.data_ptr() + (1 << 34), 1, 540 * 640, 640, 320, 0)
        ctx.push_nv12_tensor(t, 0)   # and the context is still usable
        assert ctx.frames_pushed == 2
    with yuv_ctx(capi.ESD_FMT_NV12, 640, 360) as ctx:
        t = torch.zeros((1, 540, 640), dtype=torch.uint8, device=DEV)
        with pytest.raises(capi.EsdError):
            ctx.push_i420_device(t.data_ptr(), t.data_ptr() + 360 * 640, t.data_ptr() + 450 * 640, 1, 540 * 640, 640, 320, 0)
