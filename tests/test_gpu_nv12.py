"""NV12 input (device half): the fused kernel converts decoder surfaces exactly like
cv2.cvtColor(COLOR_YUV2BGR_NV12) for the source pixels its downscale taps read.  Reference = real cv2 conversion followed
by the BGR oracle; bit-exact sums, histograms, float64 scores and cut lists."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
cv2 = pytest.importorskip("cv2")

from eioku_b200 import capi  # noqa: E402
import synthclip as synth
from eioku_b200.detectors import AdaptiveDetector, ContentDetector, HashDetector, HistogramDetector, ThresholdDetector  # noqa: E402
from eioku_b200.scene_manager import BatchVideo, SceneManager, TensorVideo  # noqa: E402
from oracle import c_oracle as co  # noqa: E402
from oracle import psd_cv2 as P  # noqa: E402

DEV = "cuda:0"
ALL = capi.ESD_DET_CONTENT | capi.ESD_DET_ADAPTIVE | capi.ESD_DET_HIST


def nv12_ctx(w, h, dst=None, detectors=ALL, **kw):
    cfg = capi.default_config()
    cfg.detectors = detectors
    cfg.src_width, cfg.src_height = w, h
    cfg.src_format = capi.ESD_FMT_NV12
    if dst is not None:
        cfg.dst_width, cfg.dst_height = dst
    for k, v in kw.items():
        setattr(cfg, k, v)
    return capi.EsdContext(cfg, 0)


def to_bgr(nv12_frames):
    return np.stack([cv2.cvtColor(f, cv2.COLOR_YUV2BGR_NV12) for f in nv12_frames])


@pytest.mark.parametrize("w,h,dst", [(1920, 1080, None), (1280, 720, None), (3840, 2160, None), (642, 362, (256, 144)),
                                     (1920, 1080, (274, 154)), (300, 200, (256, 171)), (34, 18, (17, 9)), (2050, 40, (1024, 20))])
def test_nv12_random_frames_vs_cv2_then_oracle(w, h, dst):
    rng = np.random.default_rng(w * 31 + h)
    n = 5
    nv12 = rng.integers(0, 256, (n, h * 3 // 2, w), dtype=np.uint8)
    nv12[1, :h] = rng.integers(0, 32, (h, w))        # near-black luma: the max(0, Y - 16) clamp
    nv12[2, h:] = rng.integers(0, 2, (h // 2, w)) * 255   # extreme chroma: saturation on every channel
    with nv12_ctx(w, h, dst) as ctx:
        dw, dh = ctx.dst_size
        ctx.push_nv12_tensor(torch.from_numpy(nv12).to(DEV), 0)
        sc = ctx.read_scores(0, n)
        hsv = ctx.debug_last_hsv()
        geo = ctx.geometry
        rows = ctx.touched_rows()
    bgr = to_bgr(nv12)
    sums, hist, last_hsv = co.score_frames(bgr, dw, dh, bins=256)
    assert np.array_equal(sc["sums3"].astype(np.int64), sums)
    assert np.array_equal(sc["hist"], hist)
    assert np.array_equal(hsv, last_hsv)
    # touched rows: Y rows first, then UV rows as rows of the contiguous NV12 frame; bytes per frame well under BGR's
    assert geo.row_bytes == w and np.all(np.diff(rows) > 0) and rows[-1] < h * 3 // 2
    assert geo.compact_frame_bytes == len(rows) * w
    assert geo.alg_bytes_per_frame <= geo.compact_frame_bytes


def test_nv12_layouts_pitch_separate_planes_and_ingest():
    """Pitched surfaces with the UV plane at an aligned height (decoder style), unaligned bases, and host frames through
    the ingest ring all give the numbers of the dense contiguous layout."""
    w, h, n = 1280, 720, 12
    rng = np.random.default_rng(7)
    nv12 = rng.integers(0, 256, (n, h * 3 // 2, w), dtype=np.uint8)
    with nv12_ctx(w, h) as ref:
        ref.push_nv12_tensor(torch.from_numpy(nv12).to(DEV), 0)
        want = ref.read_scores(0, n)
    # decoder-style surface: pitch 1536, UV plane at aligned height 736
    pitch, ah = 1536, 736
    surf = torch.zeros((n, ah * 3 // 2, pitch), dtype=torch.uint8, device=DEV)
    surf[:, :h, :w] = torch.from_numpy(nv12[:, :h]).to(DEV)
    surf[:, ah:ah + h // 2, :w] = torch.from_numpy(nv12[:, h:]).to(DEV)
    with nv12_ctx(w, h) as ctx:
        ctx.push_nv12_device(surf.data_ptr(), surf.data_ptr() + ah * pitch, n, surf.stride(0), pitch, 0,
                             torch.cuda.current_stream().cuda_stream)
        got = ctx.read_scores(0, n)
        for k in want:
            assert np.array_equal(np.nan_to_num(got[k], nan=-1), np.nan_to_num(want[k], nan=-1)), k
    # unaligned base and pitch (row pitch w + 3, first byte at offset 5)
    raw = torch.zeros(5 + n * (h * 3 // 2) * (w + 3) + 64, dtype=torch.uint8, device=DEV)
    view = raw[5:5 + n * (h * 3 // 2) * (w + 3)].view(n, h * 3 // 2, w + 3)
    view[:, :, :w] = torch.from_numpy(nv12).to(DEV)
    with nv12_ctx(w, h) as ctx:
        ctx.push_nv12_tensor(view[:, :, :w], 0)
        got = ctx.read_scores(0, n)
        for k in want:
            assert np.array_equal(np.nan_to_num(got[k], nan=-1), np.nan_to_num(want[k], nan=-1)), k
    # host frames: pageable and pinned through the ring (touched Y + UV rows only), in uneven pushes
    for pinned in (False, True):
        host = torch.from_numpy(nv12.copy())
        if pinned:
            host = host.pin_memory()
        with nv12_ctx(w, h) as ctx:
            ctx.ingest_open(3, 5)
            ctx.ingest_push_nv12_numpy(host.numpy()[:7], 0)
            ctx.ingest_push_nv12_numpy(host.numpy()[7:], 7)
            got = ctx.read_scores(0, n)
            h2d, _ = ctx.ingest_stats()
            assert h2d == n * ctx.geometry.compact_frame_bytes < n * w * h * 3 // 2
            for k in want:
                assert np.array_equal(np.nan_to_num(got[k], nan=-1), np.nan_to_num(want[k], nan=-1)), k
            ctx.ingest_close()
        # host tap gather: threads copy the two luma bytes / two chroma pairs per destination column of the touched rows
        for threads in (1, 5):
            with nv12_ctx(w, h) as ctx:
                ctx.ingest_open(3, 5)
                ctx.ingest_set_gather(threads)
                ctx.ingest_push_nv12_numpy(host.numpy()[:7], 0)
                ctx.ingest_push_nv12_numpy(host.numpy()[7:], 7)
                got = ctx.read_scores(0, n)
                h2d, _ = ctx.ingest_stats()
                assert h2d == n * len(ctx.touched_rows()) * 1024 < n * ctx.geometry.compact_frame_bytes   # 4 * 256 bytes per row
                for k in want:
                    assert np.array_equal(np.nan_to_num(got[k], nan=-1), np.nan_to_num(want[k], nan=-1)), (k, threads)
                ctx.ingest_close()


def test_nv12_clip_through_scene_manager_all_detectors():
    """A synthetic 1080p clip as NV12 through SceneManager (device batches and host batches): every detector's cut list
    and per-frame scores equal PySceneDetect's logic on cv2-converted frames."""
    w, h, n, seed = 1920, 1080, 240, 1002
    sch = synth.build_schedule(seed, n, min_len=20, max_len=70)
    bgr_syn = torch.empty((n, h, w, 3), dtype=torch.uint8, device=DEV)
    synth.fill(bgr_syn, seed, sch.descs)
    nv12 = synth.bgr_to_test_nv12(bgr_syn)
    nv12_np = nv12.cpu().numpy()
    assert np.array_equal(nv12_np, synth.bgr_to_test_nv12(bgr_syn.cpu().numpy()))
    refs = {"content": P.ContentDetector(threshold=27.0, min_scene_len=15), "adaptive": P.AdaptiveDetector(),
            "hist": P.HistogramDetector(), "threshold": P.ThresholdDetector(threshold=12, add_final_scene=True)}
    want = {k: [] for k in refs}
    for k in range(n):
        small = cv2.resize(cv2.cvtColor(nv12_np[k], cv2.COLOR_YUV2BGR_NV12), (256, 144), interpolation=cv2.INTER_LINEAR)
        for name, d in refs.items():
            want[name] += d.process_frame(k, small)
    want["threshold"] += refs["threshold"].post_process(n - 1)
    assert len(want["content"]) >= 3
    for source in ("device", "host"):
        dets = {"content": ContentDetector(threshold=27.0, min_scene_len=15), "adaptive": AdaptiveDetector(),
                "hist": HistogramDetector(), "threshold": ThresholdDetector(threshold=12, add_final_scene=True), "hash": HashDetector()}
        sm = SceneManager(batch_frames=64)
        for d in dets.values():
            sm.add_detector(d)
        frames = nv12 if source == "device" else nv12_np
        if source == "host":
            sm._ingest_threads = 3   # host frames through the NV12 tap gather
        assert sm.detect_scenes(TensorVideo(frames, 30.0, pixel_format="nv12"), collect_scores=True) == n
        for name in refs:
            assert sm.cuts_of(dets[name]) == want[name], (source, name)
        assert np.array_equal(sm.scores["content_val"].view(np.uint64), np.array(refs["content"].scores).view(np.uint64))
        assert np.array_equal(sm.scores["sums3"].astype(np.int64), np.stack(refs["content"].sums))
        hd = np.array(refs["hist"].diffs)
        assert np.array_equal(sm.scores["hist_diff"][1:].view(np.uint64), hd[1:].view(np.uint64))
        assert len(sm.cuts_of(dets["hash"])) >= 1
        sm.close()


def test_nv12_errors():
    cfg = capi.default_config()
    cfg.src_format = capi.ESD_FMT_NV12
    for (w, h, dst) in [(256, 144, (256, 144)), (641, 360, (256, 144)), (640, 361, (256, 144))]:
        cfg.src_width, cfg.src_height = w, h
        cfg.dst_width, cfg.dst_height = dst
        with pytest.raises(capi.EsdError):
            capi.EsdContext(cfg, 0)
    cfg.src_format = 7
    cfg.src_width, cfg.src_height, cfg.dst_width, cfg.dst_height = 640, 360, 0, 0
    with pytest.raises(capi.EsdError):
        capi.EsdContext(cfg, 0)
    with nv12_ctx(640, 360) as ctx:
        with pytest.raises(ValueError):
            ctx.push_nv12_tensor(torch.zeros((2, 360, 640), dtype=torch.uint8, device=DEV), 0)
        with pytest.raises(capi.EsdError):
            ctx.process_frame_host(np.zeros((360, 640, 3), np.uint8), 0, capi.ESD_DET_CONTENT)
    bgr_cfg = capi.default_config()
    bgr_cfg.src_width, bgr_cfg.src_height = 640, 360
    with capi.EsdContext(bgr_cfg, 0) as ctx:
        t = torch.zeros((1, 540, 640), dtype=torch.uint8, device=DEV)
        with pytest.raises(capi.EsdError):
            ctx.push_nv12_device(t.data_ptr(), t.data_ptr() + 360 * 640, 1, 540 * 640, 640, 0)


def test_nv12_clip_vs_committed_cv2_golden():
    """The same kind of clip against the committed golden (tests/golden/make_golden.py nv12: real cv2 conversion + PySceneDetect's
    logic, generated in the build container): integer sums, float64 scores, hist_diff and the three cut lists."""
    import hashlib

    g = load_golden("nv12_c2_1080p_head.npz")
    w, h, n, seed = int(g["width"]), int(g["height"]), int(g["n_frames"]), int(g["seed"])
    sch = synth.build_schedule(seed, n, min_len=20, max_len=70)
    bgr = torch.empty((n, h, w, 3), dtype=torch.uint8, device=DEV)
    synth.fill(bgr, seed, sch.descs)
    nv12 = synth.bgr_to_test_nv12(bgr)
    assert hashlib.sha256(nv12.cpu().numpy().tobytes()).digest() == bytes(g["input_sha256"])
    dets = [ContentDetector(threshold=27.0, min_scene_len=15), AdaptiveDetector(), HistogramDetector()]
    sm = SceneManager(batch_frames=100)
    for d in dets:
        sm.add_detector(d)
    assert sm.detect_scenes(TensorVideo(nv12, 30.0, pixel_format="nv12"), collect_scores=True) == n
    assert np.array_equal(sm.scores["sums3"], g["sums3"])
    assert np.array_equal(sm.scores["content_val"].view(np.uint64), g["content_val"].view(np.uint64))
    assert np.array_equal(sm.scores["hist_diff"][1:].view(np.uint64), g["hist_diff"][1:].view(np.uint64))
    assert sm.cuts_of(dets[0]) == g["cuts_content"].tolist() and len(g["cuts_content"]) >= 3
    assert sm.cuts_of(dets[1]) == g["cuts_adaptive"].tolist()
    assert sm.cuts_of(dets[2]) == g["cuts_hist"].tolist()
    sm.close()


@pytest.mark.parametrize("oy,ouv", [(0, 15), (15, 0), (15, 15), (7, 9), (1, 2), (8, 4), (3, 12)])
def test_nv12_every_plane_misalignment(oy, ouv):
    """memory safety rests on parity over ragged layouts.  Y plane starting at byte residue `oy` and UV
    plane at residue `ouv` (mod 16), an odd row pitch, a frame stride that is not a multiple of 16: the bulk copies are
    rounded to 16-byte boundaries and the misalignments travel in the stage metadata -- results must equal the dense case."""
    w, h, n = 322, 182, 6
    rng = np.random.default_rng(oy * 16 + ouv)
    nv12 = rng.integers(0, 256, (n, h * 3 // 2, w), dtype=np.uint8)
    with nv12_ctx(w, h, (129, 73)) as ctx:
        ctx.push_nv12_tensor(torch.from_numpy(nv12).to(DEV), 0)
        want = ctx.read_scores(0, n)
    pitch = w + 3                                   # odd pitch: every row has another residue
    gap = (ouv - (oy + h * pitch)) % 16             # puts the UV plane's first byte at residue ouv
    fs = oy + h * pitch + gap + (h // 2) * pitch + 5   # frame stride, not a multiple of 16 in general
    raw = torch.zeros(n * fs + 64, dtype=torch.uint8, device=DEV)
    base = raw.data_ptr()
    assert base % 256 == 0
    frames = raw[:n * fs].view(n, fs)
    y = torch.from_numpy(nv12[:, :h]).to(DEV)
    uv = torch.from_numpy(nv12[:, h:]).to(DEV)
    for r in range(h):
        frames[:, oy + r * pitch: oy + r * pitch + w] = y[:, r]
    uv0 = oy + h * pitch + gap
    for r in range(h // 2):
        frames[:, uv0 + r * pitch: uv0 + r * pitch + w] = uv[:, r]
    assert (base + oy) % 16 == oy and (base + uv0) % 16 == ouv
    with nv12_ctx(w, h, (129, 73)) as ctx:
        ctx.push_nv12_device(base + oy, base + uv0, n, fs, pitch, 0, torch.cuda.current_stream().cuda_stream)
        got = ctx.read_scores(0, n)
        hsv = ctx.debug_last_hsv()
    for k in want:
        assert np.array_equal(np.nan_to_num(got[k], nan=-1), np.nan_to_num(want[k], nan=-1)), k
    bgr = to_bgr(nv12)
    sums, hist, last_hsv = co.score_frames(bgr, 129, 73, bins=256)
    assert np.array_equal(got["sums3"].astype(np.int64), sums) and np.array_equal(hsv, last_hsv)
