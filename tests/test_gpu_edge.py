"""GPU edge cases: empty / single-frame / tiny / ragged inputs, maximum widths, context reuse, overflow and
metric export -- all against the oracle, through the C ABI."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from eioku_b200 import capi  # noqa: E402
import synthclip as synth
from eioku_b200.detectors import AdaptiveDetector, ContentDetector, HistogramDetector, StatsManager  # noqa: E402
from eioku_b200.scene_manager import SceneManager, TensorVideo  # noqa: E402
from oracle import c_oracle as co  # noqa: E402
from oracle import closed_form as cf  # noqa: E402
from oracle import psd_cv2 as P  # noqa: E402
from test_gpu_parity import DEV, gpu_clip, make_ctx, oracle_scores, same_f64  # noqa: E402


def _check(frames, dst=None, **kw):
    n, h, w, _ = frames.shape
    dw, dh = dst if dst else (w, h)
    sums, hist, cv, hd = oracle_scores(frames, dw, dh)
    with make_ctx(w, h, (dw, dh), **kw) as ctx:
        ctx.push_tensor(torch.from_numpy(frames).to(DEV), 0)
        sc = ctx.read_scores(0, n)
    assert np.array_equal(sc["sums3"].astype(np.int64), sums)
    assert np.array_equal(sc["hist"], hist)
    assert same_f64(sc["content_val"], cv) and same_f64(sc["hist_diff"], hd)


@pytest.mark.parametrize("w,h", [(1, 1), (2, 3), (17, 1), (1, 40), (255, 9), (5, 5)])
def test_tiny_images(w, h):
    rng = np.random.default_rng(w * 100 + h)
    _check(rng.integers(0, 256, (7, h, w, 3), dtype=np.uint8))


def test_single_and_two_frame_videos():
    rng = np.random.default_rng(1)
    for n in (1, 2):
        fr = rng.integers(0, 256, (n, 90, 160, 3), dtype=np.uint8)
        _check(fr)
        sm = SceneManager()
        sm.add_detector(ContentDetector())
        sm.add_detector(AdaptiveDetector())
        sm.add_detector(HistogramDetector())
        assert sm.detect_scenes(TensorVideo(torch.from_numpy(fr).to(DEV), 25.0)) == n
        assert sm.get_cut_list() == [] and sm.get_scene_list() == [] and sm.get_scene_list(start_in_scene=True) == [(0, n)]
        sm.close()


def test_width_thresholds_of_auto_downscale():
    rng = np.random.default_rng(2)
    for w, h in ((255, 40), (256, 40), (257, 40), (511, 33), (512, 34)):
        f = cf.compute_downscale_factor(w)
        dw, dh = cf.downscaled_size(w, h, f)
        fr = rng.integers(0, 256, (5, h, w, 3), dtype=np.uint8)
        sums, hist, cv, hd = oracle_scores(fr, dw, dh)
        with make_ctx(w, h, None) as ctx:  # dst 0,0 -> the library derives the SceneManager size itself
            assert ctx.dst_size == (dw, dh), (w, ctx.dst_size, (dw, dh))
            ctx.push_tensor(torch.from_numpy(fr).to(DEV), 0)
            sc = ctx.read_scores(0, 5)
        assert np.array_equal(sc["sums3"].astype(np.int64), sums) and np.array_equal(sc["hist"], hist)


def test_maximum_widths():
    rng = np.random.default_rng(3)
    _check(rng.integers(0, 256, (3, 6, 3840, 3), dtype=np.uint8))            # 4K rows, no resize (16 px / thread)
    _check(rng.integers(0, 256, (3, 6, 4096, 3), dtype=np.uint8))            # widest no-resize row
    _check(rng.integers(0, 256, (3, 9, 7680, 3), dtype=np.uint8), (1024, 3))  # 8K wide source, widest resize target
    cfg = capi.default_config()
    cfg.src_width, cfg.src_height, cfg.dst_width, cfg.dst_height = 5000, 4, 5000, 4
    with pytest.raises(capi.EsdError) as e:
        capi.EsdContext(cfg, 0)
    assert "too large" in str(e.value)


def test_every_bin_count_matches_opencv_order():
    rng = np.random.default_rng(4)
    a = rng.integers(0, 256, (6, 72, 128, 3), dtype=np.uint8)
    a[3] = np.clip(a[2].astype(int) + rng.integers(-9, 9, a[2].shape), 0, 255)
    for bins in (1, 2, 3, 5, 17, 100, 255, 256):
        sums, hist, cv, hd = oracle_scores(a, 128, 72, bins=bins)
        with make_ctx(128, 72, None, detectors=capi.ESD_DET_HIST, hist_bins=bins) as ctx:
            ctx.push_tensor(torch.from_numpy(a).to(DEV), 0)
            sc = ctx.read_scores(0, 6)
        assert np.array_equal(sc["hist"], hist), bins
        assert same_f64(sc["hist_diff"], hd), bins
    cfg = capi.default_config()
    cfg.detectors, cfg.src_width, cfg.src_height, cfg.hist_bins = capi.ESD_DET_HIST, 64, 48, 257
    with pytest.raises(capi.EsdError):
        capi.EsdContext(cfg, 0)


def test_context_reuse_after_reset_and_interleaved_contexts():
    w, h, n = 640, 360, 120
    s1 = synth.build_schedule(31, n, min_len=10, max_len=40)
    s2 = synth.build_schedule(32, n, min_len=10, max_len=40)
    c1, c2 = gpu_clip(31, w, h, s1.descs), gpu_clip(32, w, h, s2.descs)

    def fresh(clip):
        with make_ctx(w, h, None) as ctx:
            ctx.push_tensor(clip, 0)
            return ctx.read_scores(0, n), [ctx.get_cuts(d)[0] for d in (1, 2, 4)]

    r1, r2 = fresh(c1), fresh(c2)
    with make_ctx(w, h, None) as ctx:  # one ctx, two videos in a row
        for clip, ref in ((c1, r1), (c2, r2), (c1, r1)):
            ctx.reset()
            ctx.push_tensor(clip[:50], 0)
            ctx.push_tensor(clip[50:], 50)
            got = ctx.read_scores(0, n)
            assert all(np.array_equal(got[k], ref[0][k], equal_nan=True) for k in got)
            assert [ctx.get_cuts(d)[0] for d in (1, 2, 4)] == ref[1]
    a, b = make_ctx(w, h, None), make_ctx(w, h, None)  # two ctxs interleaved on one device
    st2 = torch.cuda.Stream()
    for lo in range(0, n, 30):
        a.push_tensor(c1[lo:lo + 30], lo)
        with torch.cuda.stream(st2):
            b.push_tensor(c2[lo:lo + 30], lo)
    ga, gb = a.read_scores(0, n), b.read_scores(0, n)
    assert all(np.array_equal(ga[k], r1[0][k], equal_nan=True) for k in ga)
    assert all(np.array_equal(gb[k], r2[0][k], equal_nan=True) for k in gb)
    a.close(); b.close()


def test_cut_list_overflow_is_reported():
    w, h, n = 64, 48, 64
    fr = np.zeros((n, h, w, 3), np.uint8)
    fr[1::2] = 255  # every frame is a cut candidate
    with make_ctx(w, h, None, detectors=capi.ESD_DET_CONTENT, content_min_scene_len=0, max_cuts=10) as ctx:
        ctx.push_tensor(torch.from_numpy(fr).to(DEV), 0)
        with pytest.raises(capi.EsdError) as e:
            ctx.get_cuts(capi.ESD_DET_CONTENT)
        assert e.value.status == -6
    with make_ctx(w, h, None, detectors=capi.ESD_DET_CONTENT, content_min_scene_len=0) as ctx:
        ctx.push_tensor(torch.from_numpy(fr).to(DEV), 0)
        assert ctx.get_cuts(capi.ESD_DET_CONTENT)[0] == list(range(1, n))


def test_empty_and_misordered_pushes():
    with make_ctx(64, 48, None) as ctx:
        t = torch.zeros((0, 48, 64, 3), dtype=torch.uint8, device=DEV)
        with pytest.raises(capi.EsdError):
            ctx.push_tensor(t, 0)
        assert ctx.frames_pushed == 0
        assert ctx.read_scores(0, 0) is not None
    sm = SceneManager()
    sm.add_detector(ContentDetector())
    assert sm.detect_scenes(TensorVideo(torch.zeros((0, 48, 64, 3), dtype=torch.uint8, device=DEV))) == 0
    assert sm.get_scene_list(start_in_scene=True) == []


def test_stats_manager_metrics_equal_oracle():
    w, h, n, seed = 256, 144, 80, 9
    sch = synth.build_schedule(seed, n, min_len=10, max_len=30)
    frames = co.synth_frames(seed, w, h, sch.descs)
    stats = StatsManager()
    sm = SceneManager(stats_manager=stats)
    ad = AdaptiveDetector(window_width=2)
    hd = HistogramDetector(bins=64)
    sm.add_detector(ad)
    sm.add_detector(hd)
    sm.detect_scenes(TensorVideo(torch.from_numpy(frames).to(DEV), 30.0))
    o_ad = P.AdaptiveDetector(window_width=2, backend="closed_form")
    o_hd = P.HistogramDetector(bins=64, backend="closed_form")
    P.detect(frames, [o_ad, o_hd], backend="closed_form", auto_downscale=False)
    npx = float(w * h)
    for k in range(1, n):
        cvk, dh_, ds_, dl_ = stats.get_metrics(k, ["content_val", "delta_hue", "delta_sat", "delta_lum"])
        assert cvk == o_ad.scores[k]
        assert [dh_, ds_, dl_] == [np.int64(v) / npx for v in o_ad.sums[k]]
        assert stats.get_metrics(k, ["hist_diff [bins=64]"])[0] == o_hd.diffs[k]
    for t, r in o_ad.ratios.items():
        assert stats.get_metrics(t, ["adaptive_ratio (w=2)"])[0] == r
    assert not stats.metrics_exist(0, ["content_val"]) and not stats.metrics_exist(n - 1, ["adaptive_ratio (w=2)"])
    sm.close()
    # stand-alone detector with a stats manager, CUDA tensors in ragged batches
    st2 = StatsManager()
    det = AdaptiveDetector(window_width=2)
    det.stats_manager = st2
    dev = torch.from_numpy(frames).to(DEV)
    cuts = []
    for lo, hi in ((0, 1), (1, 2), (2, 40), (40, 41), (41, n)):
        cuts += det.process_frames(lo, dev[lo:hi])
    ref = P.AdaptiveDetector(window_width=2, backend="closed_form")
    want, _ = P.detect(frames, [ref], backend="closed_form", auto_downscale=False)
    assert cuts == want
    for t, r in ref.ratios.items():
        assert st2.get_metrics(t, ["adaptive_ratio (w=2)"])[0] == r
    det.close()


def test_ingest_of_cropped_host_views():
    """Host frames that are a crop of a wider/taller array (row pitch > row bytes, unaligned base, frame stride > frame):
    every ingest route (CPU row staging, pinned DMA, host tap gather) must equal the device-resident result."""
    rng = np.random.default_rng(21)
    big = rng.integers(0, 256, (9, 130, 700, 3), dtype=np.uint8)
    view = big[:, 5:125, 13:653]  # 640x120 crop
    assert not view.flags["C_CONTIGUOUS"] and view.strides[2] == 3
    dense = np.ascontiguousarray(view)
    n, h, w, _ = dense.shape
    with make_ctx(w, h, None) as ref:
        ref.push_tensor(torch.from_numpy(dense).to(DEV), 0)
        want = ref.read_scores(0, n)
    pinned_big = torch.from_numpy(big).pin_memory().numpy()
    for src_big in (big, pinned_big):
        src = src_big[:, 5:125, 13:653]
        for threads in (0, 4):
            with make_ctx(w, h, None) as ctx:
                ctx.ingest_open(2, 4)
                ctx.ingest_set_gather(threads)
                ctx.ingest_push_host(src.ctypes.data, n, src.strides[0], src.strides[1], 0)
                got = ctx.read_scores(0, n)
                ctx.ingest_close()
            for k in want:
                assert np.array_equal(got[k], want[k], equal_nan=True), (k, threads)


def test_bad_frame_pointers_are_rejected_before_launch():
    """Pageable host pointers and spans that overrun the device allocation are refused with an error instead of
    faulting inside the kernel."""
    w, h = 320, 180
    with make_ctx(w, h, None) as ctx:
        host = np.zeros((2, h, w, 3), np.uint8)
        with pytest.raises(capi.EsdError) as e:
            ctx.push_device(host.ctypes.data, 2, host.strides[0], host.strides[1], 0)
        assert "pageable" in str(e.value)
        own = torch.empty(3 * h * w * 3 + (1 << 22), dtype=torch.uint8, device=DEV)  # private cudaMalloc-sized block
        torch.cuda.synchronize()
        # the caching allocator must not carve `big` out of a larger block an earlier test left cached: its end has to be
        # the end of a real cudaMalloc allocation
        torch.cuda.empty_cache()
        big = torch.empty(1 << 28, dtype=torch.uint8, device=DEV)  # its own 256 MiB segment
        end = big.data_ptr() + big.numel()
        with pytest.raises(capi.EsdError) as e:
            ctx.push_device(end - 2 * h * w * 3, 3, h * w * 3, w * 3, 0)  # third frame lies past the allocation
        assert "allocation ends" in str(e.value)
        assert ctx.frames_pushed == 0
        ctx.push_device(end - 3 * h * w * 3, 3, h * w * 3, w * 3, 0)      # exactly fits: accepted
        assert ctx.frames_pushed == 3
        pinned = torch.zeros((2, h, w, 3), dtype=torch.uint8).pin_memory()
        ctx.push_device(pinned.data_ptr(), 2, pinned.stride(0), pinned.stride(1), 3)  # pinned host memory is reachable (zero-copy)
        assert ctx.read_scores(3, 2)["sums3"][1].tolist() == [0, 0, 0]
        del own


@pytest.mark.parametrize("w,h,dst", [(322, 182, (129, 73)), (4, 3, None), (5, 2, None), (1280, 36, (256, 7))])
def test_every_base_misalignment_and_odd_pitch(w, h, dst):
    """Frames whose first byte sits at each of the 16 residues mod 16 (15 included), with a row pitch and a frame stride
    that are not multiples of 16 and -- for the tiny geometries -- rows shorter than one 16-byte bulk-copy unit."""
    n = 5
    rng = np.random.default_rng(w * 7 + h)
    fr = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    dw, dh = dst if dst else (w, h)
    sums, hist, cv, hd = oracle_scores(fr, dw, dh)
    pitch = w * 3 + (1 if (w * 3) % 2 == 0 else 2)          # odd
    fs = h * pitch + 7
    dev_fr = torch.from_numpy(fr).to(DEV).view(n, h, w * 3)
    for off in range(16):
        raw = torch.zeros(off + n * fs + 64, dtype=torch.uint8, device=DEV)
        assert raw.data_ptr() % 256 == 0
        frames = raw[off:off + n * fs].view(n, fs)
        for r in range(h):
            frames[:, r * pitch: r * pitch + w * 3] = dev_fr[:, r]
        with make_ctx(w, h, (dw, dh)) as ctx:
            ctx.push_device(raw.data_ptr() + off, n, fs, pitch, 0, torch.cuda.current_stream().cuda_stream)
            sc = ctx.read_scores(0, n)
        assert np.array_equal(sc["sums3"].astype(np.int64), sums), off
        assert np.array_equal(sc["hist"], hist), off
        assert same_f64(sc["content_val"], cv) and same_f64(sc["hist_diff"], hd), off


def test_red_zone_allocator_catches_an_out_of_bounds_write():
    """compute-sanitizer is closed on this pool; the library's own stand-in (csrc/guard_alloc.h, ESD_GUARD=1) surrounds every
    device buffer with guard zones verified at free time.  In a child process: guards active; an undamaged buffer passes; a write
    one byte past the end aborts.  (Run the whole GPU suite under ESD_GUARD=1 to check every kernel: profiles/r02_guard_suite.log.)"""
    import subprocess
    import sys

    code = ("import os, sys; sys.path.insert(0, %r); from eioku_b200 import capi; L = capi.load_library(); "
            "import torch; torch.zeros(1, device='cuda'); r = L.esd_debug_guard_selftest(int(sys.argv[1])); print('selftest', r)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, ESD_GUARD="1")
    ok = subprocess.run([sys.executable, "-c", code % root, "0"], env=env, capture_output=True, text=True, timeout=120)
    assert ok.returncode == 0 and "selftest 1" in ok.stdout, ok.stderr[-500:]
    bad = subprocess.run([sys.executable, "-c", code % root, "1"], env=env, capture_output=True, text=True, timeout=120)
    assert bad.returncode != 0 and "RED ZONE DAMAGED" in bad.stderr, (bad.returncode, bad.stderr[-500:])
    off = subprocess.run([sys.executable, "-c", code % root, "1"], env=dict(os.environ, ESD_GUARD="0"), capture_output=True, text=True, timeout=120)
    assert off.returncode == 0 and "selftest 0" in off.stdout
