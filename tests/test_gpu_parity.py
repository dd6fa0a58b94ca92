"""GPU parity tests proper: the CUDA path (through the C ABI, via ctypes) against the oracle and
the committed cv2-derived golden vectors.  Bit-exact for integer sums, histograms, float64 scores
and cut lists.  Run on the B200 box: python -m pytest tests -m gpu."""
import hashlib
import os

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from eioku_b200 import capi  # noqa: E402
import synthclip as synth
from eioku_b200.detectors import (AdaptiveDetector, ContentDetector, FlashFilter, HistogramDetector,  # noqa: E402
                                  ThresholdDetector)
from eioku_b200.scene_manager import SceneManager, TensorVideo  # noqa: E402
from eioku_b200.service import detect_scenes_frames  # noqa: E402
from oracle import c_oracle as co  # noqa: E402
from oracle import closed_form as cf  # noqa: E402
from oracle import psd_cv2 as P  # noqa: E402

DEV = "cuda:0"
ALL = capi.ESD_DET_CONTENT | capi.ESD_DET_ADAPTIVE | capi.ESD_DET_HIST


def make_ctx(w, h, dst=None, detectors=ALL, **kw):
    cfg = capi.default_config()
    cfg.detectors = detectors
    cfg.src_width, cfg.src_height = w, h
    if dst is not None:
        cfg.dst_width, cfg.dst_height = dst
    for k, v in kw.items():
        setattr(cfg, k, v)
    return capi.EsdContext(cfg, 0)


def gpu_clip(seed, w, h, descs):
    out = torch.empty((len(descs), h, w, 3), dtype=torch.uint8, device=DEV)
    synth.fill(out, seed, descs)
    return out


def oracle_scores(frames_np, dw, dh, bins=256):
    """Closed-form C oracle + float stage restated in oracle/psd_cv2.py (no cv2 needed)."""
    sums, hist, _ = co.score_frames(frames_np, dw, dh, bins=bins)
    npx = float(dw * dh)
    cv = np.zeros(len(frames_np))
    for i in range(1, len(frames_np)):
        comps = [np.int64(sums[i, c]) / npx for c in range(3)] + [0.0]
        cv[i] = sum(c * w for c, w in zip(comps, (1.0, 1.0, 1.0, 0.0))) / sum(abs(w) for w in (1.0, 1.0, 1.0, 0.0))
    hd = np.full(len(frames_np), np.nan)
    hn = [P.normalize_l2_f32(h) for h in hist]
    for i in range(1, len(frames_np)):
        hd[i] = P.compare_hist_correl(hn[i - 1], hn[i])
    return sums, hist, cv, hd


def same_f64(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64)[~np.isnan(a)], b.view(np.uint64)[~np.isnan(b)]) and \
        np.array_equal(np.isnan(a), np.isnan(b))


# ------------------------------------------------------------------------------------------------
def test_synth_gpu_equals_cpu_twin():
    for (w, h, n, seed) in [(320, 180, 400, 11), (1280, 720, 12, 1001)]:
        sch = synth.build_schedule(seed, n, min_len=20, max_len=60) if w == 320 else synth.build_schedule(seed, n)
        g = gpu_clip(seed, w, h, sch.descs).cpu().numpy()
        c = co.synth_frames(seed, w, h, sch.descs)
        assert np.array_equal(g, c)


@pytest.mark.parametrize("name", ["stage_1920x1080_to_256x144", "stage_1920x1080_to_274x154", "stage_1280x720_to_256x144",
                                  "stage_3840x2160_to_256x144", "stage_854x480_to_285x160", "stage_300x200_to_256x171"])
def test_stage_golden_per_pixel(name):
    """resize + BGR2HSV per pixel and Y histogram against cv2 outputs frozen in tests/golden."""
    g = load_golden(name + ".npz")
    src, dst = name.split("_")[1], name.split("_")[3]
    w, h = map(int, src.split("x"))
    dw, dh = map(int, dst.split("x"))
    img = np.random.default_rng(int(g["seed"])).integers(0, 256, (h, w, 3), dtype=np.uint8)
    small = g["small"]
    with make_ctx(w, h, (dw, dh)) as ctx:
        ctx.push_tensor(torch.from_numpy(img).to(DEV), 0)
        hsv = ctx.debug_last_hsv()
        sc = ctx.read_scores(0, 1)
    assert hashlib.sha256(hsv.tobytes()).digest() == bytes(g["hsv_sha256"])
    assert np.array_equal(hsv, cf.bgr2hsv_u8(small))
    assert np.array_equal(sc["hist"][0], g["y_hist"])
    assert sc["sums3"][0].tolist() == [0, 0, 0] and sc["content_val"][0] == 0.0 and np.isnan(sc["hist_diff"][0])


CLIPS = ["c1_720p", "c2_1080p_head", "c4_4k_head"]


@pytest.mark.parametrize("name", CLIPS)
def test_clip_golden_scores_and_cuts(name):
    """BASELINE configs 1/2/4 (seeded synthetic clips): every per-frame number and every cut list equals
    what PySceneDetect's logic on real cv2 produced (tests/golden/make_golden.py)."""
    g = load_golden(f"clip_{name}.npz")
    w, h, n, seed = int(g["width"]), int(g["height"]), int(g["n_frames"]), int(g["seed"])
    sch = synth.build_schedule(seed, n)
    batch = {1280: 257, 1920: 128, 3840: 48}[w]
    dets = [ContentDetector(threshold=27.0, min_scene_len=15), AdaptiveDetector(adaptive_threshold=3.0, window_width=2),
            HistogramDetector(threshold=0.05, bins=256, min_scene_len=15),
            ThresholdDetector(threshold=12, min_scene_len=15, add_final_scene=True)]
    sm = SceneManager(batch_frames=batch)
    for d in dets:
        sm.add_detector(d)

    def batches():
        for a in range(0, n, batch):
            yield gpu_clip(seed, w, h, sch.descs[a:a + batch])

    from eioku_b200.scene_manager import BatchVideo
    got_n = sm.detect_scenes(BatchVideo(batches(), (w, h), 30.0), collect_scores=True)
    assert got_n == n
    sc = sm.scores
    assert same_f64(sc["average_rgb"], g["average_rgb"])
    assert sm.cuts_of(dets[3]) == g["cuts_threshold"].tolist()
    assert np.array_equal(sc["sums3"], g["sums3"])
    assert same_f64(sc["content_val"], g["content_val"])
    assert same_f64(sc["adaptive_val"], g["content_val"])
    assert same_f64(sc["adaptive_ratio"], g["adaptive_ratio"])
    assert same_f64(sc["hist_diff"], g["hist_diff"])
    assert np.array_equal(sc["hist"][0], g["hist_first"]) and np.array_equal(sc["hist"][-1], g["hist_last"])
    assert hashlib.sha256(np.ascontiguousarray(sc["hist"]).tobytes()).digest() == bytes(g["hist_sha256"])
    assert sm.cuts_of(dets[0]) == g["cuts_content"].tolist()
    assert sm.cuts_of(dets[1]) == g["cuts_adaptive"].tolist()
    assert sm.cuts_of(dets[2]) == g["cuts_hist"].tolist()
    sm.close()
    # luma_only and the legacy SUPPRESS filter, same frames
    for det, key in ((ContentDetector(luma_only=True), "cuts_content_luma"),
                     (ContentDetector(filter_mode=FlashFilter.Mode.SUPPRESS), "cuts_content_suppress")):
        sm2 = SceneManager(batch_frames=batch)
        sm2.add_detector(det)
        sm2.detect_scenes(BatchVideo(batches(), (w, h), 30.0), collect_scores=True)
        assert sm2.cuts_of(det) == g[key].tolist()
        if key == "cuts_content_luma":
            assert same_f64(sm2.scores["content_val"], g["luma_val"])
        sm2.close()


def test_int_downscale_mode_golden():
    g = load_golden("clip_c2_1080p_int_head.npz")
    n, seed = int(g["n_frames"]), int(g["seed"])
    sch = synth.build_schedule(seed, n)
    sm = SceneManager(batch_frames=100, downscale_mode="int")
    det = ContentDetector()
    sm.add_detector(det)
    sm.detect_scenes(TensorVideo(gpu_clip(seed, 1920, 1080, sch.descs), 30.0), collect_scores=True)
    assert (sm._ctx.geometry.dst_width, sm._ctx.geometry.dst_height) == tuple(g["dst"].tolist()) == (274, 154)
    assert np.array_equal(sm.scores["sums3"], g["sums3"])
    assert same_f64(sm.scores["content_val"], g["content_val"])
    assert sm.cuts_of(det) == g["cuts_content"].tolist()
    sm.close()


@pytest.mark.parametrize("w,h,dst", [(101, 37, None), (333, 200, (256, 154)), (640, 360, (256, 144)), (257, 64, (256, 64)),
                                     (64, 48, (64, 48)), (1000, 30, (500, 15)), (513, 77, (171, 26)), (2000, 16, (2000, 16)),
                                     (2001, 9, (2001, 9)), (1366, 11, (1366, 11)), (1922, 7, (1922, 7)), (1024, 5, (1024, 5))])
def test_random_sizes_vs_oracle(w, h, dst):
    """Ragged / odd / unaligned geometries (row bytes not a multiple of 16, 2x area path, no-resize)."""
    rng = np.random.default_rng(w * 31 + h)
    n = 9
    frames = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    frames[4] = frames[3]  # duplicate frame -> zero deltas
    frames[6] = 0
    frames[7] = 255
    dw, dh = dst if dst else (w, h)
    sums, hist, cv, hd = oracle_scores(frames, dw, dh)
    with make_ctx(w, h, (dw, dh)) as ctx:
        ctx.push_tensor(torch.from_numpy(frames).to(DEV), 0)
        sc = ctx.read_scores(0, n)
        last = ctx.debug_last_hsv()
    assert np.array_equal(sc["sums3"].astype(np.int64), sums)
    assert sums[4].tolist() == [0, 0, 0]
    assert np.array_equal(sc["hist"], hist)
    assert same_f64(sc["content_val"], cv)
    assert same_f64(sc["hist_diff"], hd)
    assert np.array_equal(last, co.bgr2hsv(co.resize_linear(frames[-1], dw, dh)))


def test_pitched_and_strided_frames():
    """Row pitch > row bytes and frame stride > frame bytes, unaligned base pointer."""
    w, h, n = 300, 120, 6
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    pitch = w * 3 + 13
    fstride = pitch * h + 7
    buf = torch.zeros(n * fstride + 64, dtype=torch.uint8, device=DEV)
    base = 5
    host = np.zeros(n * fstride + 64, np.uint8)
    for f in range(n):
        for y in range(h):
            o = base + f * fstride + y * pitch
            host[o:o + w * 3] = frames[f, y].ravel()
    buf.copy_(torch.from_numpy(host))
    dw, dh = 150, 60
    sums, hist, cv, hd = oracle_scores(frames, dw, dh)
    with make_ctx(w, h, (dw, dh)) as ctx:
        ctx.push_device(buf.data_ptr() + base, n, fstride, pitch, 0, torch.cuda.current_stream().cuda_stream)
        sc = ctx.read_scores(0, n)
    assert np.array_equal(sc["sums3"].astype(np.int64), sums)
    assert np.array_equal(sc["hist"], hist)


def test_batching_and_tuning_invariance():
    """One batch, many ragged batches, frame by frame, every split mode / group size: same bits."""
    w, h, n, seed = 640, 360, 150, 77
    sch = synth.build_schedule(seed, n, min_len=10, max_len=40)
    clip = gpu_clip(seed, w, h, sch.descs)
    ref = None
    variants = [dict(), dict(split_mode=capi.ESD_SPLIT_CHUNKS), dict(rows_per_group=1), dict(rows_per_group=7, pipeline_stages=2),
                dict(rows_per_group=16, pipeline_stages=8, ctas_per_sm=1), dict(split_mode=capi.ESD_SPLIT_CHUNKS, rows_per_group=3)]
    for vi, kw in enumerate(variants):
        for sizes in ([n], [1] * 5 + [2, 3, 50, 1, 89], [37] * 4 + [2]):
            if vi > 0 and sizes != [n] and vi != 3:
                continue
            with make_ctx(w, h, None, **kw) as ctx:
                pos = 0
                for s in sizes:
                    ctx.push_tensor(clip[pos:pos + s], 1000 + pos)
                    pos += s
                assert pos == n
                sc = ctx.read_scores(1000, n)
                cuts = [ctx.get_cuts(d)[0] for d in (1, 2, 4)]
            if ref is None:
                ref = (sc, cuts)
                fr = clip.cpu().numpy()
                sums, hist, cv, hd = oracle_scores(fr, 256, 144)
                assert np.array_equal(sc["sums3"].astype(np.int64), sums) and np.array_equal(sc["hist"], hist)
                assert same_f64(sc["content_val"], cv) and same_f64(sc["hist_diff"], hd)
                assert len(cuts[0]) >= 2
            else:
                for k in ref[0]:
                    assert same_f64(sc[k], ref[0][k]) if sc[k].dtype == np.float64 else np.array_equal(sc[k], ref[0][k]), (kw, sizes, k)
                assert cuts == ref[1], (kw, sizes)


def test_plugin_surface_frame_by_frame_vs_oracle():
    """process_frame(frame_num, frame_img) with numpy frames, exactly as PySceneDetect calls it."""
    w, h, n, seed = 256, 144, 220, 5
    sch = synth.build_schedule(seed, n, min_len=12, max_len=50, noise_amp=3)
    frames = co.synth_frames(seed, w, h, sch.descs)
    pairs = [(ContentDetector(threshold=20.0, min_scene_len=8), P.ContentDetector(threshold=20.0, min_scene_len=8, backend="closed_form")),
             (ContentDetector(threshold=20.0, min_scene_len=8, filter_mode=FlashFilter.Mode.SUPPRESS),
              P.ContentDetector(threshold=20.0, min_scene_len=8, filter_mode=P.FILTER_SUPPRESS, backend="closed_form")),
             (ContentDetector(weights=ContentDetector.Components(0.3, 0.7, 1.9, 0.0), min_scene_len=0),
              P.ContentDetector(weights=P.Components(0.3, 0.7, 1.9, 0.0), min_scene_len=0, backend="closed_form")),
             (AdaptiveDetector(adaptive_threshold=2.0, window_width=3, min_content_val=8.0, min_scene_len=5),
              P.AdaptiveDetector(adaptive_threshold=2.0, window_width=3, min_content_val=8.0, min_scene_len=5, backend="closed_form")),
             (HistogramDetector(threshold=0.01, bins=100, min_scene_len=4), P.HistogramDetector(threshold=0.01, bins=100, min_scene_len=4, backend="closed_form")),
             (ThresholdDetector(threshold=130, min_scene_len=3, fade_bias=0.4, method=ThresholdDetector.Method.CEILING),
              P.ThresholdDetector(threshold=130, min_scene_len=3, fade_bias=0.4, method=P.ThresholdDetector.CEILING))]
    for mine, theirs in pairs:
        got, want = [], []
        for k in range(n):
            a = mine.process_frame(k, frames[k])
            b = theirs.process_frame(k, frames[k])
            assert a == b, (type(mine).__name__, k, a, b)
            got += a
            want += b
        assert mine.post_process(n - 1) == []
        assert len(want) >= 1, type(mine).__name__
        mine.close()
    assert AdaptiveDetector(window_width=4).event_buffer_length == 4
    assert ContentDetector().get_metrics() == ["content_val", "delta_hue", "delta_sat", "delta_lum", "delta_edges"]
    assert HistogramDetector(bins=64).get_metrics() == ["hist_diff [bins=64]"]


def test_service_dict_vs_oracle():
    w, h, n, seed = 1280, 720, 400, 1001
    sch = synth.build_schedule(seed, n)
    frames_np = co.synth_frames(seed, w, h, sch.descs)
    for cfg, mk in (({}, lambda: [P.ContentDetector(backend="closed_form")]),
                    ({"detector": "adaptive", "window_width": 2}, lambda: [P.AdaptiveDetector(backend="closed_form")]),
                    ({"detector": "content+hist", "threshold": 30.0, "hist_threshold": 0.02, "min_scene_len": 10},
                     lambda: [P.ContentDetector(threshold=30.0, min_scene_len=10, backend="closed_form"),
                              P.HistogramDetector(threshold=0.02, min_scene_len=10, backend="closed_form")])):
        want = P.detect_scenes_dicts(list(frames_np), mk(), 30.0, backend="closed_form")
        got_dev = detect_scenes_frames(torch.from_numpy(frames_np).to(DEV), cfg, fps=30.0, batch_frames=96)
        got_host = detect_scenes_frames(frames_np, cfg, fps=30.0, batch_frames=96)  # through the pinned ingest ring
        got_gather = detect_scenes_frames(frames_np, dict(cfg, ingest_threads=3), fps=30.0, batch_frames=96)  # host tap gather
        assert got_dev == want and got_host == want and got_gather == want
        assert all(s["duration_ms"] > 0 and s["end_ms"] >= s["start_ms"] >= 0 for s in got_dev["scenes"])
        assert [s["scene_index"] for s in got_dev["scenes"]] == list(range(len(got_dev["scenes"])))


def test_ingest_ring_touched_rows_only():
    w, h, n = 1920, 1080, 40
    sch = synth.build_schedule(1002, n)
    dev = gpu_clip(1002, w, h, sch.descs)
    host = dev.cpu()
    pinned = host.pin_memory()
    with make_ctx(w, h, None) as ref:
        ref.push_tensor(dev, 0)
        want = ref.read_scores(0, n)
        geo = ref.geometry
    assert geo.n_touched_rows == 288 and geo.alg_bytes_per_frame == 1658880
    for src in (host.numpy(), pinned.numpy()):
        for threads, per_frame in ((0, 1658880), (5, 288 * 1536)):
            with make_ctx(w, h, None) as ctx:
                ctx.ingest_open(3, 16)
                ctx.ingest_set_gather(threads)  # 0: DMA of touched rows; >0: host threads gather the tap bytes only
                ctx.ingest_push_numpy(src[:25], 0)
                ctx.ingest_push_numpy(src[25:], 25)
                got = ctx.read_scores(0, n)
                nbytes, _ = ctx.ingest_stats()
                ctx.ingest_close()
            assert nbytes == n * per_frame
            for k in want:
                assert np.array_equal(got[k], want[k], equal_nan=True), (k, threads)
    # gather on ragged geometry: unaligned rows, clamped last column, odd pitch
    rng = np.random.default_rng(77)
    for (gw, gh, dst) in ((333, 200, (256, 154)), (257, 64, (256, 64)), (1000, 30, (500, 15)), (1001, 31, (143, 5)), (700, 33, (100, 11))):
        fr = rng.integers(0, 256, (11, gh, gw, 3), dtype=np.uint8)
        with make_ctx(gw, gh, dst) as ref:
            ref.push_tensor(torch.from_numpy(fr).to(DEV), 0)
            w2 = ref.read_scores(0, 11)
        with make_ctx(gw, gh, dst) as ctx:
            ctx.ingest_open(2, 4)
            ctx.ingest_set_gather(3)
            ctx.ingest_push_numpy(fr, 0)
            g2 = ctx.read_scores(0, 11)
            ctx.ingest_close()
        for k in w2:
            assert np.array_equal(g2[k], w2[k], equal_nan=True), (gw, k)


def test_frame_range_shards_and_global_decision():
    """Config-3 style: 4 shards with a window_width+1 halo scored independently, one decision pass."""
    from eioku_b200 import sharding

    w, h, n, seed = 640, 360, 600, 1003
    sch = synth.build_schedule(seed, n, min_len=20, max_len=90)
    clip = gpu_clip(seed, w, h, sch.descs)
    ww = 2
    with make_ctx(w, h, None) as ctx:
        ctx.push_tensor(clip, 0)
        want = ctx.read_scores(0, n)
        want_cuts = {d: ctx.get_cuts(d)[0] for d in (1, 2, 4)}
        shards = sharding.frame_range_shards(n, 4, ww)
        parts = []
        for sh in shards:
            with make_ctx(w, h, None) as sctx:
                sctx.push_tensor(clip[sh.load_start:sh.load_end], sh.load_start)
                parts.append(sharding.owned_slice(sctx.read_scores(sh.load_start, sh.load_end - sh.load_start), sh))
        merged = sharding.fix_video_start(sharding.merge_owned(parts, shards))
        for k in ("sums3", "content_val", "adaptive_val", "hist_diff"):
            assert np.array_equal(merged[k], want[k], equal_nan=True), k
        # the ratio needs the full halo: every owned frame except the video's first/last w has one
        assert np.array_equal(merged["adaptive_ratio"], want["adaptive_ratio"], equal_nan=True)
        cuts_c, _ = ctx.decide_arrays(capi.ESD_DET_CONTENT, 0, merged["content_val"])
        cuts_a, ratio = ctx.decide_arrays(capi.ESD_DET_ADAPTIVE, 0, merged["adaptive_val"])
        cuts_h, _ = ctx.decide_arrays(capi.ESD_DET_HIST, 0, merged["hist_diff"])
    assert cuts_c == want_cuts[1] and cuts_a == want_cuts[2] and cuts_h == want_cuts[4]
    assert np.array_equal(ratio, want["adaptive_ratio"], equal_nan=True)
    assert len(cuts_c) >= 3


def test_filter_vectors_on_device():
    """FlashFilter MERGE/SUPPRESS state machine unit vectors through esd_decide_arrays."""
    import json

    from conftest import GOLDEN
    vecs = json.load(open(os.path.join(GOLDEN, "filter_vectors.json")))["vectors"]
    ctxs = {}
    for v in vecs:
        key = (v["length"], v["mode"])
        if key not in ctxs:
            ctxs[key] = make_ctx(64, 48, None, detectors=capi.ESD_DET_CONTENT, content_min_scene_len=v["length"],
                                 content_filter_mode=v["mode"], content_threshold=27.0)
        scores = np.where(np.array(v["above"]) > 0, 30.0, 1.0)
        cuts, _ = ctxs[key].decide_arrays(capi.ESD_DET_CONTENT, v["start"], scores)
        assert cuts == v["cuts"], v
    for c in ctxs.values():
        c.close()


def test_errors_are_loud():
    with pytest.raises(ValueError):
        ContentDetector(kernel_size=4)
    with pytest.raises(ValueError):
        AdaptiveDetector(window_width=0)
    with pytest.raises(ValueError):
        HistogramDetector().process_frame(0, np.zeros((10, 10, 3), np.float32))
    with pytest.raises(ValueError):
        HistogramDetector().process_frame(0, np.zeros((10, 10, 4), np.uint8))
    with make_ctx(64, 48, None) as ctx:
        fr = torch.zeros((2, 48, 64, 3), dtype=torch.uint8, device=DEV)
        ctx.push_tensor(fr, 10)
        with pytest.raises(capi.EsdError):
            ctx.push_tensor(fr, 20)  # non-sequential frame number
        with pytest.raises(capi.EsdError):
            ctx.read_scores(0, 5)
        with pytest.raises(ValueError):
            ctx.push_tensor(torch.zeros((1, 50, 64, 3), dtype=torch.uint8, device=DEV), 12)
    cfg = capi.default_config()
    cfg.src_width, cfg.src_height, cfg.dst_width, cfg.dst_height = 1920, 1080, 1920, 1080
    cfg.content_weights[3] = 1.0  # delta_edges at full 1080p detector resolution does not fit on an SM
    with pytest.raises(capi.EsdError) as e:
        capi.EsdContext(cfg, 0)
    assert "delta_edges" in str(e.value)


@pytest.mark.slow
@pytest.mark.parametrize("name,batch", [("c2_1080p_full", 1024), ("c4_4k_full", 256)])
def test_full_length_baseline_configs_vs_cv2_golden(name, batch):
    """BASELINE config 2 (18 000 x 1080p, 10 min) and config 4 (3 600 x 4K60) at FULL length: per-frame integer
    sums, float64 content_val / adaptive_ratio / hist_diff and all cut lists equal the cv2 run frozen in
    tests/golden (generated by tests/golden/make_golden.py --full)."""
    g = load_golden(f"clip_{name}.npz")
    w, h, n, seed = int(g["width"]), int(g["height"]), int(g["n_frames"]), int(g["seed"])
    sch = synth.build_schedule(seed, n)
    dets = [ContentDetector(threshold=27.0, min_scene_len=15), AdaptiveDetector(adaptive_threshold=3.0, window_width=2),
            HistogramDetector(threshold=0.05, bins=256, min_scene_len=15),
            ThresholdDetector(threshold=12, min_scene_len=15, add_final_scene=True)]
    sm = SceneManager(batch_frames=batch)
    for d in dets:
        sm.add_detector(d)

    def batches():
        for a in range(0, n, batch):
            yield gpu_clip(seed, w, h, sch.descs[a:a + batch])

    from eioku_b200.scene_manager import BatchVideo
    assert sm.detect_scenes(BatchVideo(batches(), (w, h), 30.0), collect_scores=True) == n
    sc = sm.scores
    assert same_f64(sc["average_rgb"], g["average_rgb"])
    assert sm.cuts_of(dets[3]) == g["cuts_threshold"].tolist() and len(g["cuts_threshold"]) >= 1
    assert np.array_equal(sc["sums3"], g["sums3"])
    assert same_f64(sc["content_val"], g["content_val"])
    assert same_f64(sc["adaptive_ratio"], g["adaptive_ratio"])
    assert same_f64(sc["hist_diff"], g["hist_diff"])
    assert hashlib.sha256(np.ascontiguousarray(sc["hist"]).tobytes()).digest() == bytes(g["hist_sha256"])
    assert sm.cuts_of(dets[0]) == g["cuts_content"].tolist() and len(g["cuts_content"]) > 20
    assert sm.cuts_of(dets[1]) == g["cuts_adaptive"].tolist()
    assert sm.cuts_of(dets[2]) == g["cuts_hist"].tolist()
    scenes = sm.get_scene_list(start_in_scene=True)
    assert scenes[0][0] == 0 and scenes[-1][1] == n and len(scenes) == len(sm.get_cut_list()) + 1
    sm.close()


@pytest.mark.slow
def test_full_size_1080p_properties():
    """BASELINE config-2 size (1080p, 2048-frame batches): size-independent properties of the full-size path:
    idempotence, |a-b| == |b-a| under frame reversal, zero deltas for duplicated frames, histogram mass."""
    w, h, n, seed = 1920, 1080, 2048, 1002
    sch = synth.build_schedule(seed, n)
    clip = gpu_clip(seed, w, h, sch.descs)
    with make_ctx(w, h, None) as ctx:
        ctx.push_tensor(clip, 0)
        a = ctx.read_scores(0, n)
        cuts_a = ctx.get_cuts(capi.ESD_DET_CONTENT)[0]
        ctx.reset()
        ctx.push_tensor(clip, 0)
        b = ctx.read_scores(0, n)
        assert all(np.array_equal(a[k], b[k], equal_nan=True) for k in a)
        assert ctx.get_cuts(capi.ESD_DET_CONTENT)[0] == cuts_a
        ctx.reset()
        ctx.push_tensor(torch.flip(clip, dims=[0]), 0)
        r = ctx.read_scores(0, n)
        assert np.array_equal(r["sums3"][1:], a["sums3"][1:][::-1])
        assert np.array_equal(r["hist"], a["hist"][::-1])
        ctx.reset()
        dup = clip[100:101].expand(64, h, w, 3).contiguous()
        ctx.push_tensor(dup, 0)
        d = ctx.read_scores(0, 64)
        assert not d["sums3"].any() and np.all(d["hist_diff"][1:] == 1.0)
    assert np.all(a["hist"].sum(axis=1) == 256 * 144)
    g = load_golden("clip_c2_1080p_head.npz")
    m = int(g["n_frames"])
    assert np.array_equal(a["sums3"][:m], g["sums3"]) and same_f64(a["content_val"][:m], g["content_val"])
    hard = [c for c in sch.hard_cuts if c >= 15]
    assert set(hard) <= set(cuts_a)


def test_threshold_detector_state_machine_vectors():
    """ThresholdDetector fade state machine (FLOOR / CEILING, fade_bias, min_scene_len, add_final_scene) on random
    brightness traces: device decision pass vs the oracle class fed 1x1 frames of the same averages."""
    rng = np.random.default_rng(12)
    for trial in range(24):
        n = int(rng.integers(5, 400))
        method = trial % 2
        thr = int(rng.integers(5, 60))
        L = int(rng.choice([0, 1, 7, 15]))
        bias = float(rng.choice([0.0, -1.0, 1.0, 0.35, -0.6]))
        level = rng.integers(0, 2, n)  # slow random telegraph + noise around the threshold
        for i in range(1, n):
            if rng.random() > 0.08:
                level[i] = level[i - 1]
        avg = np.where(level > 0, thr + rng.integers(0, 40, n), thr - rng.integers(1, 5, n)).clip(0, 255).astype(np.float64)
        start = int(rng.choice([0, 1, 500]))
        o = P.ThresholdDetector(threshold=thr, min_scene_len=L, fade_bias=bias, add_final_scene=True, method=method)
        want = []
        for i in range(n):
            want += o.process_frame(start + i, np.full((1, 1, 3), avg[i], np.uint8))
        with make_ctx(64, 48, None, detectors=capi.ESD_DET_THRESHOLD, thresh_threshold=float(thr), thresh_min_scene_len=L,
                      thresh_fade_bias=bias, thresh_method=method, thresh_add_final_scene=1) as ctx:
            cuts, _ = ctx.decide_arrays(capi.ESD_DET_THRESHOLD, start, avg)
        assert cuts == want, (trial, method, thr, L, bias)


def test_threshold_detector_frames_and_post_process():
    """Fade to black at the end of a clip: the final scene is closed by post_process when add_final_scene is set."""
    w, h, n = 320, 180, 90
    fr = np.full((n, h, w, 3), 120, np.uint8)
    fr[30:40] = 3          # fade out / in inside the clip
    fr[70:] = 1            # ends faded out
    fr[:, ::7, ::5, 1] += 9
    for kw in (dict(add_final_scene=True), dict(add_final_scene=False), dict(add_final_scene=True, min_scene_len=60, fade_bias=1.0)):
        mine = ThresholdDetector(**kw)
        theirs = P.ThresholdDetector(**kw)
        got, want = [], []
        for k in range(n):
            got += mine.process_frame(k, fr[k])
            want += theirs.process_frame(k, fr[k])
        assert got == want
        assert mine.post_process(n - 1) == theirs.post_process(n - 1)
        sm = SceneManager()
        det = ThresholdDetector(**kw)
        sm.add_detector(det)
        sm.auto_downscale = False
        sm.detect_scenes(TensorVideo(torch.from_numpy(fr).to(DEV), 30.0), collect_scores=True)
        assert sm.cuts_of(det) == want + theirs.post_process(n - 1)
        assert same_f64(sm.scores["average_rgb"], np.array(theirs.averages))
        sm.close(); mine.close()


def test_delta_edges_canny_dilate_vs_oracle():
    """ContentDetector with a delta_edges weight -- numpy.median thresholds, cv2.Canny
    and cv2.dilate restated on the device; edge-change counts, float64 content_val and cuts equal the oracle
    (whose Canny/dilate restatement is itself pinned to cv2 in tests/test_oracle.py)."""
    w, h, n, seed = 1280, 720, 140, 1001
    sch = synth.build_schedule(seed, n, min_len=15, max_len=50)
    clip = gpu_clip(seed, w, h, sch.descs)
    frames = clip.cpu().numpy()
    W = (1.0, 0.5, 1.0, 0.8)
    for ks in (None, 3, 7):
        o = P.ContentDetector(threshold=20.0, min_scene_len=10, weights=P.Components(*W), kernel_size=ks, backend="closed_form")
        want, _ = P.detect(frames, [o], backend="closed_form")
        sm = SceneManager(batch_frames=37)
        det = ContentDetector(threshold=20.0, min_scene_len=10, weights=ContentDetector.Components(*W), kernel_size=ks)
        sm.add_detector(det)
        sm.detect_scenes(TensorVideo(clip, 30.0), collect_scores=True)
        assert (sm.scores["edge_counts"].astype(np.int64) * 255).tolist() == o.edge_sums, ks
        assert same_f64(sm.scores["content_val"], np.array(o.scores)), ks
        assert sm.cuts_of(det) == want and len(want) >= 2
        assert max(o.edge_sums) > 0
        sm.close()
    # frame by frame, stand-alone, on small odd-sized frames (dilation window clipped at the borders)
    rng = np.random.default_rng(8)
    import cv2 as _cv2  # only to blur the random test frames
    small = np.stack([_cv2.GaussianBlur(rng.integers(0, 256, (37, 53, 3), dtype=np.uint8), (0, 0), 1.2 + 0.2 * (k % 4)) for k in range(12)])
    mine = ContentDetector(weights=ContentDetector.Components(0.0, 0.0, 0.0, 1.0), threshold=5.0, min_scene_len=0)
    theirs = P.ContentDetector(weights=P.Components(0.0, 0.0, 0.0, 1.0), threshold=5.0, min_scene_len=0, backend="closed_form")
    for k in range(12):
        assert mine.process_frame(k, small[k]) == theirs.process_frame(k, small[k]), k
    mine.close()


def test_model_manager_on_a_real_video_file(tmp_path):
    """The ml-service surface end to end: `await ModelManager().detect_scenes(video_path, config)` on an .mp4 written
    and decoded by OpenCV (decode is out of scope, so it stays on the CPU), against the oracle on the same decoded frames."""
    import asyncio

    cv2 = pytest.importorskip("cv2")
    from eioku_b200.service import ModelManager

    w, h, n, seed = 640, 360, 180, 41
    sch = synth.build_schedule(seed, n, min_len=25, max_len=60, noise_amp=0)
    frames = co.synth_frames(seed, w, h, sch.descs)
    path = str(tmp_path / "clip.mp4")
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), 25.0, (w, h))
    if not vw.isOpened():
        pytest.skip("no mp4v encoder in this OpenCV build")
    for f in frames:
        vw.write(f)
    vw.release()
    cap = cv2.VideoCapture(path)
    decoded = []
    while True:
        ok, fr = cap.read()
        if not ok:
            break
        decoded.append(fr)
    cap.release()
    assert len(decoded) == n
    for cfg, mk in (({}, lambda: [P.ContentDetector(backend="closed_form")]),
                    ({"detector": "adaptive+threshold", "fade_threshold": 20, "min_scene_len": 10},
                     lambda: [P.AdaptiveDetector(min_scene_len=10, backend="closed_form"), P.ThresholdDetector(threshold=20, min_scene_len=10)])):
        want = P.detect_scenes_dicts(decoded, mk(), 25.0, backend="closed_form")
        got = asyncio.run(ModelManager().detect_scenes(path, cfg))
        assert got == want and len(got["scenes"]) >= 3
    import eioku_b200
    stats_csv = str(tmp_path / "stats.csv")
    scenes = eioku_b200.detect(path, ContentDetector(), stats_file_path=stats_csv)  # scenedetect.detect() counterpart
    ref_det = P.ContentDetector(backend="closed_form")
    ref_cuts, _ = P.detect(decoded, [ref_det], backend="closed_form")
    assert scenes == P.get_scenes_from_cuts(ref_cuts, 0, n)
    rows = open(stats_csv).read().strip().split("\n")
    assert rows[0].startswith("Frame Number,Timecode,content_val") and len(rows) == n  # header + frames 2..n (frame 1 has no score)
    assert rows[1].split(",")[:2] == ["2", "00:00:00.040"] and float(rows[1].split(",")[2]) == ref_det.scores[1]
    np.save(str(tmp_path / "clip.npy"), np.stack(decoded))
    assert asyncio.run(ModelManager().detect_scenes(str(tmp_path / "clip.npy"), {"fps": 25.0})) == \
        P.detect_scenes_dicts(decoded, [P.ContentDetector(backend="closed_form")], 25.0, backend="closed_form")


def test_adaptive_and_hist_decisions_on_random_score_traces():
    """AdaptiveDetector rolling-window ratio (incl. the zero-average branches) + min_scene_len rule and the
    HistogramDetector rule (incl. its `if not self._last_scene_cut` quirk) on random score arrays: device decision
    pass through esd_decide_arrays vs the oracle classes replaying the same scores."""
    rng = np.random.default_rng(99)

    class ReplayAdaptive(P.AdaptiveDetector):
        def _calculate_frame_score(self, frame_num, frame_img):
            return frame_img  # np.float64 score (python float 0.0 for the first frame, like PySceneDetect)

    for trial in range(30):
        n = int(rng.integers(1, 300))
        w = int(rng.choice([1, 2, 3, 5]))
        L = int(rng.choice([0, 1, 5, 15]))
        thr = float(rng.choice([1.5, 3.0, 7.0]))
        mcv = float(rng.choice([0.0, 5.0, 15.0]))
        start = int(rng.choice([0, 7, 1000]))
        base = rng.gamma(1.5, 2.0, n)
        base[rng.random(n) < 0.08] *= rng.uniform(5, 30)
        if trial % 3 == 0:
            base[rng.random(n) < 0.5] = 0.0  # stretches of identical frames -> average_is_zero branches
        base[0] = 0.0
        scores = base.astype(np.float64)
        o = ReplayAdaptive(adaptive_threshold=thr, min_scene_len=L, window_width=w, min_content_val=mcv, backend="closed_form")
        want = []
        for i in range(n):
            want += o.process_frame(start + i, np.float64(scores[i]) if i else 0.0)
        with make_ctx(64, 48, None, detectors=capi.ESD_DET_ADAPTIVE, adaptive_threshold=thr, adaptive_min_scene_len=L,
                      adaptive_window_width=w, adaptive_min_content_val=mcv) as ctx:
            cuts, ratio = ctx.decide_arrays(capi.ESD_DET_ADAPTIVE, start, scores)
        assert cuts == want, (trial, w, L, thr, mcv)
        for t, r in o.ratios.items():
            assert np.float64(r).view(np.uint64) == ratio[t - start].view(np.uint64), (trial, t)
        assert np.isnan(ratio[:min(w, n)]).all() and (n <= w or np.isnan(ratio[max(n - w, 0):]).all())

    for trial in range(30):
        n = int(rng.integers(1, 300))
        L = int(rng.choice([0, 1, 5, 15]))
        thr = float(rng.choice([0.05, 0.3, 0.9]))
        start = int(rng.choice([0, 1, 42]))
        diffs = np.clip(1.0 - rng.gamma(0.5, 0.1, n), -1.0, 1.0)
        diffs[0] = np.nan
        T = max(0.0, min(1.0, 1.0 - thr))
        last = None
        want = []
        for i in range(n):  # HistogramDetector.process_frame with the histogram comparison replaced by diffs[i]
            fn = start + i
            if not last:
                last = fn
            if i > 0 and diffs[i] <= T and (fn - last) >= L:
                want.append(fn)
                last = fn
        with make_ctx(64, 48, None, detectors=capi.ESD_DET_HIST, hist_threshold=thr, hist_min_scene_len=L) as ctx:
            cuts, _ = ctx.decide_arrays(capi.ESD_DET_HIST, start, diffs)
        assert cuts == want, (trial, L, thr, start)
