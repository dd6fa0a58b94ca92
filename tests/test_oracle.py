"""CPU suite: the oracle against real cv2 (when importable) and against the committed cv2-derived
golden vectors -- this is what pins the oracle."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
import synthclip as synth
from oracle import c_oracle as co
from oracle import closed_form as cf
from oracle import psd_cv2 as P

STAGES = ["stage_1920x1080_to_256x144", "stage_1920x1080_to_274x154", "stage_1280x720_to_256x144",
          "stage_3840x2160_to_256x144", "stage_854x480_to_285x160", "stage_300x200_to_256x171"]


def _stage_geom(name):
    src, dst = name.split("_")[1], name.split("_")[3]
    return tuple(map(int, src.split("x"))) + tuple(map(int, dst.split("x")))


@pytest.mark.parametrize("name", STAGES)
def test_resize_hsv_y_vs_golden(name):
    w, h, dw, dh = _stage_geom(name)
    g = load_golden(name + ".npz")
    img = np.random.default_rng(int(g["seed"])).integers(0, 256, (h, w, 3), dtype=np.uint8)
    small_c = co.resize_linear(img, dw, dh)
    assert np.array_equal(small_c, g["small"])  # C restatement == cv2.resize
    if w <= 1920:
        assert np.array_equal(cf.resize_linear_u8(img, dw, dh), g["small"])  # numpy restatement == cv2.resize
    hsv = co.bgr2hsv(small_c)
    assert hashlib.sha256(hsv.tobytes()).digest() == bytes(g["hsv_sha256"])
    assert np.array_equal(cf.bgr2hsv_u8(small_c), hsv)
    y = co.bgr2y(small_c)
    assert hashlib.sha256(y.tobytes()).digest() == bytes(g["y_sha256"])
    assert np.array_equal(cf.y_histogram(cf.bgr2y_u8(small_c)), g["y_hist"])
    assert np.array_equal(hsv.reshape(-1, 3).sum(0, dtype=np.int64), g["hsv_sums"])


def test_exhaustive_cube_vs_cv2_checksums():
    """All 2^24 BGR values through the C restatement; checksums were taken from cv2 4.13.0."""
    ref = json.load(open(os.path.join(GOLDEN, "cube.json")))
    x = np.arange(1 << 24, dtype=np.uint32)
    img = np.stack([x & 255, (x >> 8) & 255, (x >> 16) & 255], -1).astype(np.uint8)
    assert hashlib.sha256(co.bgr2hsv(img).tobytes()).hexdigest() == ref["hsv_sha256"]
    assert hashlib.sha256(co.bgr2y(img).tobytes()).hexdigest() == ref["y_sha256"]
    sub = img[::17]
    assert np.array_equal(cf.bgr2hsv_u8(sub), co.bgr2hsv(sub))
    assert np.array_equal(cf.bgr2y_u8(sub), co.bgr2y(sub))
    assert co.bgr2hsv(img)[..., 0].max() == 179


def test_axis_tables_c_equals_numpy():
    for src, dst in [(1920, 256), (1080, 144), (1920, 274), (3840, 256), (854, 285), (300, 256), (512, 256), (257, 256), (10, 7)]:
        a = co.axis_tables(src, dst)
        b = cf.linear_axis_tables(src, dst)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        assert np.all(a[2].astype(int) + a[3].astype(int) == 2048)
    assert len(cf.touched_rows(1080, 144)) == 288 and len(cf.touched_rows(720, 144)) == 288
    assert len(cf.touched_rows(2160, 144)) * 3840 * 3 == 3317760  # 4K: 288 rows touched


def test_downscale_factor_and_size():
    assert cf.compute_downscale_factor(1920) == 7.5 and cf.compute_downscale_factor(1920, mode="int") == 7
    assert cf.compute_downscale_factor(255) == 1 and cf.compute_downscale_factor(256) == 1.0
    assert cf.downscaled_size(1920, 1080, 7.5) == (256, 144)
    assert cf.downscaled_size(1920, 1080, 7) == (274, 154)
    assert cf.downscaled_size(1280, 720, 5.0) == (256, 144)
    assert cf.downscaled_size(3840, 2160, 15.0) == (256, 144)
    assert cf.downscaled_size(200, 100, 1) == (200, 100)


def test_live_cv2_when_available():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(42)
    for (w, h) in [(1920, 1080), (641, 333), (512, 288), (3840, 2160)]:
        f = cf.compute_downscale_factor(w)
        dw, dh = cf.downscaled_size(w, h, f)
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        small = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(co.resize_linear(img, dw, dh), small)
        assert np.array_equal(co.bgr2hsv(small), cv2.cvtColor(small, cv2.COLOR_BGR2HSV))
        assert np.array_equal(co.bgr2y(small), cv2.cvtColor(small, cv2.COLOR_BGR2YUV)[..., 0])
    for bins in (256, 255, 100, 17, 3):
        a = rng.integers(0, 256, (144, 256, 3), dtype=np.uint8)
        b = np.clip(a.astype(int) + rng.integers(-40, 40, a.shape), 0, 255).astype(np.uint8)
        ca, ha = P.calculate_histogram(a, bins, "cv2")
        cb, hb = P.calculate_histogram(b, bins, "cv2")
        ca2, ha2 = P.calculate_histogram(a, bins, "closed_form")
        assert np.array_equal(ca, ca2) and np.array_equal(ha.view(np.uint32), ha2.view(np.uint32))
        assert cv2.compareHist(ha, hb, cv2.HISTCMP_CORREL) == P.compare_hist_correl(ha, hb)


def test_filter_vectors_regression():
    vecs = json.load(open(os.path.join(GOLDEN, "filter_vectors.json")))["vectors"]
    assert len(vecs) > 50
    for v in vecs:
        f = P.FlashFilter(v["mode"], v["length"])
        cuts = []
        for i, a in enumerate(v["above"]):
            cuts += f.filter(v["start"] + i, bool(a))
        assert cuts == v["cuts"]
    # hand-checked cases of SURVEY.md A.5
    f = P.FlashFilter(P.FILTER_SUPPRESS, 15)
    assert sum((f.filter(i, i in (3, 20, 25, 50)) for i in range(60)), []) == [20, 50]
    f = P.FlashFilter(P.FILTER_MERGE, 15)
    assert sum((f.filter(i, i in (20, 25, 27, 70)) for i in range(90)), []) == [20, 70]


def test_closed_form_backend_reproduces_clip_golden_prefix():
    """The cv2-free oracle (C integer stages + restated float stages) equals the cv2 run that produced
    tests/golden/clip_c1_720p.npz on the clip's first 300 frames."""
    g = load_golden("clip_c1_720p.npz")
    n = 300
    seed, w, h = int(g["seed"]), int(g["width"]), int(g["height"])
    sch = synth.build_schedule(seed, int(g["n_frames"]))
    frames = co.synth_frames(seed, w, h, sch.descs[:n])
    sums, hist, _ = co.score_frames(frames, 256, 144, bins=256)
    assert np.array_equal(sums.astype(np.uint64), g["sums3"][:n])
    assert np.array_equal(hist[0], g["hist_first"])
    cd = P.ContentDetector(backend="closed_form")
    ad = P.AdaptiveDetector(backend="closed_form")
    hd = P.HistogramDetector(backend="closed_form")
    cuts, _ = P.detect(frames, [cd, ad, hd], backend="closed_form")
    assert np.array_equal(np.array(cd.scores).view(np.uint64), g["content_val"][:n].view(np.uint64))
    assert np.array_equal(np.array(hd.diffs)[1:].view(np.uint64), g["hist_diff"][1:n].view(np.uint64))
    for t, r in ad.ratios.items():
        assert np.float64(r).view(np.uint64) == g["adaptive_ratio"][t].view(np.uint64)
    want = sorted(set(c for k in ("cuts_content", "cuts_adaptive", "cuts_hist") for c in g[k].tolist() if c < n - 20))
    assert [c for c in cuts if c < n - 20] == want


def test_synthetic_clip_is_meaningful():
    """Within-scene scores stay far below the threshold, hard cuts far above, dissolves below."""
    g = load_golden("clip_c1_720p.npz")
    n = int(g["n_frames"])
    sch = synth.build_schedule(int(g["seed"]), n)
    s = g["content_val"]
    inside = np.ones(n, bool)
    for c in sch.hard_cuts:
        inside[c] = False
    for a, b in sch.dissolves + sch.fades + sch.flashes:
        inside[a:b + 1] = False
    assert s[inside].max() < 12.0 and s[inside][1:].mean() < 5.0
    assert all(s[c] > 35.0 for c in sch.hard_cuts)
    for a, b in sch.dissolves:
        assert s[a:b].max() < 27.0
    assert set(c for c in sch.hard_cuts if c >= 15) <= set(g["cuts_content"].tolist())
    assert len(sch.hard_cuts) >= 8 and len(sch.dissolves) >= 1


def test_schedule_is_deterministic():
    a = synth.build_schedule(1002, 5000)
    b = synth.build_schedule(1002, 5000)
    assert np.array_equal(a.descs, b.descs)
    assert hashlib.sha256(a.descs.tobytes()).hexdigest()[:16] == "84395418ff2fce5f"  # frozen: goldens depend on it
    assert a.descs.shape == (5000, 8) and (a.descs[:, 3] >= 1).all()
    assert a.fades and a.dissolves and a.flashes


def test_threshold_detector_oracle_hand_cases():
    """ThresholdDetector restatement on hand-checked traces (fade out at 10, back in at 20 -> cut at 15)."""
    def run(avgs, **kw):
        d = P.ThresholdDetector(**kw)
        cuts = []
        for i, a in enumerate(avgs):
            cuts += d.process_frame(i, np.full((2, 2, 3), a, np.uint8))
        return cuts, d.post_process(len(avgs) - 1)
    trace = [100] * 10 + [2] * 10 + [100] * 30
    assert run(trace, min_scene_len=15) == ([15], [])
    assert run(trace, min_scene_len=15, fade_bias=1.0) == ([20], [])
    assert run(trace, min_scene_len=15, fade_bias=-1.0) == ([10], [])
    assert run(trace, min_scene_len=25) == ([], [])                       # fade-in too close to the start
    assert run(trace + [1] * 5, min_scene_len=15, add_final_scene=True) == ([15], [50])
    assert run(trace + [1] * 5, min_scene_len=15) == ([15], [])
    assert run([1] * 5 + [90] * 40, min_scene_len=3) == ([2], [])          # starts faded out: cut at (5 + 0) / 2
    assert run([10, 30, 10, 30, 10, 30], threshold=20, min_scene_len=0, method=P.ThresholdDetector.CEILING)[0] == [1, 3]


def test_canny_dilate_restatement_vs_cv2():
    """ContentDetector._detect_edges (median thresholds, cv2.Canny, cv2.dilate) restated in numpy == real cv2."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    assert cf.estimated_kernel_size(256, 144) == 5 and cf.estimated_kernel_size(1920, 1080) == 13
    for t in range(24):
        h, w, k = int(rng.integers(6, 160)), int(rng.integers(6, 280)), int(rng.choice([3, 5, 7, 9]))
        lum = cv2.GaussianBlur(rng.integers(0, 256, (h, w), dtype=np.uint8), (0, 0), 0.8 + 0.3 * (t % 6))
        low, high = cf.canny_thresholds(lum)
        assert np.array_equal(cf.canny_u8(lum, low, high), cv2.Canny(lum, low, high))
        assert np.array_equal(cf.detect_edges(lum, k), cv2.dilate(cv2.Canny(lum, low, high), np.ones((k, k), np.uint8)))
    sch = synth.build_schedule(1001, 60, min_len=15, max_len=40)
    frames = co.synth_frames(1001, 640, 360, sch.descs)
    W = P.Components(1.0, 1.0, 1.0, 1.0)
    a, b = P.ContentDetector(weights=W, backend="cv2"), P.ContentDetector(weights=W, backend="closed_form")
    ca, _ = P.detect(frames, [a], backend="cv2")
    cb, _ = P.detect(frames, [b], backend="closed_form")
    assert ca == cb and a.edge_sums == b.edge_sums and max(a.edge_sums) > 0
    assert np.array_equal(np.array(a.scores).view(np.uint64), np.array(b.scores).view(np.uint64))


# ------------------------------------------------------------------------------------------------ HashDetector stages
def test_gray_restatement_exhaustive_vs_cv2():
    """cv2 BGR2GRAY uses the 15-bit coefficients (not the 14-bit Y of BGR2YUV): all 2^24 BGR values."""
    cv2 = pytest.importorskip("cv2")
    x = np.arange(1 << 24, dtype=np.uint32)
    img = np.stack([x & 255, (x >> 8) & 255, (x >> 16) & 255], -1).astype(np.uint8).reshape(4096, 4096, 3)
    assert np.array_equal(cv2.cvtColor(img, cv2.COLOR_BGR2GRAY), cf.bgr2gray_u8(img))


@pytest.mark.parametrize("w,h,s", [(256, 144, 32), (274, 154, 32), (285, 160, 32), (100, 77, 32), (64, 64, 32), (128, 96, 32),
                                   (32, 32, 32), (33, 40, 32), (300, 200, 64), (320, 180, 6), (200, 120, 20)])
def test_inter_area_restatement_vs_cv2(w, h, s):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(w * 7 + h)
    for _ in range(3):
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        assert np.array_equal(cf.resize_area_u8(img, s, s), cv2.resize(img, (s, s), interpolation=cv2.INTER_AREA))


def test_hash_closed_form_vs_cv2_golden():
    """The float64-DCT restatement reproduces the cv2-derived golden hashes on every stable bit, and the golden's
    thumbnails hash matches the restated integer stages."""
    import hashlib

    import synthclip as synth
    from oracle import c_oracle as co
    g = load_golden("hash_c2_1080p_head.npz")
    w, h, seed = int(g["width"]), int(g["height"]), int(g["seed"])
    n = 48
    sch = synth.build_schedule(seed, n)
    frames = co.synth_frames(seed, w, h, sch.descs)
    dw, dh = [int(v) for v in g["dst"]]
    want = np.unpackbits(g["bits"], axis=1, bitorder="little")[:n].astype(bool)
    unstable = np.unpackbits(g["unstable"], axis=1, bitorder="little")[:n].astype(bool)
    det = P.HashDetector(backend="closed_form")
    for k in range(n):
        small = cf.resize_linear_u8(frames[k], dw, dh)
        det.process_frame(k, small)
    got = np.array(det.hashes).reshape(n, -1)
    assert not ((got != want) & ~unstable).any()
    if not unstable.any():
        assert np.array_equal(np.array(det.dists)[1:], g["hash_dist"][1:n])


def test_nv12_to_bgr_restatement_exhaustive_vs_cv2():
    """Every (Y, U, V) combination: the closed form of cv2.cvtColor(COLOR_YUV2BGR_NV12) the CUDA kernel restates."""
    cv2 = pytest.importorskip("cv2")
    W = H = 4096
    c = np.arange(2 ** 22).reshape(H // 2, W // 2)
    uvid, grp = c % 65536, c // 65536
    uv = np.zeros((H // 2, W), np.uint8)
    uv[:, 0::2] = (uvid & 255).astype(np.uint8)
    uv[:, 1::2] = (uvid >> 8).astype(np.uint8)
    y = np.zeros((H, W), np.uint8)
    for dy in range(2):
        for dx in range(2):
            y[dy::2, dx::2] = (grp * 4 + dy * 2 + dx).astype(np.uint8)
    nv12 = np.vstack([y, uv])
    assert np.array_equal(cv2.cvtColor(nv12, cv2.COLOR_YUV2BGR_NV12), cf.nv12_to_bgr_u8(nv12))


def test_nv12_closed_form_pipeline_vs_cv2_golden():
    """Oracle restatement of the NV12 path (closed-form NV12 -> BGR, C resize / HSV sums) against the cv2-derived golden."""
    g = load_golden("nv12_c2_1080p_head.npz")
    w, h, seed = int(g["width"]), int(g["height"]), int(g["seed"])
    n = 40
    sch = synth.build_schedule(seed, int(g["n_frames"]), min_len=20, max_len=70)
    nv12 = synth.bgr_to_test_nv12(co.synth_frames(seed, w, h, sch.descs[:n]))
    bgr = np.stack([cf.nv12_to_bgr_u8(f) for f in nv12])
    dw, dh = [int(v) for v in g["dst"]]
    sums, _, _ = co.score_frames(bgr, dw, dh)
    assert np.array_equal(sums.astype(np.uint64), g["sums3"][:n])
