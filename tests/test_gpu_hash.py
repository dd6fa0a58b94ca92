"""GPU parity of the HashDetector path: fused kernel gray plane -> hash_kernel (INTER_AREA thumbnail,
DCT, median bits) -> hash_dist_kernel -> decide_kernel, against real cv2 (via oracle/psd_cv2.HashDetector) and the
committed cv2-derived goldens.

Parity bar: the integer stages (BGR2GRAY, INTER_AREA thumbnail) are bit-exact; the DCT is tolerance-parity -- cv2.dct's
own float32 output differs between its IPP and plain code paths -- so hash bits must equal cv2's wherever the
coefficient is more than MARGIN away from the median, hash_dist may differ by at most the number of such unstable bits
in the two frames involved, and cut lists must be identical whenever no frame sits that close to the threshold."""
import hashlib

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
cv2 = pytest.importorskip("cv2")

from eioku_b200 import capi  # noqa: E402
import synthclip as synth
from eioku_b200.detectors import ContentDetector, HashDetector, StatsManager  # noqa: E402
from eioku_b200.scene_manager import BatchVideo, SceneManager  # noqa: E402
from eioku_b200.service import detect_scenes_frames  # noqa: E402
from oracle import closed_form as cf  # noqa: E402
from oracle import psd_cv2 as P  # noqa: E402

DEV = "cuda:0"
MARGIN = 4e-6  # |coefficient - median| below which a hash bit is implementation-dependent (tests/golden/make_golden.py)


def cv2_thumb(frame, s):
    return cv2.resize(cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY), (s, s), interpolation=cv2.INTER_AREA)


def check_bits(got, want, margins, what=""):
    """got/want bool [n, s, s]; margins float [n, s, s].  Returns the per-frame count of unstable bits."""
    unstable = np.asarray(margins) <= MARGIN
    bad = (got != want) & ~unstable
    assert not bad.any(), f"{what}: {int(bad.sum())} stable hash bits differ from cv2 (first at {np.argwhere(bad)[0]})"
    return unstable.reshape(len(got), -1).sum(1)


@pytest.mark.parametrize("w,h,size,lowpass", [
    (256, 144, 16, 2),   # the 16:9 detector resolution: ResizeArea_ (scale 8 x 4.5)
    (274, 154, 16, 2),   # <= 0.6.1 integer downscale of 1080p: fractional on both axes
    (256, 171, 16, 2), (100, 77, 16, 2), (285, 160, 8, 4), (33, 40, 16, 2),
    (64, 64, 16, 2),     # 2x2 box: the (sum + 2) >> 2 special case
    (128, 96, 16, 2),    # integral scales 4 x 3: ResizeAreaFast_
    (32, 32, 16, 2),     # no resize at all
    (320, 180, 3, 2),    # odd bit count (9): the median is the middle element
    (300, 200, 32, 2),   # the largest supported hash: 1024 bits from a 64 x 64 DCT
    (200, 120, 5, 4),
    (1920, 1080, 16, 2),  # stand-alone detector on full-resolution frames: 60-byte integer column sums, 3 bands of rows
    (640, 360, 16, 2), (1000, 2000, 16, 2),
])
def test_hash_stages_vs_cv2(w, h, size, lowpass):
    rng = np.random.default_rng(w * 131 + h * 7 + size)
    n = 9
    # smooth-ish images plus noise (pure noise makes every coefficient large; gradients exercise near-median ones)
    base = rng.integers(0, 256, (n, 6, 8, 3), dtype=np.uint8)
    frames = np.stack([cv2.resize(b, (w, h), interpolation=cv2.INTER_CUBIC) for b in base])
    frames = np.clip(frames.astype(np.int16) + rng.integers(-6, 7, frames.shape), 0, 255).astype(np.uint8)
    frames[3] = 0                       # all black: max_value == 0 -> 1
    frames[4] = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    frames[5] = frames[4]               # identical frames: distance 0
    det = HashDetector(size=size, lowpass=lowpass, min_scene_len=2)
    ref = P.HashDetector(size=size, lowpass=lowpass, min_scene_len=2)
    cuts = det.process_frames(0, frames[:4])
    s = size * lowpass
    for k in range(4):   # test hook: thumbnails of the most recent push
        assert np.array_equal(det._ctx.debug_hash_input(k), cv2_thumb(frames[k], s)), f"thumbnail of frame {k}"
    cuts += det.process_frames(4, frames[4:])
    for k in range(4, n):
        assert np.array_equal(det._ctx.debug_hash_input(k), cv2_thumb(frames[k], s)), f"thumbnail of frame {k}"
    want_cuts = []
    for k in range(n):
        want_cuts += ref.process_frame(k, frames[k])
    bits, dist = det._ctx.read_hash(0, n)
    nun = check_bits(bits, np.array(ref.hashes), np.array(ref.margins), f"{w}x{h} size {size}")
    # esd_read_hash_margin: the device reports how far each frame's weakest bit is from flipping; it agrees with cv2's own
    # coefficients up to the DCT noise, and every frame it calls stable (margin > 2 * MARGIN) hashes exactly like cv2
    margin = det._ctx.read_hash_margin(0, n)
    ref_min = np.array(ref.margins, np.float64).reshape(n, -1).min(1)
    assert margin.dtype == np.float32 and np.all(np.abs(margin - ref_min) <= MARGIN), (margin, ref_min)
    for k in range(n):
        if margin[k] > 2 * MARGIN:
            assert np.array_equal(bits[k], np.array(ref.hashes[k])), k
    assert margin[3] == 0.0  # the black frame: every coefficient equals the median
    assert np.isnan(dist[0])
    for k in range(1, n):
        assert abs(dist[k] - ref.dists[k]) * size * size <= nun[k] + nun[k - 1] + 1e-9
    if nun.sum() == 0:
        assert np.array_equal(dist[1:], np.array(ref.dists[1:]))
        assert cuts == want_cuts
    assert dist[5] == 0.0
    det.close()


@pytest.mark.parametrize("name", ["c1_720p", "c2_1080p_head", "c4_4k_head_s8l4"])
def test_hash_clip_golden(name):
    """BASELINE clips through SceneManager (fused downscale) with HashDetector + ContentDetector from one pass, against
    the cv2-derived golden: thumbnails bit-exact, stable bits identical, hash_dist and cut list identical up to the
    documented DCT tolerance."""
    g = load_golden(f"hash_{name}.npz")
    w, h, n, seed = int(g["width"]), int(g["height"]), int(g["n_frames"]), int(g["seed"])
    size, lowpass = int(g["size"]), int(g["lowpass"])
    sch = synth.build_schedule(seed, n)
    batch = {1280: 257, 1920: 128, 3840: 40}[w]
    sm = SceneManager(batch_frames=batch, stats_manager=StatsManager())
    hd = HashDetector(threshold=0.395, size=size, lowpass=lowpass, min_scene_len=15)
    sm.add_detector(hd)
    sm.add_detector(ContentDetector())
    thumbs = hashlib.sha256()
    ctx_box = {}

    def batches():
        for a in range(0, n, batch):
            if ctx_box and a > 0:   # thumbnails of the batch pushed last
                for k in range(a - batch, a):
                    thumbs.update(sm._ctx.debug_hash_input(k).tobytes())
            out = torch.empty((len(sch.descs[a:a + batch]), h, w, 3), dtype=torch.uint8, device=DEV)
            synth.fill(out, seed, sch.descs[a:a + batch])
            ctx_box["on"] = True
            yield out

    assert sm.detect_scenes(BatchVideo(batches(), (w, h), 30.0), collect_scores=True) == n
    bits = sm.scores["hash_bits"].reshape(n, -1)
    want = np.unpackbits(g["bits"], axis=1, bitorder="little")[:, :size * size].astype(bool)
    unstable = np.unpackbits(g["unstable"], axis=1, bitorder="little")[:, :size * size].astype(bool)
    bad = (bits != want) & ~unstable
    assert not bad.any(), f"{int(bad.sum())} stable bits differ"
    nun = unstable.sum(1)
    dist, wd = sm.scores["hash_dist"], g["hash_dist"]
    assert np.isnan(dist[0]) and np.isnan(wd[0])
    tol = (nun[1:] + nun[:-1]) / float(size * size)
    assert np.all(np.abs(dist[1:] - wd[1:]) <= tol + 1e-12)
    exact = tol == 0
    assert np.array_equal(dist[1:][exact], wd[1:][exact])
    # no frame of these clips sits within the tolerance of the threshold, so the cut lists must be identical
    assert not np.any(np.abs(wd[1:] - 0.395) <= tol)
    assert sm.cuts_of(hd) == g["cuts_hash"].tolist()
    key = hd.get_metrics()[0]
    assert key == f"hash_dist [size={size} lowpass={lowpass}]"
    assert sm.stats_manager.get_metrics(5, [key])[0] == float(dist[5])
    assert sm.stats_manager.get_metrics(0, [key])[0] is None


def test_hash_thumbnails_sha_vs_golden():
    """Integer stage over a whole clip: sha256 of every INTER_AREA thumbnail equals cv2's (config-2 head, fused 1080p
    downscale feeding the gray plane)."""
    g = load_golden("hash_c2_1080p_head.npz")
    w, h, n, seed = int(g["width"]), int(g["height"]), int(g["n_frames"]), int(g["seed"])
    sch = synth.build_schedule(seed, n)
    cfg = capi.default_config()
    cfg.detectors = capi.ESD_DET_HASH
    cfg.src_width, cfg.src_height = w, h
    sha = hashlib.sha256()
    with capi.EsdContext(cfg, 0) as ctx:
        for a in range(0, n, 100):
            out = torch.empty((100, h, w, 3), dtype=torch.uint8, device=DEV)
            synth.fill(out, seed, sch.descs[a:a + 100])
            ctx.push_tensor(out, a)
            for k in range(a, a + 100):
                sha.update(ctx.debug_hash_input(k).tobytes())
        with pytest.raises(capi.EsdError):
            ctx.debug_hash_input(0)  # only the most recent push is kept
    assert sha.digest() == bytes(g["thumbs_sha256"])


def test_hash_batch_split_invariance_and_reset():
    rng = np.random.default_rng(5)
    frames = torch.from_numpy(rng.integers(0, 256, (40, 90, 160, 3), dtype=np.uint8)).to(DEV)
    frames[10:20] = frames[9]
    cfg = capi.default_config()
    cfg.detectors = capi.ESD_DET_HASH | capi.ESD_DET_CONTENT
    cfg.src_width, cfg.src_height = 160, 90
    cfg.hash_min_scene_len = 3
    with capi.EsdContext(cfg, 0) as a, capi.EsdContext(cfg, 0) as b:
        a.push_tensor(frames, 100)
        pos = 100
        for m in (1, 2, 7, 13, 17):
            b.push_tensor(frames[pos - 100:pos - 100 + m], pos)
            pos += m
        ba, da = a.read_hash(100, 40)
        bb, db = b.read_hash(100, 40)
        assert np.array_equal(ba, bb) and np.array_equal(da[1:], db[1:]) and np.isnan(da[0]) and np.isnan(db[0])
        assert a.get_cuts(capi.ESD_DET_HASH) == b.get_cuts(capi.ESD_DET_HASH)
        assert a.get_cuts(capi.ESD_DET_CONTENT) == b.get_cuts(capi.ESD_DET_CONTENT)
        assert np.all(da[11:20] == 0.0)
        b.reset()
        b.push_tensor(frames[:5], 0)
        _, d2 = b.read_hash(0, 5)
        assert np.isnan(d2[0]) and np.array_equal(d2[1:], da[1:5])


def test_hash_decision_random_traces():
    """esd_decide_arrays(ESD_DET_HASH): `dist >= threshold and frame - last_cut >= min_scene_len`, last_cut starting at
    the first frame (HashDetector.process_frame)."""
    rng = np.random.default_rng(17)
    for L in (0, 1, 7, 15):
        for thr in (0.1, 0.395):
            cfg = capi.default_config()
            cfg.detectors = capi.ESD_DET_HASH
            cfg.src_width, cfg.src_height = 64, 64
            cfg.hash_threshold, cfg.hash_min_scene_len = thr, L
            with capi.EsdContext(cfg, 0) as ctx:
                for start in (0, 500):
                    n = 9000
                    d = rng.integers(0, 257, n) / 256.0
                    d[rng.random(n) < 0.9] *= 0.2
                    d[0] = np.nan
                    want, last = [], start
                    for i in range(1, n):
                        if d[i] >= thr and (start + i - last) >= L:
                            want.append(start + i)
                            last = start + i
                    got, _ = ctx.decide_arrays(capi.ESD_DET_HASH, start, d)
                    assert got == want


def test_hash_through_service_and_validation():
    rng = np.random.default_rng(9)
    a = rng.integers(0, 256, (1, 72, 128, 3), dtype=np.uint8).repeat(30, 0)
    b = rng.integers(0, 256, (1, 72, 128, 3), dtype=np.uint8).repeat(30, 0)
    clip = np.concatenate([a, b, a])
    out = detect_scenes_frames(clip, {"detector": "hash", "fps": 30.0})
    assert [s["start_ms"] for s in out["scenes"]] == [0, 1000, 2000]
    out2 = detect_scenes_frames(clip, {"detector": "hash+content", "size": 8, "lowpass": 4, "hash_threshold": 0.3})
    assert len(out2["scenes"]) == 3
    for kw, exc in [(dict(size=3, lowpass=3), capi.EsdError), (dict(size=40, lowpass=1), capi.EsdError),
                    (dict(size=16, lowpass=8), capi.EsdError)]:
        det = HashDetector(**kw)
        with pytest.raises(exc):
            det.process_frames(0, clip[:2])
    det = HashDetector()   # thumbnail larger than the frame
    with pytest.raises(capi.EsdError):
        det.process_frames(0, clip[:2, :20, :20].copy())
