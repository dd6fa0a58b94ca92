"""Generates the committed golden vectors under tests/golden/ from REAL cv2 (4.13.0 in this
image) -- run in the build container, never on the GPU box:

    python tests/golden/make_golden.py [--full]

What is frozen (SURVEY.md section 8c "golden vectors to create"):
  stage_*.npz   cv2.resize / cvtColor outputs on seeded random frames (per-stage parity)
  cube.json     sha256 of cv2's BGR2HSV and BGR2YUV-Y over all 2^24 BGR values
  clip_*.npz    per-frame (sum|dH|, sum|dS|, sum|dV|), content_val, adaptive_ratio, Y-hist checksum,
                hist_diff and the three detectors' cut lists for the seeded synthetic clips of the
                BASELINE configs, computed by oracle/psd_cv2.py (PySceneDetect logic on real cv2)
                on frames from the CPU twin of the clip generator.
  filter_vectors.json   FlashFilter / min_scene_len state-machine unit vectors.
  nv12_*.npz    the NV12 input path: cv2.cvtColor(COLOR_YUV2BGR_NV12) + resize + detector logic on NV12 test content.
  hash_*.npz    HashDetector on the same clips, from real cv2 (cvtColor GRAY, resize INTER_AREA,
                dct): per-frame hash bits, the bits whose DCT coefficient lies within 4e-6 of the median
                ("unstable": cv2.dct's float32 rounding is build-dependent), hash_dist, cut list, and a sha256
                of the INTER_AREA thumbnails (integer stage, bit-exact).
--full also produces the full-length config-2 (18 000 x 1080p) and config-4 (3 600 x 4K) files.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import time

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import synthclip as synth  # noqa: E402
from oracle import c_oracle as co  # noqa: E402
from oracle import psd_cv2 as P  # noqa: E402

VERSIONS = {"cv2": cv2.__version__, "numpy": np.__version__, "python": sys.version.split()[0]}


def stage_vectors():
    cases = [(1920, 1080, 256, 144), (1920, 1080, 274, 154), (1280, 720, 256, 144), (3840, 2160, 256, 144),
             (854, 480, 285, 160), (300, 200, 256, 171)]
    for (w, h, dw, dh) in cases:
        rng = np.random.default_rng(w * 10007 + h)
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        small = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
        hsv = cv2.cvtColor(small, cv2.COLOR_BGR2HSV)
        y = cv2.cvtColor(small, cv2.COLOR_BGR2YUV)[..., 0]
        np.savez_compressed(os.path.join(HERE, f"stage_{w}x{h}_to_{dw}x{dh}.npz"), seed=w * 10007 + h, small=small,
                            hsv_sha256=np.frombuffer(hashlib.sha256(hsv.tobytes()).digest(), np.uint8),
                            y_sha256=np.frombuffer(hashlib.sha256(np.ascontiguousarray(y).tobytes()).digest(), np.uint8),
                            hsv_sums=hsv.reshape(-1, 3).sum(0, dtype=np.int64),
                            y_hist=np.bincount(y.ravel(), minlength=256).astype(np.uint32))


def cube():
    x = np.arange(1 << 24, dtype=np.uint32)
    img = np.stack([x & 255, (x >> 8) & 255, (x >> 16) & 255], -1).astype(np.uint8).reshape(4096, 4096, 3)
    hsv = cv2.cvtColor(img, cv2.COLOR_BGR2HSV)
    y = cv2.cvtColor(img, cv2.COLOR_BGR2YUV)[..., 0]
    gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    # per-row 256-bin histograms of Y (uint32 [4096][256]): what a GPU context with frames = rows of the cube can be asked for
    rowhist = np.stack([np.bincount(r, minlength=256) for r in y]).astype(np.uint32)
    out = {"hsv_sha256": hashlib.sha256(hsv.tobytes()).hexdigest(), "y_sha256": hashlib.sha256(np.ascontiguousarray(y).tobytes()).hexdigest(),
           "gray_sha256": hashlib.sha256(np.ascontiguousarray(gray).tobytes()).hexdigest(),
           "y_rowhist_sha256": hashlib.sha256(rowhist.tobytes()).hexdigest(),
           "layout": "index = b | g<<8 | r<<16, reshaped 4096x4096x3", "versions": VERSIONS}
    json.dump(out, open(os.path.join(HERE, "cube.json"), "w"), indent=1)


def clip(name, seed, w, h, n, chunk=64, downscale_mode="float"):
    t0 = time.time()
    sch = synth.build_schedule(seed, n)
    dets = {
        "content": P.ContentDetector(threshold=27.0, min_scene_len=15),
        "content_suppress": P.ContentDetector(threshold=27.0, min_scene_len=15, filter_mode=P.FILTER_SUPPRESS),
        "content_luma": P.ContentDetector(threshold=27.0, min_scene_len=15, luma_only=True),
        "adaptive": P.AdaptiveDetector(adaptive_threshold=3.0, window_width=2),
        "hist": P.HistogramDetector(threshold=0.05, bins=256, min_scene_len=15),
        "threshold": P.ThresholdDetector(threshold=12, min_scene_len=15, fade_bias=0.0, add_final_scene=True),
    }
    cuts = {k: [] for k in dets}

    def frames():
        for a in range(0, n, chunk):
            fr = co.synth_frames(seed, w, h, sch.descs[a:a + chunk])
            for f in fr:
                yield f

    from oracle import closed_form as cf
    factor = cf.compute_downscale_factor(w, mode=downscale_mode)
    dw, dh = cf.downscaled_size(w, h, factor)
    for k, frame in enumerate(frames()):
        small = cv2.resize(frame, (dw, dh), interpolation=cv2.INTER_LINEAR) if factor > 1 else frame
        for name_d, det in dets.items():
            cuts[name_d] += det.process_frame(k, small)
    cuts["threshold"] += dets["threshold"].post_process(n - 1)
    ad = dets["adaptive"]
    ratio = np.full(n, np.nan)
    for t, r in ad.ratios.items():
        ratio[t] = r
    hd = dets["hist"]
    counts = np.stack(hd.counts)
    np.savez_compressed(
        os.path.join(HERE, f"clip_{name}.npz"), seed=seed, width=w, height=h, n_frames=n, dst=np.array([dw, dh]),
        sums3=np.stack(dets["content"].sums).astype(np.uint64),
        content_val=np.array(dets["content"].scores), luma_val=np.array(dets["content_luma"].scores),
        adaptive_ratio=ratio, hist_diff=np.array(hd.diffs),
        hist_first=counts[0], hist_last=counts[-1],
        hist_sha256=np.frombuffer(hashlib.sha256(counts.astype(np.uint32).tobytes()).digest(), np.uint8),
        cuts_content=np.array(cuts["content"], np.int64), cuts_content_suppress=np.array(cuts["content_suppress"], np.int64),
        cuts_content_luma=np.array(cuts["content_luma"], np.int64), cuts_adaptive=np.array(cuts["adaptive"], np.int64),
        cuts_hist=np.array(cuts["hist"], np.int64), cuts_threshold=np.array(cuts["threshold"], np.int64),
        average_rgb=np.array(dets["threshold"].averages),
        hard_cuts=np.array(sch.hard_cuts, np.int64), versions=json.dumps(VERSIONS))
    print(f"clip_{name}: {n} frames {w}x{h} -> {dw}x{dh} in {time.time() - t0:.1f}s; cuts content={len(cuts['content'])} "
          f"adaptive={len(cuts['adaptive'])} hist={len(cuts['hist'])} threshold={len(cuts['threshold'])}", flush=True)


HASH_MARGIN = 4e-6  # cv2.dct IPP vs plain differ by <= ~2e-7 on [0,1] inputs; float64 DCT vs cv2 by <= ~1.5e-6


def hash_clip(name, seed, w, h, n, chunk=64, size=16, lowpass=2):
    t0 = time.time()
    sch = synth.build_schedule(seed, n)
    det = P.HashDetector(threshold=0.395, size=size, lowpass=lowpass, min_scene_len=15)
    from oracle import closed_form as cf
    factor = cf.compute_downscale_factor(w)
    dw, dh = cf.downscaled_size(w, h, factor)
    cuts = []
    thumbs = hashlib.sha256()
    k = 0
    for a in range(0, n, chunk):
        for f in co.synth_frames(seed, w, h, sch.descs[a:a + chunk]):
            small = cv2.resize(f, (dw, dh), interpolation=cv2.INTER_LINEAR) if factor > 1 else f
            cuts += det.process_frame(k, small)
            g = cv2.cvtColor(small, cv2.COLOR_BGR2GRAY)
            thumbs.update(cv2.resize(g, (size * lowpass, size * lowpass), interpolation=cv2.INTER_AREA).tobytes())
            k += 1
    bits = np.array(det.hashes).reshape(n, -1)
    unstable = np.array(det.margins).reshape(n, -1) <= HASH_MARGIN
    np.savez_compressed(os.path.join(HERE, f"hash_{name}.npz"), seed=seed, width=w, height=h, n_frames=n, dst=np.array([dw, dh]),
                        size=size, lowpass=lowpass, bits=np.packbits(bits, axis=1, bitorder="little"),
                        unstable=np.packbits(unstable, axis=1, bitorder="little"), hash_dist=np.array(det.dists),
                        cuts_hash=np.array(cuts, np.int64), thumbs_sha256=np.frombuffer(thumbs.digest(), np.uint8),
                        margin=HASH_MARGIN, versions=json.dumps(VERSIONS))
    print(f"hash_{name}: {n} frames, {len(cuts)} cuts, {int(unstable.sum())} unstable bits of {unstable.size} "
          f"in {time.time() - t0:.1f}s", flush=True)


def nv12_clip(name, seed, w, h, n):
    """NV12 input: the synthetic clip re-expressed as NV12 test content (synth.bgr_to_test_nv12), converted
    by real cv2 (COLOR_YUV2BGR_NV12), downscaled and scored by PySceneDetect's logic -- what the fused NV12 kernel must equal."""
    t0 = time.time()
    sch = synth.build_schedule(seed, n, min_len=20, max_len=70)
    from oracle import closed_form as cf
    dw, dh = cf.downscaled_size(w, h, cf.compute_downscale_factor(w))
    dets = {"content": P.ContentDetector(threshold=27.0, min_scene_len=15), "adaptive": P.AdaptiveDetector(),
            "hist": P.HistogramDetector()}
    cuts = {k: [] for k in dets}
    k = 0
    sha = hashlib.sha256()
    for a in range(0, n, 60):
        nv12 = synth.bgr_to_test_nv12(co.synth_frames(seed, w, h, sch.descs[a:a + 60]))
        sha.update(nv12.tobytes())
        for f in nv12:
            small = cv2.resize(cv2.cvtColor(f, cv2.COLOR_YUV2BGR_NV12), (dw, dh), interpolation=cv2.INTER_LINEAR)
            for nm, d in dets.items():
                cuts[nm] += d.process_frame(k, small)
            k += 1
    np.savez_compressed(os.path.join(HERE, f"nv12_{name}.npz"), seed=seed, width=w, height=h, n_frames=n, dst=np.array([dw, dh]),
                        sums3=np.stack(dets["content"].sums).astype(np.uint64), content_val=np.array(dets["content"].scores),
                        hist_diff=np.array(dets["hist"].diffs), cuts_content=np.array(cuts["content"], np.int64),
                        cuts_adaptive=np.array(cuts["adaptive"], np.int64), cuts_hist=np.array(cuts["hist"], np.int64),
                        input_sha256=np.frombuffer(sha.digest(), np.uint8), versions=json.dumps(VERSIONS))
    print(f"nv12_{name}: {n} frames in {time.time() - t0:.1f}s; cuts {[len(v) for v in cuts.values()]}", flush=True)


def filter_vectors():
    rng = np.random.default_rng(7)
    vecs = []
    patterns = []
    for L in (0, 1, 5, 15):
        for _ in range(6):
            n = 120
            p = rng.random(n) < rng.choice([0.02, 0.1, 0.4])
            patterns.append((L, p.astype(int).tolist()))
    patterns.append((15, [0] * 10 + [1] + [0] * 3 + [1, 1, 1] + [0] * 40 + [1] + [0] * 5 + [1] * 20 + [0] * 30))
    patterns.append((15, [1] * 60))
    patterns.append((15, [0] * 14 + [1] + [0] * 14 + [1] + [0] * 30))
    for L, p in patterns:
        for mode in (P.FILTER_MERGE, P.FILTER_SUPPRESS):
            for start in (0, 1000):
                f = P.FlashFilter(mode, L)
                cuts = []
                for i, a in enumerate(p):
                    cuts += f.filter(start + i, bool(a))
                vecs.append({"length": L, "mode": mode, "start": start, "above": p, "cuts": cuts})
    json.dump({"vectors": vecs, "note": "oracle/psd_cv2.FlashFilter (PySceneDetect 0.6.4 semantics, SURVEY.md A.5)"},
              open(os.path.join(HERE, "filter_vectors.json"), "w"))


if __name__ == "__main__":
    full = "--full" in sys.argv
    only = [a for a in sys.argv[1:] if not a.startswith("--")]
    if not only or "stage" in only: stage_vectors()
    if not only or "cube" in only: cube()
    if not only or "filter" in only: filter_vectors()
    if not only or "clips" in only:
        clip("c1_720p", 1001, 1280, 720, 1800)          # BASELINE config 1, full length
        clip("c2_1080p_head", 1002, 1920, 1080, 600)    # first 600 frames of config 2
        clip("c4_4k_head", 1004, 3840, 2160, 240)       # first 240 frames of config 4
        clip("c2_1080p_int_head", 1002, 1920, 1080, 300, downscale_mode="int")  # <= 0.6.1 downscale (274x154)
    if not only or "hash" in only:
        hash_clip("c1_720p", 1001, 1280, 720, 1800)
        hash_clip("c2_1080p_head", 1002, 1920, 1080, 600)
        hash_clip("c4_4k_head_s8l4", 1004, 3840, 2160, 120, size=8, lowpass=4)
    if not only or "nv12" in only:
        nv12_clip("c2_1080p_head", 1002, 1920, 1080, 240)
    if full:
        clip("c2_1080p_full", 1002, 1920, 1080, 18000)
        clip("c4_4k_full", 1004, 3840, 2160, 3600)
