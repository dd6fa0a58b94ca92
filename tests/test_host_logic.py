"""CPU suite: host-side logic of the product and the C-ABI boundary (no compute calls without a GPU)."""
import asyncio
import ctypes
import io
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from conftest import ROOT
from eioku_b200 import capi, sharding
from eioku_b200.detectors import (AdaptiveDetector, ContentDetector, FlashFilter, HistogramDetector, StatsManager,
                                  ThresholdDetector)
from eioku_b200.scene_manager import SceneManager, compute_downscale_factor, get_scenes_from_cuts
from eioku_b200.service import ModelManager, build_detectors, frame_to_timecode, scenes_to_boundaries, scenes_to_dicts
from oracle import psd_cv2 as P


def _has_gpu():
    import torch

    return torch.cuda.is_available()


# ------------------------------------------------------------------------------------------- C ABI
def test_library_exports_every_declared_symbol(lib_built):
    hdr = open(os.path.join(ROOT, "include", "esd.h")).read()
    declared = sorted(set(re.findall(r"ESD_API\s+[\w\s\*]+?\b(esd_\w+)\s*\(", hdr)))
    assert len(declared) >= 25
    L = capi.load_library()
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(capi.EXPORTED_SYMBOLS) == declared
    assert L.esd_abi_version() == capi.ESD_ABI_VERSION == int(re.search(r"#define ESD_ABI_VERSION (\d+)", hdr).group(1))
    assert L.esd_strerror(-4) == b"call out of order"
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (\w+)", out))
    assert set(declared) <= exported
    assert not [s for s in exported if not s.startswith("esd_")], "only the C ABI may be exported"


def test_ctypes_struct_layout_matches_header(tmp_path):
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "esd.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(esd_config), offsetof(esd_config, content_threshold), offsetof(esd_config, adaptive_weights),"
                   "offsetof(esd_config, hist_bins), offsetof(esd_config, max_cuts), sizeof(esd_geometry));"
                   'printf("%zu %zu\\n", offsetof(esd_config, hash_threshold), offsetof(esd_config, hash_min_scene_len));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = list(map(int, subprocess.check_output([str(exe)]).split()))
    C = capi.EsdConfig
    assert got == [ctypes.sizeof(C), C.content_threshold.offset, C.adaptive_weights.offset, C.hist_bins.offset,
                   C.max_cuts.offset, ctypes.sizeof(capi.EsdGeometry), C.hash_threshold.offset, C.hash_min_scene_len.offset]


def test_defaults_are_pyscenedetect_defaults():
    cfg = capi.default_config()
    assert (cfg.content_threshold, cfg.content_min_scene_len, cfg.content_filter_mode) == (27.0, 15, 0)
    assert list(cfg.content_weights) == [1.0, 1.0, 1.0, 0.0] and cfg.content_weight_div == 3.0
    assert (cfg.adaptive_threshold, cfg.adaptive_window_width, cfg.adaptive_min_content_val, cfg.adaptive_min_scene_len) == (3.0, 2, 15.0, 15)
    assert (cfg.hist_threshold, cfg.hist_bins, cfg.hist_min_scene_len) == (0.05, 256, 15)
    assert (cfg.hash_threshold, cfg.hash_size, cfg.hash_lowpass, cfg.hash_min_scene_len) == (0.395, 16, 2, 15)


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly():
    cfg = capi.default_config()
    cfg.src_width, cfg.src_height = 64, 48
    with pytest.raises(capi.EsdError) as e:
        capi.EsdContext(cfg, 0)
    assert "no CUDA device" in str(e.value) and "no CPU fallback" in str(e.value)
    assert capi.load_library().esd_device_count() == 0
    det = ContentDetector()
    with pytest.raises(Exception):
        det.process_frame(0, np.zeros((48, 64, 3), np.uint8))


def test_missing_library_is_an_import_error(tmp_path):
    with pytest.raises(ImportError):
        capi.load_library(str(tmp_path / "libesd.so"))


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "eioku_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")) and f != "synth_core.h":
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "esd_oracle" not in text and "libesd_oracle" not in text, f


# ------------------------------------------------------------------------------------------- plugin surface
def test_detector_constructors_and_metrics():
    d = ContentDetector()
    assert d.get_metrics() == ["content_val", "delta_hue", "delta_sat", "delta_lum", "delta_edges"]
    assert d.event_buffer_length == 0 and d.post_process(10) == [] and d.is_processing_required(3)
    assert ContentDetector(luma_only=True)._weights == ContentDetector.LUMA_ONLY_WEIGHTS
    a = AdaptiveDetector(window_width=3, luma_only=True)
    assert a.event_buffer_length == 3 and a.get_metrics()[-1] == "adaptive_ratio_lum (w=3)"
    assert AdaptiveDetector().get_metrics()[-1] == "adaptive_ratio (w=2)"
    assert AdaptiveDetector(min_delta_hsv=9.0).min_content_val == 9.0
    h = HistogramDetector(threshold=0.3, bins=64)
    assert h.get_metrics() == ["hist_diff [bins=64]"] and h._threshold == 0.7
    assert HistogramDetector(threshold=2.0)._threshold == 0.0
    with pytest.raises(ValueError):
        AdaptiveDetector(window_width=0)
    e = ContentDetector(weights=ContentDetector.Components(1.0, 1.0, 1.0, 0.5), kernel_size=7)
    ecfg = capi.default_config()
    e._fill_config(ecfg)
    assert ecfg.edge_kernel_size == 7 and ecfg.content_weights[3] == 0.5
    with pytest.raises(ValueError):
        AdaptiveDetector(weights=ContentDetector.Components(1.0, 1.0, 1.0, 0.5), kernel_size=5)._fill_config(ecfg)
    with pytest.raises(ValueError):
        ContentDetector(kernel_size=4)
    cfg = capi.default_config()
    cfg.detectors = 0
    ContentDetector(threshold=31.5, min_scene_len=7, weights=ContentDetector.Components(0.1, 0.2, 0.3, 0.0),
                    filter_mode=FlashFilter.Mode.SUPPRESS)._fill_config(cfg)
    AdaptiveDetector(adaptive_threshold=2.5, window_width=4, luma_only=True)._fill_config(cfg)
    HistogramDetector(threshold=0.1, bins=32, min_scene_len=3)._fill_config(cfg)
    t = ThresholdDetector(threshold=12.9, fade_bias=-0.5, add_final_scene=True, method=ThresholdDetector.Method.CEILING)
    assert t.threshold == 12 and t.get_metrics() == ["average_rgb"] and t.post_process(5) == []
    t._fill_config(cfg)
    assert (cfg.thresh_threshold, cfg.thresh_fade_bias, cfg.thresh_add_final_scene, cfg.thresh_method) == (12.0, -0.5, 1, 1)
    assert cfg.detectors == 15 and cfg.content_threshold == 31.5 and cfg.content_filter_mode == 1
    # the divisor is the host interpreter's sum(abs(w)) -- same expression as PySceneDetect
    assert cfg.content_weight_div == sum(abs(w) for w in (0.1, 0.2, 0.3, 0.0))
    assert list(cfg.adaptive_weights) == [0.0, 0.0, 1.0, 0.0] and cfg.adaptive_window_width == 4
    assert (cfg.hist_threshold, cfg.hist_bins, cfg.hist_min_scene_len) == (0.1, 32, 3)


def test_scene_manager_geometry_and_scene_lists():
    assert compute_downscale_factor(1920) == 7.5 and compute_downscale_factor(1920, mode="int") == 7
    assert compute_downscale_factor(200) == 1
    sm = SceneManager()
    assert sm._target_size(1920, 1080) == (256, 144) and sm._target_size(3840, 2160) == (256, 144)
    assert sm._target_size(255, 100) == (255, 100)
    assert SceneManager(downscale_mode="int")._target_size(1920, 1080) == (274, 154)
    sm.downscale = 2
    assert not sm.auto_downscale and sm._target_size(1920, 1080) == (960, 540)
    with pytest.raises(ValueError):
        sm.downscale = 0
    sm.add_detector(ContentDetector())
    with pytest.raises(ValueError):
        sm.add_detector(ContentDetector())
    for cuts, s, e in (([], 0, 100), ([10, 50], 0, 100), ([99], 7, 100)):
        assert get_scenes_from_cuts(cuts, s, e) == P.get_scenes_from_cuts(cuts, s, e)
    assert get_scenes_from_cuts([10, 50], 0, 100) == [(0, 10), (10, 50), (50, 100)]
    assert sm.get_scene_list() == []


def test_scene_dicts_match_reference_schema():
    scenes = [(0, 150), (150, 375), (375, 376)]
    d = scenes_to_dicts(scenes, 30.0)
    assert d == P.scenes_to_dicts(scenes, 30.0)
    assert d[0] == {"scene_index": 0, "start_ms": 0, "end_ms": 5000, "duration_ms": 5000}
    assert d[1] == {"scene_index": 1, "start_ms": 5000, "end_ms": 12500, "duration_ms": 7500}  # scene_v1.py examples
    for s in d:  # SceneV1: all >= 0, duration > 0; task_handler: start <= end, both present
        assert set(s) == {"scene_index", "start_ms", "end_ms", "duration_ms"}
        assert s["duration_ms"] > 0 and s["end_ms"] >= s["start_ms"] >= 0
        assert all(isinstance(v, int) for v in s.values())
    assert scenes_to_dicts([(0, 1001)], 29.97)[0]["end_ms"] == int(1001 / 29.97 * 1000)


def test_scene_boundaries_timecodes():
    assert frame_to_timecode(0, 30.0) == "00:00:00.000"
    assert frame_to_timecode(150, 30.0) == "00:00:05.000"
    assert frame_to_timecode(1001, 29.97) == "00:00:33.400"
    assert frame_to_timecode(30 * 3600 + 30 * 61 + 7, 30.0) == "01:01:01.233"
    assert frame_to_timecode(1799, 30.0) == "00:00:59.967"
    assert frame_to_timecode(215999, 60.0) == "00:59:59.983"
    assert scenes_to_boundaries([(0, 150), (150, 375)], 30.0) == [
        {"scene": 0, "start": "00:00:00.000", "end": "00:00:05.000"}, {"scene": 1, "start": "00:00:05.000", "end": "00:00:12.500"}]


def test_provenance_hashes_and_response(tmp_path):
    xxhash = pytest.importorskip("xxhash")
    import json

    from eioku_b200.service import provenance_hashes, scene_detection_response

    cfg = {"threshold": 27.0, "detector": "content", "min_scene_len": 15}
    f = tmp_path / "v.bin"
    f.write_bytes(bytes(range(256)) * 4099)
    ch, ih = provenance_hashes(str(f), cfg)
    assert ch == xxhash.xxh64(json.dumps(cfg, sort_keys=True).encode()).hexdigest()[:16]
    assert ih == xxhash.xxh64(f.read_bytes()).hexdigest()[:16]
    assert provenance_hashes("/no/such/file.mp4", {})[1] == xxhash.xxh64(b"/no/such/file.mp4").hexdigest()[:16]
    scenes = scenes_to_dicts([(0, 150), (150, 375)], 30.0)
    r = scene_detection_response(str(f), cfg, scenes, run_id="r1")
    assert set(r) == {"run_id", "config_hash", "input_hash", "producer", "producer_version", "scenes"}  # responses.py:135-143
    assert r["producer"] == "scenedetect" and r["run_id"] == "r1" and r["config_hash"] == ch and r["input_hash"] == ih
    assert r["scenes"] == [{"scene_index": 0, "start_ms": 0, "end_ms": 5000}, {"scene_index": 1, "start_ms": 5000, "end_ms": 12500}]


def test_build_detectors_from_task_config():
    legacy = build_detectors({"threshold": 0.7, "min_scene_length": 0.6})  # video_discovery_service.py:425-428
    assert len(legacy) == 1 and isinstance(legacy[0], ContentDetector) and legacy[0]._threshold == 27.0
    d = build_detectors({"detector": "content", "threshold": 30, "min_scene_len": 20, "luma_only": True, "filter_mode": "suppress"})[0]
    assert (d._threshold, d._min_scene_len, d._weights, d._filter_mode) == (30.0, 20, ContentDetector.LUMA_ONLY_WEIGHTS, FlashFilter.Mode.SUPPRESS)
    a = build_detectors({"detector": "adaptive", "adaptive_threshold": 2.0, "window_width": 3, "min_content_val": 10})[0]
    assert isinstance(a, AdaptiveDetector) and (a.adaptive_threshold, a.window_width, a.min_content_val) == (2.0, 3, 10.0)
    both = build_detectors({"detector": "hist+content", "bins": 128, "hist_threshold": 0.1})
    assert isinstance(both[0], HistogramDetector) and both[0]._bins == 128 and isinstance(both[1], ContentDetector)
    th = build_detectors({"detector": "threshold", "threshold": 20, "fade_bias": 0.5, "add_final_scene": True})[0]
    assert isinstance(th, ThresholdDetector) and (th.threshold, th.fade_bias, th.add_final_scene) == (20, 0.5, True)
    hs = build_detectors({"detector": "hash", "threshold": 0.3, "size": 8, "lowpass": 4, "min_scene_len": 10})[0]
    assert type(hs).__name__ == "HashDetector" and (hs._threshold, hs._size, hs._factor, hs._min_scene_len) == (0.3, 8, 4, 10)
    assert hs.get_metrics() == ["hash_dist [size=8 lowpass=4]"]
    with pytest.raises(ValueError):
        build_detectors({"detector": "optical-flow"})


def test_model_manager_raises_like_the_reference():
    mm = ModelManager(decoder=lambda path, cfg: (_ for _ in ()).throw(RuntimeError("Failed to open video: " + path)))
    with pytest.raises(RuntimeError):
        asyncio.run(mm.detect_scenes("/nonexistent.mp4", {}))


def test_stats_manager_csv():
    sm = StatsManager()
    sm.set_metrics(0, {"content_val": 0.0})
    sm.set_metrics(1, {"content_val": 1.5, "delta_hue": 2.0})
    assert sm.metrics_exist(1, ["content_val", "delta_hue"]) and not sm.metrics_exist(0, ["delta_hue"])
    assert sm.get_metrics(1, ["delta_hue"]) == [2.0]
    buf = io.StringIO()
    sm.save_to_csv(buf)
    lines = buf.getvalue().strip().split("\n")
    # scenedetect's stats-file layout: Frame Number (1-based), Timecode, then the metric keys sorted; str() per value
    assert lines[0] == "Frame Number,Timecode,content_val,delta_hue"
    assert lines[1] == "1,00:00:00.000,0.0,None" and lines[2] == "2,00:00:00.033,1.5,2.0"
    buf = io.StringIO()
    sm.save_to_csv(buf, fps=25.0)
    assert buf.getvalue().strip().split("\n")[2] == "2,00:00:00.040,1.5,2.0"


# ------------------------------------------------------------------------------------------- sharding
def test_frame_range_shards_and_halo():
    sh = sharding.frame_range_shards(18000, 8, window_width=2)
    assert [s.own_end - s.own_start for s in sh] == [2250] * 8
    assert sh[0].load_start == 0 and sh[0].load_end == 2252
    assert sh[3].load_start == sh[3].own_start - 3 and sh[3].load_end == sh[3].own_end + 2  # window_width+1 / window_width
    assert sh[7].load_end == 18000
    assert sharding.frame_range_shards(5, 8, 1)[0].own_end in (0, 1)
    tiny = sharding.frame_range_shards(3, 8, 2)
    assert sum(s.own_end - s.own_start for s in tiny) == 3


def test_partition_videos_lpt():
    lengths = [9000] * 512
    parts = sharding.partition_videos(lengths, 8)
    assert sorted(sum(parts, [])) == list(range(512)) and all(len(p) == 64 for p in parts)
    parts = sharding.partition_videos([10, 1, 1, 1, 7, 3, 3], 3)
    loads = [sum([10, 1, 1, 1, 7, 3, 3][i] for i in p) for p in parts]
    assert max(loads) == 10 and sorted(sum(parts, [])) == list(range(7))


def test_merge_owned_checks_tiling():
    sh = sharding.frame_range_shards(10, 2, 1)
    parts = [{"content_val": np.arange(5.0)}, {"content_val": np.arange(5.0, 10.0)}]
    m = sharding.merge_owned(parts, sh)
    assert np.array_equal(m["content_val"], np.arange(10.0))
    with pytest.raises(ValueError):
        sharding.merge_owned(parts, [sh[0], sharding.FrameShard(1, 6, 10, 4, 10)])


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    from eioku_b200 import sharding as sh
    import synthclip as synth
    from oracle import c_oracle as co
    from oracle import psd_cv2 as PP

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, w, h, seed, ww = 160, 320, 180, 21, 2
    sch = synth.build_schedule(seed, n, min_len=12, max_len=40)

    def score_shard(s):
        frames = co.synth_frames(seed, w, h, sch.descs[s.load_start:s.load_end])
        det = PP.AdaptiveDetector(window_width=ww, min_scene_len=5, backend="closed_form")
        hd = PP.HashDetector(threshold=0.3, min_scene_len=5, backend="closed_form")
        for k, f in enumerate(frames):
            det.process_frame(s.load_start + k, f)
            hd.process_frame(s.load_start + k, f)
        return {"adaptive_val": np.array(det.scores), "sums3": np.stack(det.sums), "hash_dist": np.array(hd.dists)}

    def decide(m):
        # single global pass over the concatenated scores with the oracle's state machine
        class Replay(PP.AdaptiveDetector):
            def _calculate_frame_score(self, frame_num, frame_img):
                return np.float64(frame_img) if frame_num > 0 else 0.0
        det = Replay(window_width=ww, min_scene_len=5, backend="closed_form")
        cuts = []
        for k, v in enumerate(m["adaptive_val"]):
            cuts += det.process_frame(k, v)
        hcuts, last = [], 0   # HashDetector's rule on the merged distances (first frame: NaN, never a cut)
        for k, v in enumerate(m["hash_dist"]):
            if v >= 0.3 and k - last >= 5:
                hcuts.append(k)
                last = k
        return {"adaptive": cuts, "hash": hcuts, "hash_dist0": float(m["hash_dist"][0]), "n": len(m["adaptive_val"]),
                "sum0": m["sums3"][0].tolist()}

    out = sh.sharded_detect(score_shard, decide, n, window_width=ww)
    if rank == 0:
        frames = co.synth_frames(seed, w, h, sch.descs)
        ref = PP.AdaptiveDetector(window_width=ww, min_scene_len=5, backend="closed_form")
        want, _ = PP.detect(frames, [ref], backend="closed_form", auto_downscale=False)
        href = PP.HashDetector(threshold=0.3, min_scene_len=5, backend="closed_form")
        hwant, _ = PP.detect(frames, [href], backend="closed_form", auto_downscale=False)
        q.put((out, want, hwant))
    dist.destroy_process_group()


def test_sharded_detect_gloo_world2():
    """N>1 host path on CPU: two gloo ranks score halo'd frame ranges with the oracle, rank 0 merges and
    runs one global decision pass; must equal the single-process result."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out, want, hwant = q.get(timeout=300)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert out["n"] == 160 and out["sum0"] == [0, 0, 0]
    assert out["adaptive"] == want and len(want) >= 2
    assert out["hash"] == hwant and len(hwant) >= 1 and np.isnan(out["hash_dist0"])


# ------------------------------------------------------------------------------------------- later round-1 additions
def test_hash_detector_fills_config_and_metric_key():
    from eioku_b200.detectors import HashDetector

    cfg = capi.default_config()
    cfg.detectors = 0
    d = HashDetector(threshold=0.3, size=8, lowpass=4, min_scene_len=9)
    d._fill_config(cfg)
    assert cfg.detectors == capi.ESD_DET_HASH
    assert (cfg.hash_threshold, cfg.hash_size, cfg.hash_lowpass, cfg.hash_min_scene_len) == (0.3, 8, 4, 9)
    assert d.get_metrics() == ["hash_dist [size=8 lowpass=4]"] and d.event_buffer_length == 0
    assert d.defer(32).event_buffer_length == 31
    a = AdaptiveDetector(window_width=3).defer(8)
    assert a.event_buffer_length == 3 + 7


def test_nv12_video_geometry_and_test_content():
    import synthclip as synth
    from eioku_b200.scene_manager import TensorVideo

    rng = np.random.default_rng(0)
    bgr = rng.integers(0, 256, (3, 8, 12, 3), dtype=np.uint8)
    nv12 = synth.bgr_to_test_nv12(bgr)
    assert nv12.shape == (3, 12, 12)
    assert np.array_equal(nv12[:, :8], bgr[..., 1])
    assert np.array_equal(nv12[:, 8:, 0::2], bgr[:, 0::2, 0::2, 0]) and np.array_equal(nv12[:, 8:, 1::2], bgr[:, 0::2, 0::2, 2])
    v = TensorVideo(nv12, 25.0, pixel_format="nv12")
    assert v.frame_size == (12, 8) and v.read_batch(2).shape == (2, 12, 12)
    with pytest.raises(ValueError):
        TensorVideo(bgr, pixel_format="nv12")
    with pytest.raises(ValueError):
        TensorVideo(nv12, pixel_format="yuyv")
    # planar I420 carries the same samples: the (H*3/2, W) array cv2.cvtColor(COLOR_YUV2BGR_I420) takes
    i420 = synth.nv12_to_i420(nv12)
    assert i420.shape == nv12.shape and np.array_equal(i420[:, :8], nv12[:, :8])
    assert np.array_equal(i420[:, 8:10].reshape(3, 4, 6), nv12[:, 8:, 0::2]) and np.array_equal(i420[:, 10:].reshape(3, 4, 6), nv12[:, 8:, 1::2])
    assert TensorVideo(i420, pixel_format="i420").frame_size == (12, 8)
    # the closed-form NV12 -> BGR conversion agrees with cv2 on this content too
    cv2 = pytest.importorskip("cv2")
    from oracle import closed_form as cf

    assert np.array_equal(cf.nv12_to_bgr_u8(nv12[0]), cv2.cvtColor(nv12[0], cv2.COLOR_YUV2BGR_NV12))
    # and cv2 converts the planar form to the very same picture: one arithmetic, two chroma layouts
    assert np.array_equal(cv2.cvtColor(i420[0], cv2.COLOR_YUV2BGR_I420), cv2.cvtColor(nv12[0], cv2.COLOR_YUV2BGR_NV12))


def test_scene_artifact_payloads_from_a_finished_manager():
    """SceneV1 {scene_index, method, score, frame_number} (reference spec, artifact-envelope design.md:159-167)."""
    from eioku_b200.detectors import HashDetector
    from eioku_b200.scene_manager import SceneManager
    from eioku_b200.service import scene_artifact_payloads

    sm = SceneManager.__new__(SceneManager)   # no device needed: fill in what detect_scenes leaves behind
    c, h = ContentDetector(), HashDetector()
    sm._detector_list = [c, h]
    sm._start_pos, sm._last_pos = 100, 199
    sm._cuts_by_detector = {"ContentDetector": [130, 160], "HashDetector": [130, 181]}
    sm.scores = {"content_val": np.arange(100, dtype=np.float64), "hash_dist": np.full(100, 0.5)}
    recs = scene_artifact_payloads(sm)
    assert recs == [
        {"scene_index": 0, "method": "start", "score": 0.0, "frame_number": 100},
        {"scene_index": 1, "method": "content", "score": 30.0, "frame_number": 130},
        {"scene_index": 2, "method": "content", "score": 60.0, "frame_number": 160},
        {"scene_index": 3, "method": "hash", "score": 0.5, "frame_number": 181},
    ]
    sm._start_pos = None
    assert scene_artifact_payloads(sm) == []


# ------------------------------------------------------------------------------------------- round 2: decode library boundary
def test_decode_library_exports_every_declared_symbol(lib_built):
    from eioku_b200 import decode

    hdr = open(os.path.join(ROOT, "include", "esd_decode.h")).read()
    declared = sorted(set(re.findall(r"ESD_DEC_API\s+[\w\s\*]+?\b(esd_\w+)\s*\(", hdr)))
    assert sorted(decode.EXPORTED_SYMBOLS) == declared and len(declared) == 7
    L = decode.load_library()
    for name in declared:
        assert hasattr(L, name), name
    assert L.esd_decode_abi_version() == decode.ESD_DECODE_ABI_VERSION == int(re.search(r"#define ESD_DECODE_ABI_VERSION (\d+)", hdr).group(1))
    out = subprocess.run(["nm", "-D", "--defined-only", decode.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (\w+)", out))
    assert set(declared) <= exported and not [s for s in exported if not s.startswith("esd_")]
    src = "#include <stddef.h>\n#include <stdio.h>\n#include \"esd_decode.h\"\nint main(void){printf(\"%zu %zu %zu\\n\", sizeof(esd_mjpeg_info), offsetof(esd_mjpeg_info, n_frames), offsetof(esd_mjpeg_info, backend));return 0;}\n"
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "l.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "l.c"), "-o", os.path.join(d, "l")])
        got = list(map(int, subprocess.check_output([os.path.join(d, "l")]).split()))
    assert got == [ctypes.sizeof(decode.MjpegInfo), decode.MjpegInfo.n_frames.offset, decode.MjpegInfo.backend.offset]


def test_mjpeg_sniffing_and_no_cpu_fallback(tmp_path):
    cv2 = pytest.importorskip("cv2")
    from eioku_b200 import decode

    frames = np.random.default_rng(0).integers(0, 255, (4, 48, 64, 3), dtype=np.uint8)
    for fourcc, want in (("MJPG", True), ("mp4v", False)):
        p = str(tmp_path / f"{fourcc}.avi")
        w = cv2.VideoWriter(p, cv2.VideoWriter_fourcc(*fourcc), 30.0, (64, 48))
        for f in frames:
            w.write(f)
        w.release()
        assert decode.is_mjpeg_avi(p) is want
    assert not decode.is_mjpeg_avi(str(tmp_path / "nope.avi"))
    if not _has_gpu():
        with pytest.raises(decode.DecodeError) as e:
            decode.MjpegVideo(str(tmp_path / "MJPG.avi"))
        assert "no CPU fallback" in str(e.value)


def test_all_gather_scores_single_rank_and_schema_packing():
    import torch

    local = torch.arange(12, dtype=torch.float64).reshape(2, 6)
    assert torch.equal(sharding.all_gather_scores(local, [4]), local[:, :4])
    schema = [("adaptive_val", 1, np.float64), ("sums3", 3, np.uint64)]
    part = {"adaptive_val": np.array([0.5, np.nan, 2.0]), "sums3": np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9400320]], np.uint64)}
    m = sharding._pack(part, schema, 3, 5)
    assert tuple(m.shape) == (4, 5)
    back = sharding._unpack(m[:, :3], schema)
    assert back["sums3"].dtype == np.uint64 and np.array_equal(back["sums3"], part["sums3"])
    assert np.array_equal(np.nan_to_num(back["adaptive_val"], nan=-1), np.nan_to_num(part["adaptive_val"], nan=-1))
    with pytest.raises(ValueError):
        sharding._pack({"x": np.array([1 << 60], np.uint64)}, [("x", 1, np.uint64)], 1, 1)
