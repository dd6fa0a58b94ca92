"""The reference's OWN validators, executed on this repo's output (CPU suite; skips where /root/reference is absent,
i.e. on the GPU box).

The reference holds no golden vectors for the scene path; what it does hold, and what can execute here
(pydantic is installed, the modules import nothing else), is the output contract:
  * backend/src/domain/schemas/scene_v1.py:7-16            SceneV1 (all >= 0, duration_ms > 0)
  * ml-service/src/models/responses.py:68-73, 135-143      Scene, SceneDetectionResponse
  * backend/src/workers/artifact_transformer.py:65-140     ArtifactTransformer.transform_ml_result("scene_detection")
  * ml-service/src/domain/artifacts.py:7-73                ArtifactEnvelope (__post_init__ validation)
  * backend/tests/test_artifact_transformer.py:393-432     the reference's own example payload (0-5000 / 5000-12500 ms)
The inputs are the cut lists of the committed cv2 goldens (every detector, every BASELINE clip) pushed through the
product's host glue: get_scenes_from_cuts -> scenes_to_dicts -> scene_detection_response -> scene_artifact_envelopes.
"""
import importlib
import importlib.util
import json
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference not present (GPU box)")


def _load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def ref():
    scene_v1 = _load_by_path("ref_scene_v1", f"{REF}/backend/src/domain/schemas/scene_v1.py")
    responses = _load_by_path("ref_responses", f"{REF}/ml-service/src/models/responses.py")
    artifacts = _load_by_path("ref_artifacts", f"{REF}/ml-service/src/domain/artifacts.py")
    # the transformer uses package-relative imports: import it as the reference's own `src` package
    sys.path.insert(0, f"{REF}/backend")
    try:
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        transformer = importlib.import_module("src.workers.artifact_transformer")
    finally:
        sys.path.remove(f"{REF}/backend")
    return {"SceneV1": scene_v1.SceneV1, "Scene": responses.Scene, "Response": responses.SceneDetectionResponse,
            "Envelope": artifacts.ArtifactEnvelope, "Transformer": transformer.ArtifactTransformer}


def _golden_cases():
    """(clip, detector key, cut list, n_frames, fps) for every committed clip golden."""
    out = []
    for name, fps in (("clip_c1_720p.npz", 30.0), ("clip_c2_1080p_head.npz", 30.0), ("clip_c2_1080p_full.npz", 30.0),
                      ("clip_c4_4k_head.npz", 60.0), ("clip_c4_4k_full.npz", 60.0), ("clip_c2_1080p_int_head.npz", 30.0)):
        p = os.path.join(GOLDEN, name)
        if not os.path.exists(p):
            continue
        g = np.load(p)
        for k in g.files:
            if k.startswith("cuts_"):
                out.append((name, k, [int(c) for c in g[k]], int(g["n_frames"]), fps))
    assert out
    return out


def _result_for(cuts, n_frames, fps, video_path="/videos/clip.mp4", config=None):
    from eioku_b200 import service
    from eioku_b200.scene_manager import get_scenes_from_cuts

    scenes = service.scenes_to_dicts(get_scenes_from_cuts(sorted(set(cuts)), 0, n_frames), fps)
    return {"scenes": scenes}, service.scene_detection_response(video_path, config or {"detector": "content"}, scenes, run_id="run_001")


def test_every_golden_scene_list_passes_the_reference_validators(ref):
    from eioku_b200 import service

    n_scenes = 0
    for name, key, cuts, n_frames, fps in _golden_cases():
        result, response = _result_for(cuts, n_frames, fps)
        scenes = result["scenes"]
        assert len(scenes) == len(set(cuts)) + 1, (name, key)
        prev_end = 0
        for i, s in enumerate(scenes):
            v = ref["SceneV1"](**s)  # backend payload schema: ge=0, duration_ms gt=0
            assert v.scene_index == i and v.start_ms == prev_end and v.end_ms - v.start_ms == v.duration_ms, (name, key, s)
            prev_end = v.end_ms
            n_scenes += 1
        assert prev_end == int(n_frames / fps * 1000)
        # ml-service response model: producer defaults to "scenedetect"; extra key duration_ms is not part of Scene
        r = ref["Response"](**response)
        assert r.producer == "scenedetect" and len(r.scenes) == len(scenes)
        assert [(x.scene_index, x.start_ms, x.end_ms) for x in r.scenes] == [(s["scene_index"], s["start_ms"], s["end_ms"]) for s in scenes]
        # the backend's transformer (Redis-result path): needs the provenance keys + scenes with duration_ms
        envs = ref["Transformer"].transform_ml_result(task_id="t1", task_type="scene_detection", video_id="video_001",
                                                      ml_result={**response, "scenes": scenes})
        assert len(envs) == len(scenes)
        assert all(e["artifact_type"] == "scene.detection" for e in envs)
        assert [(e["span_start_ms"], e["span_end_ms"]) for e in envs] == [(s["start_ms"], s["end_ms"]) for s in scenes]
        # the ml-service task handler's envelope (task_handler.py:257-331) as this repo's pure function
        mine = service.scene_artifact_envelopes({**response, "scenes": scenes}, "video_001")
        assert len(mine) == len(scenes)
        for e, s in zip(mine, scenes):
            env = ref["Envelope"](**e)  # __post_init__ validates non-null fields, span order, schema_version
            assert env.artifact_type == "scene" and env.get_duration_ms() == s["duration_ms"]
            assert json.loads(env.payload_json) == s and env.producer == "scenedetect" and env.run_id == "run_001"
            assert env.artifact_id == f"video_001_scene_detection_run_001_{s['scene_index']}"
    assert n_scenes > 500


def test_reference_example_payloads_round_trip(ref):
    """The reference's own literals (scene_v1.py examples; test_artifact_transformer.py:393-432): 150 and 225 frames
    at 30 fps give exactly 0-5000 / 5000-12500 ms through this repo's glue, and the transformer accepts them."""
    from eioku_b200 import service
    from eioku_b200.scene_manager import get_scenes_from_cuts

    examples = ref["SceneV1"].model_config["json_schema_extra"]["examples"]
    scenes = service.scenes_to_dicts(get_scenes_from_cuts([150], 0, 375), 30.0)
    assert scenes == examples
    ml_result = {"config_hash": "config_abc123", "input_hash": "input_xyz789", "run_id": "run_001", "producer": "scenedetect",
                 "producer_version": "0.6.0", "model_profile": "balanced", "scenes": scenes}
    envs = ref["Transformer"].transform_ml_result(task_id="task_006", task_type="scene_detection", video_id="video_001", ml_result=ml_result)
    assert len(envs) == 2 and envs[0]["artifact_type"] == "scene.detection"
    assert (envs[0]["span_start_ms"], envs[0]["span_end_ms"], envs[1]["span_start_ms"], envs[1]["span_end_ms"]) == (0, 5000, 5000, 12500)


def test_envelope_function_follows_the_handlers_defaults_and_drops(ref):
    """task_handler.py:145-153 (defaults when the result carries no provenance) and :277-308 (dropped items)."""
    from eioku_b200 import service

    res = {"scenes": [{"scene_index": 0, "start_ms": 0, "end_ms": 1000, "duration_ms": 1000},
                      {"scene_index": 1, "start_ms": 1000},                                       # no end_ms -> dropped
                      {"scene_index": 2, "start_ms": 3000, "end_ms": 2000, "duration_ms": 1},     # start > end -> dropped
                      {"scene_index": 3, "start_ms": 2000, "end_ms": 2500, "duration_ms": 500}]}
    envs = service.scene_artifact_envelopes(res, "vid", run_id="r")
    assert [e["artifact_id"] for e in envs] == ["vid_scene_detection_r_0", "vid_scene_detection_r_3"]
    for e in envs:
        env = ref["Envelope"](**e)
        assert (env.producer, env.producer_version, env.model_profile, env.config_hash, env.input_hash) == ("ml-service", "1.0.0", "balanced", "", "")


def test_provenance_hashes_equal_the_reference_helpers(tmp_path):
    """ml-service/src/utils/hashing.py:12-54 executed on the same inputs (needs xxhash, a reference dependency)."""
    pytest.importorskip("xxhash")
    from eioku_b200 import service

    hashing = _load_by_path("ref_hashing", f"{REF}/ml-service/src/utils/hashing.py")
    cfg = {"detector": "adaptive", "window_width": 2, "adaptive_threshold": 3.0}
    f = tmp_path / "v.bin"
    f.write_bytes(os.urandom(3 << 20))
    ch, ih = service.provenance_hashes(str(f), cfg)
    assert ch == hashing.compute_config_hash(cfg)
    assert ih == hashing.compute_input_hash(str(f))
