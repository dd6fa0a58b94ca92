"""eioku_b200 -- B200-native scene-detection hot path of codihuston/eioku.

PySceneDetect-compatible ContentDetector / AdaptiveDetector / HistogramDetector behind the
SceneDetector plugin surface and the ml-service scene-task schema, executed by hand-written
sm_100a CUDA kernels in libesd.so (C ABI in include/esd.h).  No CPU fallback.
"""
from .detectors import (AdaptiveDetector, ContentDetector, FlashFilter, HashDetector, HistogramDetector, SceneDetector,
                        StatsManager, ThresholdDetector)
from .scene_manager import (BatchVideo, SceneManager, TensorVideo, compute_downscale_factor,
                            get_scenes_from_cuts)
from .service import ModelManager, detect, detect_scenes_frames, scenes_to_boundaries, scenes_to_dicts

__all__ = [
    "AdaptiveDetector", "ContentDetector", "FlashFilter", "HashDetector", "HistogramDetector", "SceneDetector", "StatsManager", "ThresholdDetector",
    "BatchVideo", "SceneManager", "TensorVideo", "compute_downscale_factor", "get_scenes_from_cuts",
    "ModelManager", "detect", "detect_scenes_frames", "scenes_to_boundaries", "scenes_to_dicts",
]
__version__ = "0.1.0"
