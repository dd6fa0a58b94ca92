"""eioku_b200 -- B200-native scene-detection hot path of codihuston/eioku.

PySceneDetect-compatible ContentDetector / AdaptiveDetector / HistogramDetector behind the
SceneDetector plugin surface and the ml-service scene-task schema, executed by hand-written
sm_100a CUDA kernels in libesd.so (C ABI in include/esd.h).  No CPU fallback.
"""
import os as _os

# A job drives many CUDA streams at once: one per decoder session or device shard, plus each context's own tail / copy streams.
# With the driver's default of 8 hardware work queues unrelated streams share a queue, and a 60 us scoring launch then waits
# behind another session's 35 ms entropy-decode kernel (measured: 24.0 k -> 37.2 k frames/s for 8 decode + scoring sessions,
# profiles/r02_decode_timeline.md).  Read by the driver when the CUDA context is created, so it is set on import, and only if the
# application has not chosen a value itself.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

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
