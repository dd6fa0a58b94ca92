"""CPU ORACLE (test infrastructure only) -- timing harness for the reference's CPU path.

"PySceneDetect with OpenCV, one process per host core": each worker runs oracle/psd_cv2.py's
SceneManager loop (cv2.resize -> detector.process_frame) with cv2.setNumThreads(1) over a shared
sample of frames already decoded in RAM (/dev/shm), exactly the work the GPU path does per frame.
Used only by bench.py (cpu_baseline leg and --impl reference).
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _make_detectors(P, detector, backend):
    if detector == "adaptive":
        return [P.AdaptiveDetector(backend=backend)]
    if detector == "hist":
        return [P.HistogramDetector(backend=backend)]
    return [P.ContentDetector(backend=backend)]


def _worker(args):
    path, detector, reps, start_evt_t = args
    sys.path.insert(0, ROOT)
    from oracle import psd_cv2 as P

    backend = "cv2"
    try:
        import cv2

        cv2.setNumThreads(1)
    except Exception:
        backend = "closed_form"
    frames = np.load(path, mmap_mode="r")
    n = 0
    cuts = None
    t0 = time.perf_counter()
    for _ in range(reps):
        cuts, k = P.detect(frames, _make_detectors(P, detector, backend), backend=backend)
        n += k
    return n, time.perf_counter() - t0, cuts, backend


def _decode_worker(args):
    """Decode + score: cv2.VideoCapture.read() (the reference's decode loop, model_manager.py:237-263) feeding the
    SceneManager loop, one whole file per pass."""
    path, detector, reps, _ = args
    sys.path.insert(0, ROOT)
    import cv2

    from oracle import psd_cv2 as P

    cv2.setNumThreads(1)
    n = 0
    cuts = None
    t0 = time.perf_counter()
    for _ in range(reps):
        cap = cv2.VideoCapture(path)

        def frames():
            while True:
                ok, fr = cap.read()
                if not ok:
                    return
                yield fr

        cuts, k = P.detect(frames(), _make_detectors(P, detector, "cv2"), backend="cv2")
        cap.release()
        n += k
    return n, time.perf_counter() - t0, cuts, "cv2+VideoCapture"


def available_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class Runner:
    """`cores` worker processes (spawned once) that each score every frame of a shared sample `reps` times per step."""

    def __init__(self, frames=None, detector: str = "content", cores: int | None = None, video_path: str | None = None):
        self.cores = cores or available_cores()
        self.detector = detector
        self._own_path = None
        if video_path is not None:
            self.path, self._fn = video_path, _decode_worker
        else:
            shm = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
            self.path = self._own_path = os.path.join(shm, f"esd_cpu_sample_{os.getpid()}.npy")
            np.save(self.path, np.ascontiguousarray(frames))
            self._fn = _worker
        self.pool = mp.get_context("spawn").Pool(self.cores)
        self.pool.map(_noop, range(self.cores))  # start the interpreters (and import cv2) before any clock

    def step(self, reps: int = 1):
        t0 = time.perf_counter()
        res = self.pool.map(self._fn, [(self.path, self.detector, reps, t0)] * self.cores)
        dt = time.perf_counter() - t0
        total = sum(r[0] for r in res)
        return {"frames_per_s": total / dt, "cores": self.cores, "frames_total": total, "seconds": dt, "backend": res[0][3],
                "cuts": res[0][2], "per_core_frames_per_s": float(np.mean([r[0] / r[1] for r in res]))}

    def close(self):
        self.pool.close()
        self.pool.join()
        if self._own_path:
            try:
                os.remove(self._own_path)
            except OSError:
                pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def run(frames: np.ndarray, detector: str = "content", cores: int | None = None, reps: int = 1):
    """Time `cores` processes each scoring every frame of `frames` `reps` times.

    Returns dict(frames_per_s aggregate, cores, frames_total, seconds, backend, cuts)."""
    with Runner(frames, detector, cores) as r:
        return r.step(reps)


def _noop(_):
    import numpy  # noqa: F401

    try:
        import cv2  # noqa: F401
    except Exception:
        pass
    return 0
