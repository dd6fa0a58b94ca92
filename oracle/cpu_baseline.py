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
        if detector == "adaptive":
            dets = [P.AdaptiveDetector(backend=backend)]
        elif detector == "hist":
            dets = [P.HistogramDetector(backend=backend)]
        else:
            dets = [P.ContentDetector(backend=backend)]
        cuts, k = P.detect(frames, dets, backend=backend)
        n += k
    return n, time.perf_counter() - t0, cuts, backend


def available_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run(frames: np.ndarray, detector: str = "content", cores: int | None = None, reps: int = 1):
    """Time `cores` processes each scoring every frame of `frames` `reps` times.

    Returns dict(frames_per_s aggregate, cores, frames_total, seconds, backend, cuts)."""
    cores = cores or available_cores()
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    path = os.path.join(shm, f"esd_cpu_sample_{os.getpid()}.npy")
    np.save(path, np.ascontiguousarray(frames))
    try:
        ctx = mp.get_context("spawn")
        with ctx.Pool(cores) as pool:
            pool.map(_noop, range(cores))  # start the interpreters before the clock
            t0 = time.perf_counter()
            res = pool.map(_worker, [(path, detector, reps, t0)] * cores)
            dt = time.perf_counter() - t0
    finally:
        try:
            os.remove(path)
        except OSError:
            pass
    total = sum(r[0] for r in res)
    return {"frames_per_s": total / dt, "cores": cores, "frames_total": total, "seconds": dt, "backend": res[0][3],
            "cuts": res[0][2], "per_core_frames_per_s": float(np.mean([r[0] / r[1] for r in res]))}


def _noop(_):
    import numpy  # noqa: F401

    try:
        import cv2  # noqa: F401
    except Exception:
        pass
    return 0
