#!/usr/bin/env python
"""Throughput of the GPU decode path (libesd_decode: nvJPEG, Motion-JPEG AVI) on a 1080p file made from the config-2 clip:
decode only and decode + scoring, per nvJPEG back end and per number of concurrent decoder sessions; and the reference's CPU
arm on the same file (cv2.VideoCapture decode + PySceneDetect logic, one process per core).  One JSON line."""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import synthclip as synth  # noqa: E402
from eioku_b200 import decode  # noqa: E402
from eioku_b200.detectors import ContentDetector  # noqa: E402
from eioku_b200.scene_manager import SceneManager  # noqa: E402


def make_file(path, n, w=1920, h=1080, seed=1002, quality=None):
    import cv2

    sch = synth.build_schedule(seed, n)
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 30.0, (w, h))
    if quality is not None:
        wr.set(cv2.VIDEOWRITER_PROP_QUALITY, quality)
    for a in range(0, n, 64):
        t = torch.empty((min(64, n - a), h, w, 3), dtype=torch.uint8, device="cuda:0")
        synth.fill(t, seed, sch.descs[a:a + 64])
        for f in t.cpu().numpy():
            wr.write(f)
    wr.release()
    return os.path.getsize(path)


def run_sessions(path, sessions, backend, score, batch, passes=2):
    """`sessions` threads, each decoding (and scoring) the whole file `passes` times on its own stream."""
    errs, frames = [], [0] * sessions
    barrier = threading.Barrier(sessions + 1)

    def work(i):
        try:
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                sm = None
                if score:
                    sm = SceneManager(batch_frames=batch, tuning={"reserved2": 24})
                    sm.add_detector(ContentDetector())
                v = decode.MjpegVideo(path, batch_frames=batch, backend=backend)
                # warm-up pass
                while v.read_batch(0) is not None:
                    pass
                st.synchronize()
                barrier.wait()
                for _ in range(passes):
                    v.seek(0)
                    v._pos = 0
                    if sm is not None:
                        frames[i] += sm.detect_scenes(v, reuse_context=True)
                    else:
                        while True:
                            b = v.read_batch(0)
                            if b is None:
                                break
                            frames[i] += int(b.shape[0])
                st.synchronize()
                barrier.wait()
                v.close()
                if sm is not None:
                    sm.close()
        except BaseException as e:  # noqa: BLE001
            errs.append(repr(e))
            try:
                barrier.abort()
            except Exception:
                pass

    th = [threading.Thread(target=work, args=(i,)) for i in range(sessions)]
    for t in th:
        t.start()
    try:
        barrier.wait()
        t0 = time.perf_counter()
        barrier.wait()
        dt = time.perf_counter() - t0
    except threading.BrokenBarrierError:
        dt = float("nan")
    for t in th:
        t.join()
    if errs:
        return {"error": errs[0][:200]}
    return {"frames_per_s": sum(frames) / dt, "seconds": dt}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--backends", default="native,hardware,gpu_hybrid,default")
    ap.add_argument("--sessions", default="1,2,4,8")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    path = os.path.join(shm, f"esd_probe_{os.getpid()}.avi")
    t0 = time.time()
    size = make_file(path, args.frames)
    out = {"file": {"frames": args.frames, "bytes": size, "bytes_per_frame": size // args.frames, "encode_s": time.time() - t0}}
    try:
        with decode.MjpegVideo(path, batch_frames=args.batch) as v:
            out["auto_backend"] = v.backend
            out["hw_engines"] = int(v.info.hw_engines)
        for name, be in (("native", decode.ESD_JPEG_NATIVE), ("hardware", decode.ESD_JPEG_HARDWARE), ("gpu_hybrid", decode.ESD_JPEG_GPU_HYBRID), ("default", decode.ESD_JPEG_DEFAULT)):
            if name not in args.backends.split(","):
                continue
            for s in [int(x) for x in args.sessions.split(",")]:
                for score in (False, True):
                    r = run_sessions(path, s, be, score, args.batch)
                    out[f"{name}_x{s}_{'score' if score else 'decode'}"] = r.get("frames_per_s", r.get("error"))
                    if "error" in r:
                        break
                else:
                    continue
                break
        if not args.no_cpu:
            from oracle import cpu_baseline

            with cpu_baseline.Runner(video_path=path) as runner:
                runner.step(1)
                r = runner.step(1)
            out["cpu_decode_and_score"] = {"frames_per_s": r["frames_per_s"], "cores": r["cores"], "backend": r["backend"], "seconds": r["seconds"]}
    finally:
        os.remove(path)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
