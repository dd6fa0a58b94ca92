#!/usr/bin/env python
"""Host-decoded files (codecs only the CPU can decode here): frames/s of ModelManager.detect_scenes on a 1080p MPEG-4 file with
1 capture (the reference's decode loop) and with several captures decoding frame ranges at once (`decode_workers`), next to the CPU
arm -- cv2.VideoCapture + PySceneDetect logic in ONE process, which is what one job of the reference's worker gets."""
import argparse
import asyncio
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cv2  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402

import synthclip as synth  # noqa: E402
from eioku_b200.service import ModelManager  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=3000)
    ap.add_argument("--workers", default="1,2,4,8,12,16")
    ap.add_argument("--fourcc", default="mp4v")
    args = ap.parse_args()
    w, h, seed = 1920, 1080, 1002
    path = f"/dev/shm/esd_capture_probe_{os.getpid()}.mp4"
    sch = synth.build_schedule(seed, args.frames)
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*args.fourcc), 30.0, (w, h))
    t0 = time.time()
    for a in range(0, args.frames, 64):
        t = torch.empty((min(64, args.frames - a), h, w, 3), dtype=torch.uint8, device="cuda:0")
        synth.fill(t, seed, sch.descs[a:a + 64])
        for f in t.cpu().numpy():
            wr.write(f)
    wr.release()
    out = {"file": {"frames": args.frames, "bytes": os.path.getsize(path), "fourcc": args.fourcc, "encode_s": round(time.time() - t0, 1)}}
    try:
        mm = ModelManager()
        ref = None
        for k in [int(x) for x in args.workers.split(",")]:
            asyncio.run(mm.detect_scenes(path, {"decode_workers": k}))   # warm-up (page cache, CUDA modules)
            t0 = time.perf_counter()
            res = asyncio.run(mm.detect_scenes(path, {"decode_workers": k}))
            dt = time.perf_counter() - t0
            if ref is None:
                ref = res
            out[f"workers_{k}"] = {"frames_per_s": round(args.frames / dt, 1), "seconds": round(dt, 3), "scenes": len(res["scenes"]), "same_scenes": res == ref}
        # the CPU arm for ONE job: decode + PySceneDetect logic in one process
        from oracle import psd_cv2 as P

        cap = cv2.VideoCapture(path)
        det = P.ContentDetector()
        t0 = time.perf_counter()
        k = 0
        while True:
            ok, f = cap.read()
            if not ok:
                break
            det.process_frame(k, cv2.resize(f, (256, 144), interpolation=cv2.INTER_LINEAR))
            k += 1
        dt = time.perf_counter() - t0
        out["cpu_one_process"] = {"frames_per_s": round(k / dt, 1), "seconds": round(dt, 3)}
        cap = cv2.VideoCapture(path)
        t0 = time.perf_counter()
        k = 0
        while cap.read()[0]:
            k += 1
        out["cv2_decode_only_one_capture"] = {"frames_per_s": round(k / (time.perf_counter() - t0), 1)}
    finally:
        os.remove(path)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
