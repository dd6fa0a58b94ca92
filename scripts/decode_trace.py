#!/usr/bin/env python
"""Where a decode + scoring session spends its wall time: per session, seconds inside read_batch (host staging + waiting for a
free slot), inside the push, and in the end-of-pass work (cuts D2H + synchronise), for S concurrent sessions."""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))

import torch  # noqa: E402

from decode_probe import make_file  # noqa: E402
from eioku_b200 import capi, decode  # noqa: E402
from eioku_b200.detectors import ContentDetector  # noqa: E402
from eioku_b200.scene_manager import SceneManager  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=768)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--sessions", type=int, default=8)
    ap.add_argument("--passes", type=int, default=3)
    ap.add_argument("--grid-cap", type=int, default=24)
    ap.add_argument("--sync-each-pass", type=int, default=1)
    args = ap.parse_args()
    path = f"/dev/shm/esd_trace_{os.getpid()}.avi"
    make_file(path, args.frames)
    S = args.sessions
    out = [None] * S
    gate = threading.Barrier(S + 1)

    def work(i):
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            sm = SceneManager(batch_frames=args.batch, tuning={"reserved2": args.grid_cap} if args.grid_cap else {})
            sm.add_detector(ContentDetector())
            v = decode.MjpegVideo(path, batch_frames=args.batch)
            ctx = sm.make_context(*v.frame_size)
            t = {"read": 0.0, "push": 0.0, "end": 0.0, "reset": 0.0}
            for p in range(args.passes + 1):
                if p == 1:
                    st.synchronize()
                    gate.wait()
                    t = {k: 0.0 for k in t}
                    t_start = time.perf_counter()
                a = time.perf_counter()
                ctx.reset()
                v.seek(0)
                v._pos = 0
                b = time.perf_counter()
                t["reset"] += b - a
                pos = 0
                while True:
                    a = time.perf_counter()
                    batch = v.read_batch(0)
                    b = time.perf_counter()
                    t["read"] += b - a
                    if batch is None:
                        break
                    ctx.push_tensor(batch, pos)
                    pos += int(batch.shape[0])
                    t["push"] += time.perf_counter() - b
                a = time.perf_counter()
                if args.sync_each_pass:
                    ctx.get_cuts(capi.ESD_DET_CONTENT, 0)
                    ctx.post_process(capi.ESD_DET_CONTENT, pos - 1) if hasattr(ctx, "post_process") else None
                t["end"] += time.perf_counter() - a
            st.synchronize()
            t["wall"] = time.perf_counter() - t_start
            out[i] = {k: round(x, 4) for k, x in t.items()}
            gate.wait()
            v.close()
            ctx.close()

    th = [threading.Thread(target=work, args=(i,)) for i in range(S)]
    for x in th:
        x.start()
    gate.wait()
    t0 = time.perf_counter()
    gate.wait()
    dt = time.perf_counter() - t0
    for x in th:
        x.join()
    os.remove(path)
    print(json.dumps({"sessions": S, "frames_per_s": S * args.passes * args.frames / dt, "seconds": dt, "per_session": out}), flush=True)


if __name__ == "__main__":
    main()
