"""Host tap-gather ingest rates (frames/s) for a few thread counts / ring shapes; ESD_GATHER_PF=<bytes> selects the
rolling software-prefetch distance of the gather loop (0 = the next-row page touch)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from eioku_b200 import capi
import synthclip as synth
W, H, n = 1920, 1080, 512
sch = synth.build_schedule(1002, n)
clip = torch.empty((n, H, W, 3), dtype=torch.uint8, device="cuda:0"); synth.fill(clip, 1002, sch.descs)
host_pinned = clip.cpu().pin_memory().numpy()
cfg = capi.default_config(); cfg.src_width, cfg.src_height = W, H
ref = capi.EsdContext(cfg, 0); ref.push_tensor(clip, 0); want = ref.read_scores(0, n, ["sums3"])["sums3"]; ref.close()
shapes = [(4, 128)] if os.environ.get("ESD_PROBE_SHORT") else [(4, 128), (4, 64), (6, 64)]
for slots, fps in shapes:
    for th in (0, 16, 12, 8):
        ctx = capi.EsdContext(cfg, 0); ctx.ingest_open(slots, fps); ctx.ingest_set_gather(th)
        for i in range(2): ctx.ingest_push_numpy(host_pinned, i * n)
        ctx.synchronize()
        t0 = time.perf_counter()
        for i in range(2, 6): ctx.ingest_push_numpy(host_pinned, i * n)
        ctx.synchronize(); dt = time.perf_counter() - t0
        ok = np.array_equal(ctx.read_scores(0, n, ["sums3"])["sums3"], want)
        print(f"pf={os.environ.get('ESD_GATHER_PF','-')} slots={slots} frames/slot={fps} gather_threads={th}: {4*n/dt:,.0f} frames/s  "
              f"h2d bytes/frame={ctx.ingest_stats()[0]/(6*n):,.0f} exact={ok}", flush=True)
        ctx.close()
