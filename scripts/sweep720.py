"""720p (BASELINE config 1 geometry) pipeline-shape sweep: frames/s of ContentDetector on a resident 1 800-frame clip."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eioku_b200 import capi
import synthclip as synth
W, H, n, seed = int(os.environ.get("W", 1280)), int(os.environ.get("H", 720)), 1800, 1001
sch = synth.build_schedule(seed, n)
clip = torch.empty((n, H, W, 3), dtype=torch.uint8, device="cuda:0")
for a in range(0, n, 300):
    synth.fill(clip[a:a + 300], seed, sch.descs[a:a + 300])
stream = torch.cuda.current_stream().cuda_stream
for (st, rs, rg) in [(0, 0, 0), (2, 4, 16), (3, 4, 16), (4, 4, 16), (3, 2, 16), (4, 2, 16), (6, 2, 16), (3, 4, 8), (4, 3, 16), (3, 3, 16)]:
    cfg = capi.default_config()
    cfg.src_width, cfg.src_height = W, H
    cfg.pipeline_stages, cfg.rows_per_stage, cfg.rows_per_group = st, rs, rg
    cfg.initial_capacity = 30 * n
    try:
        ctx = capi.EsdContext(cfg, 0)
    except capi.EsdError as e:
        print(st, rs, rg, "n/a", str(e)[:80]); continue
    pos = 0
    for _ in range(3):
        ctx.push_tensor(clip, pos, stream); pos += n
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ctx.push_tensor(clip, pos, stream); pos += n
    ctx.join(stream); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    alg = ctx.alg_bytes_per_frame
    print(f"{W}x{H} stages={st} rows/stage={rs} rows/group={rg}: {n/ms*1e3:,.0f} frames/s  {n*alg/ms/1e6:,.0f} GB/s", flush=True)
    ctx.close()
