import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eioku_b200 import capi
import synthclip as synth
W,H,n=1920,1080,128
sch=synth.build_schedule(1002,n)
clip=torch.empty((n,H,W,3),dtype=torch.uint8,device="cuda:0"); synth.fill(clip,1002,sch.descs)
stream=torch.cuda.current_stream().cuda_stream
for R,RS,S,occ in [(0,0,0,0),(2,1,4,0),(2,2,3,0),(4,1,4,0),(1,1,4,0),(4,2,2,0)]:
    cfg=capi.default_config(); cfg.src_width,cfg.src_height,cfg.dst_width,cfg.dst_height=W,H,W,H
    cfg.rows_per_group,cfg.rows_per_stage,cfg.pipeline_stages,cfg.ctas_per_sm=R,RS,S,occ
    cfg.initial_capacity=40*n
    ctx=capi.EsdContext(cfg,0)
    pos=0
    for _ in range(2): ctx.push_tensor(clip,pos,stream); pos+=n
    ctx.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ctx.push_tensor(clip,pos,stream); pos+=n
    ctx.join(stream); e1.record(); torch.cuda.synchronize()
    fps=10*n/(e0.elapsed_time(e1)/1e3)
    print(f"R={R} RS={RS} S={S}: {fps:,.0f} frames/s  {fps*W*H*3/1e9:,.0f} GB/s", flush=True)
    ctx.close()
