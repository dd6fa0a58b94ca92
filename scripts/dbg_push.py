import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eioku_b200 import capi, synth
W,H,n=1920,1080,1024
dev="cuda:0"
sch=synth.build_schedule(1002, 20000)
cfg=capi.default_config(); cfg.src_width, cfg.src_height = W,H
ctx=capi.EsdContext(cfg,0)
ctx.set_timing(True)
stream=torch.cuda.current_stream().cuda_stream
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
def fill(a):
    out=torch.empty((n,H,W,3),dtype=torch.uint8,device=dev)
    for b in range(0,n,256): capi.synth_fill(out[b:b+256],1002,sch.descs[a+b:a+b+256])
    return out
clip=fill(0)
pos=0
for mode in ["same","same","same","fresh","fresh","fresh","same_nosync","same_nosync"]:
    if mode=="fresh": clip=fill(pos)
    if mode!="same_nosync": torch.cuda.synchronize()
    t0=time.perf_counter()
    e0.record(); ctx.push_tensor(clip,pos,stream); ctx.join(stream); e1.record(); 
    t1=time.perf_counter()
    torch.cuda.synchronize()
    km,kn=ctx.kernel_time()
    print(mode, "event ms %.3f"%e0.elapsed_time(e1), "fused ms %.3f"%km, "host enqueue ms %.3f"%((t1-t0)*1e3), flush=True)
    pos+=n
