import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import amatsukaze_b200 as ab
from amatsukaze_b200 import synth
W,H,n=1920,1080,1800
torch.cuda.set_device(0)
lg=synth.make_logo()
t=torch.empty((n,W*H*3//2),dtype=torch.uint8,device="cuda")
for k in range(0,n,20): synth.make_frames(k,min(20,n-k),W,H,device="cuda",logo=lg,imgx=1700,imgy=60,out=t[k:k+20])
clip=ab.yv12_clip(t,W,H,n,True)
res={}
for cw in ("0","1"):
    os.environ["AMTK_EVAL_CW"]=cw
    ctx=ab.Context(0, torch.cuda.current_stream().cuda_stream)
    logo=ab.Logo.create(lg["data"],64,64,W,H,1700,60).deint().create_mask(0.35)
    out=ctx.scan_frames(clip,[logo]); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): out=ctx.scan_frames(clip,[logo])
    e1.record(); torch.cuda.synchronize()
    res[cw]=out.cpu().numpy().view(np.uint32).copy()
    print("AMTK_EVAL_CW=%s: %.4f ms per 1800-frame ScanFrame pass"%(cw,e0.elapsed_time(e1)/20))
    ctx.close()
print("identical bits:", np.array_equal(res["0"],res["1"]))
