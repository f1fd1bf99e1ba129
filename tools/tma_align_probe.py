import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import amatsukaze_b200 as ab
from amatsukaze_b200 import synth
imgx = int(sys.argv[1]); w = int(sys.argv[2]) if len(sys.argv) > 2 else 512
h = 128
lg = synth.make_logo(64, 64)
fr = synth.make_frames(0, 4, w, h, device="cuda")
ctx = ab.Context(0, torch.cuda.current_stream().cuda_stream)
logo = ab.Logo.create(lg["data"], 64, 64, w, h, imgx, 20).deint().create_mask(0.35)
try:
    out = ctx.scan_frames(ab.yv12_clip(fr, w, h, 4, True), [logo])
    torch.cuda.synchronize()
    print("imgx", imgx, "ok", out[0].cpu().numpy())
except Exception as e:
    print("imgx", imgx, "FAILED", str(e)[:120])
