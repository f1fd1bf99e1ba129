"""Equal-share vs lock-step partition of the comb kernel on 1920x1080 and 1440x1080 clips (GB/s algorithmic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import amatsukaze_b200 as ab
from amatsukaze_b200 import synth
torch.cuda.set_device(0)
for (w, h) in ((1920, 1080), (1440, 1080)):
    n = 900
    t = torch.empty((n, w * h * 3 // 2), dtype=torch.uint8, device="cuda")
    for n0 in range(0, n, 20):
        synth.make_frames(n0, min(20, n - n0), w, h, device="cuda", out=t[n0:n0 + min(20, n - n0)])
    clip = ab.yv12_clip(t, w, h, n, True)
    ref = None
    for part, l2 in ((0, 128), (1, 128), (0, 64), (1, 64), (0, 0)):
        os.environ["AMTK_COMB_PART"] = str(part)
        os.environ["AMTK_COMB_L2"] = str(l2)
        ctx = ab.Context(0, torch.cuda.current_stream().cuda_stream)
        out = ctx.comb_frames(clip)
        torch.cuda.synchronize()
        ctx.set_kernel_timing(True)
        for _ in range(5):
            out = ctx.comb_frames(clip)
        ms, k = ctx.kernel_timing()
        o = out.cpu().numpy()
        ref = o if ref is None else ref
        print("%dx%d part=%d l2=%d: %.3f ms  %.0f GB/s  same=%s" % (w, h, part, l2, ms / k, n * w * h * 1.5 / (ms / k) / 1e6, np.array_equal(o, ref)), flush=True)
        ctx.close()
    del t
