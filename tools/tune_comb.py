"""Sweep comb-kernel variants (AMTK_COMB_* knobs) on a 1080p clip; prints GB/s per variant.  Also checks parity."""
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import amatsukaze_b200 as ab
from amatsukaze_b200 import synth
if os.environ.get("AMTK_LIB"):
    ab.capi.LIB_PATH = os.environ["AMTK_LIB"]      # codegen experiments: load another build of the library

W, H = 1920, 1080
frames = int(os.environ.get("FRAMES", "600"))
torch.cuda.set_device(0)
clip_t = torch.empty((frames, W * H * 3 // 2), dtype=torch.uint8, device="cuda")
for n0 in range(0, frames, 20):
    n = min(20, frames - n0)
    synth.make_frames(n0, n, W, H, device="cuda", out=clip_t[n0:n0 + n])
clip = ab.yv12_clip(clip_t, W, H, frames, True)
prm = ab.default_comb_params()
ref = None
combos = [(8, 3, 0, 0, 128, 0), (8, 2, 0, 0, 128, 0), (8, 4, 0, 0, 128, 0), (8, 3, 0, 1, 128, 0), (8, 3, 3, 0, 128, 0), (8, 3, 0, 0, 0, 0)]
if len(sys.argv) > 1:
    combos = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
for strip, stages, ctas, acc, l2, R in combos:
    os.environ["AMTK_COMB_STRIP"] = str(strip)
    os.environ["AMTK_COMB_STAGES"] = str(stages)
    os.environ["AMTK_COMB_CTAS"] = str(ctas)
    os.environ["AMTK_COMB_SYNC"] = str(acc)
    os.environ["AMTK_COMB_L2"] = str(l2)
    os.environ["AMTK_COMB_R"] = str(R)
    ctx = ab.Context(0, torch.cuda.current_stream().cuda_stream)
    out = ctx.comb_frames(clip, prm)
    torch.cuda.synchronize()
    ctx.set_kernel_timing(True)
    for _ in range(5):
        out = ctx.comb_frames(clip, prm)
    ms, n = ctx.kernel_timing()
    o = out.cpu().numpy()
    if ref is None:
        ref = o
    gbs = frames * W * H * 1.5 / (ms / n * 1e-3) / 1e9
    print("strip=%d stages=%d ctas=%d sync=%d l2=%d R=%d: %.3f ms/launch  %.0f GB/s  %.0f fps  same=%s" % (strip, stages, ctas, acc, l2, R, ms / n, gbs, frames / (ms / n * 1e-3), np.array_equal(o, ref)), flush=True)
    ctx.close()
