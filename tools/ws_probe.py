"""Round-2 comb kernel probe: times the warp-stream kernel (AMTK_COMB_WS=1) against the round-1 CTA-ring kernel
(AMTK_COMB_WS=0) on resident clips and checks that both return identical counters.  Usage:
    python tools/ws_probe.py [WxHxFRAMES ...]      (default 1920x1080x1800 1440x1080x1800)
Extra env knobs are passed through (AMTK_COMB_R, AMTK_COMB_CTAS, AMTK_COMB_L2, AMTK_COMB_STAGES)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import amatsukaze_b200 as ab
from amatsukaze_b200 import synth

if os.environ.get("AMTK_LIB"):
    ab.capi.LIB_PATH = os.environ["AMTK_LIB"]
specs = sys.argv[1:] or ["1920x1080x1800", "1440x1080x1800"]
torch.cuda.set_device(0)
for spec in specs:
    W, H, frames = (int(x) for x in spec.split("x"))
    clip_t = torch.empty((frames, W * H * 3 // 2), dtype=torch.uint8, device="cuda")
    for n0 in range(0, frames, 20):
        n = min(20, frames - n0)
        synth.make_frames(n0, n, W, H, device="cuda", out=clip_t[n0:n0 + n])
    clip = ab.yv12_clip(clip_t, W, H, frames, True)
    prm = ab.default_comb_params()
    ref = None
    for ws in (os.environ.get("WS_LIST", "0,1").split(",")):
        os.environ["AMTK_COMB_WS"] = ws
        ctx = ab.Context(0, torch.cuda.current_stream().cuda_stream)
        out = ctx.comb_frames(clip, prm)
        torch.cuda.synchronize()
        ctx.set_kernel_timing(True)
        for _ in range(int(os.environ.get("REPS", "10"))):
            out = ctx.comb_frames(clip, prm)
        ms, n = ctx.kernel_timing()
        o = out.cpu().numpy()
        if ref is None:
            ref = o
        gbs = frames * W * H * 1.5 / (ms / n * 1e-3) / 1e9
        print("%s ws=%s: %.4f ms/launch  %.0f GB/s  %.0f fps  same=%s" % (spec, ws, ms / n, gbs, frames / (ms / n * 1e-3), np.array_equal(o, ref)), flush=True)
        if not np.array_equal(o, ref):
            bad = np.argwhere(o != ref)
            print("   first mismatches:", bad[:6].tolist(), o[bad[0][0]], ref[bad[0][0]])
        ctx.close()
    del clip_t
