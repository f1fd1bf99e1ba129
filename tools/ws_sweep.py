"""Sweep of warp-stream comb kernel variants on one resident clip; every variant is checked against the first one.
    python tools/ws_sweep.py 1920x1080x1800 "WS=1" "WS=1,WS_WARPS=7" "WS=1,WS_PF=1" ...
Each KEY=VAL sets the environment variable AMTK_COMB_<KEY> for the context that runs that variant."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import amatsukaze_b200 as ab
from amatsukaze_b200 import synth

if os.environ.get("AMTK_LIB"):
    ab.capi.LIB_PATH = os.environ["AMTK_LIB"]
spec, cfgs = sys.argv[1], sys.argv[2:]
W, H, frames = (int(x) for x in spec.split("x"))
torch.cuda.set_device(0)
clip_t = torch.empty((frames, W * H * 3 // 2), dtype=torch.uint8, device="cuda")
for n0 in range(0, frames, 20):
    n = min(20, frames - n0)
    synth.make_frames(n0, n, W, H, device="cuda", out=clip_t[n0:n0 + n])
clip = ab.yv12_clip(clip_t, W, H, frames, True)
prm = ab.default_comb_params()
ref = None
reps = int(os.environ.get("REPS", "10"))
rounds = int(os.environ.get("ROUNDS", "3"))
try:
    import pynvml
    pynvml.nvmlInit()
    _h = pynvml.nvmlDeviceGetHandleByIndex(0)
    def clk():
        return "%d MHz %.0f W" % (pynvml.nvmlDeviceGetClockInfo(_h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(_h) / 1000.0)
except Exception:
    def clk():
        return "?"
best = {}
for cfg in cfgs * rounds:
    for k in [k for k in os.environ if k.startswith("AMTK_COMB_")]:
        del os.environ[k]
    for kv in cfg.split(","):
        if kv:
            k, v = kv.split("=")
            os.environ["AMTK_COMB_" + k] = v
    try:
        ctx = ab.Context(0, torch.cuda.current_stream().cuda_stream)
        out = ctx.comb_frames(clip, prm)
        torch.cuda.synchronize()
        ctx.set_kernel_timing(True)
        for _ in range(reps):
            out = ctx.comb_frames(clip, prm)
        ms, n = ctx.kernel_timing()
        o = out.cpu().numpy()
        if ref is None:
            ref = o
        gbs = frames * W * H * 1.5 / (ms / n * 1e-3) / 1e9
        c = clk()
        best.setdefault(cfg, []).append(ms / n)
        print("%-40s %.4f ms  %.0f GB/s  %.3f of 6486  same=%s  [%s]" % (cfg, ms / n, gbs, gbs / 6486.1, np.array_equal(o, ref), c), flush=True)
        ctx.close()
    except Exception as e:  # a variant that does not exist / does not fit
        print("%-40s FAILED: %s" % (cfg, e), flush=True)
print("---- per variant: min / median ms over %d rounds" % rounds)
for cfg in cfgs:
    v = sorted(best.get(cfg, []))
    if v:
        print("%-40s min %.4f  med %.4f  -> %.0f GB/s (min)" % (cfg, v[0], v[len(v) // 2], frames * W * H * 1.5 / (v[0] * 1e-3) / 1e9))
