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
p10 = spec.endswith("p10")                     # e.g. 1920x1080x900p10: YUV420P10 (16-bit containers, 10 significant bits)
W, H, frames = (int(x) for x in spec.replace("p10", "").split("x"))
torch.cuda.set_device(0)
clip_t = torch.empty((frames, W * H * 3 // 2), dtype=torch.uint8, device="cuda")
for n0 in range(0, frames, 20):
    n = min(20, frames - n0)
    synth.make_frames(n0, n, W, H, device="cuda", out=clip_t[n0:n0 + n])
if p10:
    c16 = torch.empty((frames, W * H * 3 // 2), dtype=torch.int16, device="cuda")
    for n0 in range(0, frames, 20):
        v = clip_t[n0:n0 + 20].to(torch.int32)
        c16[n0:n0 + 20] = (v * 4 + (v & 3)).to(torch.int16)
    del clip_t
    clip_t = c16
    clip = ab.yv12_clip(clip_t, W, H, frames, True, bits=10)
else:
    clip = ab.yv12_clip(clip_t, W, H, frames, True)
BYTES_PER_FRAME = W * H * 1.5 * (2 if p10 else 1)
prm = ab.default_comb_params()
if p10:
    prm.th_move_y, prm.th_shima_y, prm.th_lshima_y = 80, 48, 144
    prm.th_move_c, prm.th_shima_c, prm.th_lshima_c = 96, 64, 192
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
        gbs = frames * BYTES_PER_FRAME / (ms / n * 1e-3) / 1e9
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
        print("%-40s min %.4f  med %.4f  -> %.0f GB/s (min)" % (cfg, v[0], v[len(v) // 2], frames * BYTES_PER_FRAME / (v[0] * 1e-3) / 1e9))
