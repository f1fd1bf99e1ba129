"""Small driver for ncu captures: a few fused steps over a shorter 1080p clip (same kernels, same tile grid)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import amatsukaze_b200 as ab
from amatsukaze_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=300)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--comb-only", action="store_true")
ap.add_argument("--p10", action="store_true", help="YUV420P10 clip (16-bit containers): the integer-lane form of the streaming kernel; implies --comb-only")
a = ap.parse_args()
W, H = 1920, 1080
torch.cuda.set_device(0)
lg = synth.make_logo()
clip_t = torch.empty((a.frames, W * H * 3 // 2), dtype=torch.uint8, device="cuda")
for n0 in range(0, a.frames, 20):
    n = min(20, a.frames - n0)
    synth.make_frames(n0, n, W, H, device="cuda", logo=lg, imgx=1700, imgy=60, out=clip_t[n0:n0 + n])
ctx = ab.Context(0, torch.cuda.current_stream().cuda_stream)
logo = ab.Logo.create(lg["data"], 64, 64, W, H, 1700, 60).deint().create_mask(0.35)
clip = ab.yv12_clip(clip_t, W, H, a.frames, True)
prm = ab.default_comb_params()
if a.p10:
    c16 = torch.empty((a.frames, W * H * 3 // 2), dtype=torch.int16, device="cuda")
    for n0 in range(0, a.frames, 20):
        v = clip_t[n0:n0 + 20].to(torch.int32)
        c16[n0:n0 + 20] = (v * 4 + (v & 3)).to(torch.int16)
    del clip_t
    clip = ab.yv12_clip(c16, W, H, a.frames, True, bits=10)
    prm.th_move_y, prm.th_shima_y, prm.th_lshima_y = 80, 48, 144
    prm.th_move_c, prm.th_shima_c, prm.th_lshima_c = 96, 64, 192
for _ in range(a.steps):
    if a.comb_only or a.p10:
        ctx.comb_frames(clip, prm)
    else:
        ctx.scan_comb_frames(clip, [logo], prm)
torch.cuda.synchronize()
print("done")
