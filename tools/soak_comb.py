"""Randomised parity soak of the streaming pass against the spec oracle: random geometry (also widths that are not multiples
of 16 -> generic kernel), frame counts, thresholds, 8-bit and YUV420P10, whole calls and range calls.
    python tools/soak_comb.py [seconds] [seed]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import amatsukaze_b200 as ab
from amatsukaze_b200 import synth
from oracle import pyoracle as po

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
torch.cuda.set_device(0)
ctx = ab.Context(0, torch.cuda.current_stream().cuda_stream)
t0, cases, px = time.time(), 0, 0
while time.time() - t0 < budget:
    w = int(rng.choice([2 * int(rng.integers(8, 400)), 16 * int(rng.integers(1, 130)), 128 * int(rng.integers(1, 16))]))
    h = 2 * int(rng.integers(3, 330))
    n = int(rng.integers(1, 7))
    p10 = bool(rng.integers(0, 2))
    mode = str(rng.choice(["interlaced", "telecine", "flat"]))
    fr = synth.make_frames(int(rng.integers(0, 100)), n, w, h, device="cuda", mode=mode, seed=int(rng.integers(1, 1 << 30)))
    prm = ab.default_comb_params()
    if p10:
        v = fr.to(torch.int32)
        fr = (v * 4 + (v & 3)).to(torch.int16).contiguous()
        if rng.integers(0, 2):
            fr.view(-1)[:: int(rng.integers(3, 50))] = int(rng.choice([0, 1023]))
        prm.th_move_y, prm.th_move_c = int(rng.integers(1, 1200)), int(rng.integers(1, 32769) if rng.integers(0, 4) == 0 else rng.integers(1, 300))
        prm.th_shima_y, prm.th_lshima_y = int(rng.integers(1, 400)), int(rng.integers(1, 7000))
        prm.th_shima_c, prm.th_lshima_c = int(rng.integers(1, 9000)), int(rng.integers(1, 100000) if rng.integers(0, 4) == 0 else rng.integers(1, 600))
        clip = ab.yv12_clip(fr, w, h, n, True, bits=10)
        a = fr.cpu().numpy().view(np.uint16)
    else:
        if rng.integers(0, 2):
            fr.view(-1)[:: int(rng.integers(3, 50))] = int(rng.choice([0, 255]))
        prm.th_move_y, prm.th_move_c = int(rng.integers(1, 129)), int(rng.integers(1, 129))
        prm.th_shima_y, prm.th_lshima_y = int(rng.integers(1, 200)), int(rng.integers(1, 2048))
        prm.th_shima_c, prm.th_lshima_c = int(rng.integers(1, 2048)), int(rng.integers(1, 300))
        clip = ab.yv12_clip(fr, w, h, n, True)
        a = fr.cpu().numpy()
    ysz, csz = w * h, (w // 2) * (h // 2)
    Y, U, V = a[:, :ysz].reshape(n, h, w), a[:, ysz:ysz + csz].reshape(n, h // 2, w // 2), a[:, ysz + csz:].reshape(n, h // 2, w // 2)
    ref = po.or_comb_clip(Y, U, V, prm.as_list())
    got = ctx.comb_frames(clip, prm).cpu().numpy()
    assert np.array_equal(got, ref), ("MISMATCH", w, h, n, p10, mode, prm.as_list(), np.argwhere(got != ref)[:4])
    if n > 2:
        k = int(rng.integers(1, n))
        part = np.concatenate([ctx.comb_frames(clip, prm, 0, k).cpu().numpy(), ctx.comb_frames(clip, prm, k, n - k).cpu().numpy()])
        assert np.array_equal(part, ref), ("RANGE MISMATCH", w, h, n, k, p10)
    cases += 1
    px += w * h * n
print("soak ok: %d random cases (%.1f Mpx) in %.0f s, all counters equal to the spec oracle" % (cases, px / 1e6, time.time() - t0))
ctx.close()
