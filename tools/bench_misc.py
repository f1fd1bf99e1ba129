"""Secondary configs of BASELINE.json (parity-test cases, not the headline bench line): device-resident timings with
CUDA events.  configs[2]: 1440x1080 combing pass; configs[3]: LogoScan accumulation; AMTAnalyzeLogo (33 evals/frame);
AMTEraseLogo in place."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import amatsukaze_b200 as ab
from amatsukaze_b200 import synth
if os.environ.get("AMTK_LIB"):
    ab.capi.LIB_PATH = os.environ["AMTK_LIB"]      # codegen experiments: load another build of the library


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def clip_of(w, h, n, mode, logo=None, imgx=0, imgy=0):
    t = torch.empty((n, w * h * 3 // 2), dtype=torch.uint8, device="cuda")
    for n0 in range(0, n, 20):
        k = min(20, n - n0)
        synth.make_frames(n0, k, w, h, device="cuda", mode=mode, logo=logo, imgx=imgx, imgy=imgy, out=t[n0:n0 + k])
    return t


torch.cuda.set_device(0)
ctx = ab.Context(0, torch.cuda.current_stream().cuda_stream)
lg = synth.make_logo(64, 64)
# configs[2]: 1440x1080i, 1800 frames (3600 fields), combing pass
w, h, n = 1440, 1080, 1800
t = clip_of(w, h, n, "telecine")
clip = ab.yv12_clip(t, w, h, n, True)
ms = timed(lambda: ctx.comb_frames(clip))
print("configs[2] comb 1440x1080 x%d: %.3f ms  %.0f frames/s  %.0f GB/s algorithmic" % (n, ms, n / ms * 1e3, n * w * h * 1.5 / ms / 1e6))
del t, clip
# configs[1] pieces on 1920x1080
w, h, n = 1920, 1080, 1800
t = clip_of(w, h, n, "flat", logo=lg, imgx=1700, imgy=60)
clip = ab.yv12_clip(t, w, h, n, True)
acc = ctx.logo_scan(64, 64, 12)
ms = timed(lambda: acc.add_frames(clip, 1700, 60))
print("configs[3] LogoScan accumulate 64x64 ROI: %.3f ms per %d frames -> %.0f frames/s (10000 frames = %.1f ms); valid so far %d"
      % (ms, n, n / ms * 1e3, ms * 10000 / n, acc.num_valid))
raw = ab.Logo.create(lg["data"], 64, 64, w, h, 1700, 60)
de, top, bot = raw.deint().create_mask(0.35), raw.field(0).create_mask(0.35), raw.field(1).create_mask(0.35)
ms = timed(lambda: ctx.scan_frames(clip, [de]))
print("LogoFrame::ScanFrame (2 evals/frame): %.3f ms per %d frames -> %.0f frames/s" % (ms, n, n / ms * 1e3))
ms = timed(lambda: ctx.analyze_frames(clip, de, top, bot), reps=3)
print("AMTAnalyzeLogo (33 evals/frame): %.3f ms per %d frames -> %.0f frames/s" % (ms, n, n / ms * 1e3))
fades = np.tile(np.array([[0.7, 0.7]], np.float32), (n, 1))
ms = timed(lambda: ctx.erase_logo(clip, raw, fades))
print("AMTEraseLogo in place: %.3f ms per %d frames -> %.0f frames/s" % (ms, n, n / ms * 1e3))
# YUV420P10 through the streaming 16-bit kernel
t16 = (t[:600].to(torch.int32) * 4 + 1).to(torch.int16).contiguous()
clip16 = ab.yv12_clip(t16, w, h, 600, True, bits=10)
ms = timed(lambda: ctx.comb_frames(clip16), reps=3)
print("comb YUV420P10 1920x1080 x600 (streaming u16 kernel): %.3f ms -> %.0f frames/s  %.0f GB/s algorithmic" % (ms, 600 / ms * 1e3, 600 * w * h * 3.0 / ms / 1e6))
del t16, clip16
os.environ["AMTK_COMB_GENERIC"] = "1"
ctx2 = ab.Context(0, torch.cuda.current_stream().cuda_stream)
ms = timed(lambda: ctx2.comb_frames(clip), reps=2)
print("comb generic kernel (fallback / 16-bit path) on the same 8-bit clip: %.3f ms -> %.0f frames/s" % (ms, n / ms * 1e3))
