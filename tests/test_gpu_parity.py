"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU oracle on the same bytes.
Bit-exact for the integer combing counters AND for the float logo scores (the kernels replicate the reference's
AVX summation tree and its sequential score sum); the 1e-5 relative tolerance of BASELINE.json is therefore not
needed and not used -- tests assert exact equality of the float bit patterns."""
import numpy as np
import pytest
import torch

import amatsukaze_b200 as ab
from amatsukaze_b200 import synth

pytestmark = pytest.mark.gpu

W, H = 256, 128
IMGX, IMGY = 160, 32


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _logos(po, w=64, h=64, imgw=W, imgh=H, imgx=IMGX, imgy=IMGY, ratio=0.35):
    lg = synth.make_logo(w, h)
    p = ab.Logo.create(lg["data"], w, h, imgw, imgh, imgx, imgy)
    o = po.OracleLogo.create(lg["data"], w, h, imgw, imgh, imgx, imgy)
    P = {"raw": p, "deint": p.deint().create_mask(ratio), "top": p.field(0).create_mask(ratio), "bot": p.field(1).create_mask(ratio)}
    O = {"raw": o, "deint": o.deint().create_mask(ratio), "top": o.field(0).create_mask(ratio), "bot": o.field(1).create_mask(ratio)}
    return lg, P, O


def _clip(frames, w, h, on_device=True, bits=8):
    return ab.yv12_clip(frames, w, h, frames.shape[0], on_device, bits)


def test_scan_frames_bit_exact(ctx, oracle):
    po = oracle
    lg, P, O = _logos(po)
    n = 48
    fr = synth.make_frames(40, n, W, H, device="cuda", logo=lg, imgx=IMGX, imgy=IMGY, logo_period=40)
    out = ctx.scan_frames(_clip(fr, W, H), [P["deint"]]).cpu().numpy()
    Y, _, _ = synth.split_planes(fr, W, H)
    ref = np.stack([O["deint"].scan_frame(Y[i]) for i in range(n)])
    assert np.array_equal(_bits(out[:, 0, :]), _bits(ref))
    assert ref[:, 0].max() > 0.5 and ref[:, 0].min() < 0.1        # logo on and off both occur


def test_scan_frames_size_mismatch_and_multi_logo(ctx, oracle):
    po = oracle
    lg, P, O = _logos(po)
    lg2 = synth.make_logo(64, 64, seed=2)
    other = ab.Logo.create(lg2["data"], 64, 64, 1920, 1080, 100, 100).deint().create_mask(0.35)   # wrong frame size
    n = 5
    fr = synth.make_frames(60, n, W, H, device="cuda", logo=lg, imgx=IMGX, imgy=IMGY)
    out = ctx.scan_frames(_clip(fr, W, H), [other, P["deint"]]).cpu().numpy()
    assert np.all(out[:, 0, 0] == 0.0) and np.all(out[:, 0, 1] == -1.0)       # LogoScan.hpp:1551-1558
    Y, _, _ = synth.split_planes(fr, W, H)
    ref = np.stack([O["deint"].scan_frame(Y[i]) for i in range(n)])
    assert np.array_equal(_bits(out[:, 1, :]), _bits(ref))


def test_analyze_frames_bit_exact(ctx, oracle):
    po = oracle
    lg, P, O = _logos(po)
    n = 10
    fr = synth.make_frames(52, n, W, H, device="cuda", logo=lg, imgx=IMGX, imgy=IMGY, logo_period=20)
    out = ctx.analyze_frames(_clip(fr, W, H), P["deint"], P["top"], P["bot"]).cpu().numpy()
    Y, _, _ = synth.split_planes(fr, W, H)
    ref = np.stack([po.or_analyze_frame(O["deint"], O["top"], O["bot"], Y[i]) for i in range(n)])
    assert np.array_equal(_bits(out), _bits(ref))


def test_comb_bit_exact_small(ctx, oracle):
    po = oracle
    n = 9
    for (w, h, mode) in ((256, 128, "interlaced"), (224, 136, "telecine"), (352, 270, "interlaced"), (96, 36, "interlaced")):
        fr = synth.make_frames(0, n, w, h, device="cuda", mode=mode)
        prm = ab.default_comb_params()
        out = ctx.comb_frames(_clip(fr, w, h), prm).cpu().numpy()
        Y, U, V = synth.split_planes(fr, w, h)
        ref = po.or_comb_clip(Y, U, V, prm.as_list())
        assert np.array_equal(out, ref), (w, h, mode, out[:3], ref[:3])
        assert ref[:, 1].sum() > 0 and ref[:, 0].sum() > 0


def test_fused_and_host_path(ctx, oracle):
    po = oracle
    lg, P, O = _logos(po)
    n = 20
    fr = synth.make_frames(45, n, W, H, device="cuda", logo=lg, imgx=IMGX, imgy=IMGY, logo_period=30)
    prm = ab.default_comb_params()
    s, c = ctx.scan_comb_frames(_clip(fr, W, H), [P["deint"]], prm)
    Y, U, V = synth.split_planes(fr, W, H)
    rs = np.stack([O["deint"].scan_frame(Y[i]) for i in range(n)])
    rc = po.or_comb_clip(Y, U, V, prm.as_list())
    assert np.array_equal(_bits(s.cpu().numpy()[:, 0]), _bits(rs)) and np.array_equal(c.cpu().numpy(), rc)
    # same through HOST buffers (library stages frames through HBM itself)
    host = fr.cpu().numpy()
    s2, c2 = ctx.scan_comb_frames(_clip(host, W, H, on_device=False), [P["deint"]], prm)
    assert np.array_equal(_bits(s2[:, 0]), _bits(rs)) and np.array_equal(c2, rc)
