"""CPU tests of the checker itself: the plain-C oracle (oracle/amtk_oracle.c) must reproduce
  (1) the committed golden vectors the REFERENCE'S OWN code produced (tests/golden/logo_golden.json), always;
  (2) the reference's own compiled code (oracle/_ref) live, when that library is present.
All float comparisons are on bit patterns."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

from amatsukaze_b200 import synth
from oracle import pyoracle as po

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "logo_golden.json")))
W, H, IMGX, IMGY = 256, 128, 160, 32
needs_ref = pytest.mark.skipif(not po.ref_available(), reason="oracle/_ref not built (needs /root/reference)")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32).ravel().tolist()


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def data():
    lg = synth.make_logo(64, 64, seed=1)
    frames = synth.make_frames(40, 24, W, H, seed=0x5EED0001, logo=lg, imgx=IMGX, imgy=IMGY, logo_period=20).numpy()
    raw = po.OracleLogo.create(lg["data"], 64, 64, W, H, IMGX, IMGY)
    logos = {"raw": raw, "deint": raw.deint().create_mask(0.35), "top": raw.field(0).create_mask(0.35),
             "bot": raw.field(1).create_mask(0.35), "deint10": raw.deint().create_mask(0.1)}
    return lg, frames, logos


def test_corr5x5_avx_tree_matches_golden():
    g = GOLD["corr5x5"]
    rng = np.random.default_rng(g["seed"])
    Y = np.concatenate([(rng.random(20 * 20) * 255).astype(np.float32), np.zeros(8, np.float32)])
    K = np.concatenate([rng.standard_normal(25).astype(np.float32), np.zeros(8, np.float32)])
    L = po.oracle_lib()
    sums, avgs, scalar = [], [], []
    for y in range(2, 18):
        for x in range(2, 18):
            a = C.c_float()
            sums.append(L.amtk_or_corr5x5(K.ctypes.data_as(po.c_float_p), Y.ctypes.data_as(po.c_float_p), x, y, 20, C.byref(a)))
            avgs.append(a.value)
            scalar.append(L.amtk_or_corr5x5_scalar_order(K.ctypes.data_as(po.c_float_p), Y.ctypes.data_as(po.c_float_p), x, y, 20, None))
    assert bits(sums) == g["sum_bits"] and bits(avgs) == g["avg_bits"]
    # the scalar summation order is NOT what the reference runs on AVX hosts and differs in the last bits
    assert bits(scalar) != g["sum_bits"]
    assert np.allclose(scalar, sums, rtol=2e-3, atol=1e-2)


def test_tables_match_golden(data):
    _, _, logos = data
    for name, t in GOLD["tables"].items():
        l = logos[name]
        ny = l.s.w * l.s.h
        assert l.s.maskpixels == t["maskpixels"] and l.s.count == t["count"]
        assert bits([l.s.blackScore])[0] == t["black_bits"]
        assert digest(l.data()[:2 * ny]) == t["ab_sha"]
        assert digest(l.mask()) == t["mask_sha"]
        assert digest(l.kernels()) == t["kernels_sha"]
        assert digest(l.scales()) == t["scales_sha"]


def test_scan_and_analyze_match_golden(data):
    _, frames, logos = data
    Y, _, _ = synth.split_planes(frames, W, H)
    for i in range(frames.shape[0]):
        assert bits(logos["deint"].scan_frame(Y[i])) == GOLD["scan_frame_bits"][i]
    for k, i in enumerate(GOLD["analyze_frames"]):
        assert bits(po.or_analyze_frame(logos["deint"], logos["top"], logos["bot"], Y[i])) == GOLD["analyze_bits"][k]
    on = np.array([np.array(b, np.uint32).view(np.float32)[0] for b in GOLD["scan_frame_bits"]])
    assert on.max() > 0.8 and on.min() < 0.2          # the fixture covers logo present AND absent


def test_fade_sweep_matches_golden(data):
    _, frames, logos = data
    Y, _, _ = synth.split_planes(frames, W, H)
    de = np.zeros(64 * 64 + 8, np.float32)
    roi = np.ascontiguousarray(Y[12])
    po.oracle_lib().amtk_or_deint_y_u8(de.ctypes.data_as(po.c_float_p), roi.reshape(-1)[IMGX + IMGY * W:].ctypes.data_as(po.c_u8_p), W, 64, 64)
    got = [logos["deint10"].evaluate(de, 255.0, np.float32(0.1) * np.float32(fi)) for fi in range(20)]
    assert bits(got) == GOLD["fade_sweep_bits"]


def test_logoscan_matches_golden():
    flat = synth.make_frames(0, 40, 128, 96, seed=0x5EED0004, mode="flat", logo=synth.make_logo(32, 32, seed=3), imgx=64, imgy=32).numpy()
    fy, fu, fv = synth.split_planes(flat, 128, 96)
    sc = po.OracleScan(32, 32, 12)
    valid = [sc.add_frame(fy[i][32:64, 64:96], fu[i][16:32, 32:48], fv[i][16:32, 32:48]) for i in range(flat.shape[0])]
    g = GOLD["scan"]
    assert valid == g["valid"] and sc.nframes == g["nframes"] and 0 < sc.nframes < len(valid)
    assert digest(sc.sums()) == g["sums_sha"]
    lg = sc.get_logo(255, clean=False)
    assert digest(lg) == g["logo_sha"] and bits(lg[:16]) == g["logo_head_bits"]
    assert digest(sc.get_logo(255, clean=True)) == g["logo_clean_sha"]


def test_logoscan_insufficient_frames_returns_none():
    sc = po.OracleScan(16, 16, 12)
    assert sc.get_logo(255) is None      # 0 frames -> NaN slopes -> the reference returns nullptr (LogoScan.hpp:391,503)


@needs_ref
def test_oracle_equals_reference_live():
    """Random logos / frames beyond the golden set, incl. 16-bit samples, odd sizes and a logo whose mask
    spills into zero-variance pixels (count < maskpixels)."""
    rng = np.random.default_rng(7)
    for case, (w, h, ratio, bitsps) in enumerate(((64, 64, 0.35, 8), (48, 40, 0.9, 8), (64, 32, 0.35, 10), (32, 64, 0.1, 8))):
        lg = synth.make_logo(w, h, seed=case)
        data = lg["data"].copy()
        if case == 1:
            data[: w * h][(rng.random(w * h) < 0.3)] *= 1.0      # keep flat areas: high maskratio forces border picks
        fw, fh, ix, iy = 320, 200, 100, 60
        r = po.RefLogo.create(data, w, h, fw, fh, ix, iy).deint().create_mask(ratio)
        o = po.OracleLogo.create(data, w, h, fw, fh, ix, iy).deint().create_mask(ratio)
        assert r.visited_count() == o.s.count and r.dims()["maskpixels"] == o.s.maskpixels
        assert np.array_equal(r.mask(), o.mask())
        assert np.array_equal(r.kernels().view(np.uint32), o.kernels().view(np.uint32))
        assert np.array_equal(r.scales().view(np.uint32), o.scales().view(np.uint32))
        assert bits([r.black_score()]) == bits([o.s.blackScore])
        maxv = float((1 << bitsps) - 1)
        for _ in range(4):
            if bitsps == 8:
                plane = rng.integers(16, 236, (fh, fw), dtype=np.uint8)
            else:
                plane = rng.integers(64, 940, (fh, fw)).astype(np.uint16)
            assert bits(po.ref_scan_frame(r, plane, maxv)) == bits(o.scan_frame(plane, maxv=maxv))
        if case == 1:
            assert o.s.count < o.s.maskpixels          # the quirk case really happened


@needs_ref
def test_oracle_logoscan_equals_reference_live():
    rng = np.random.default_rng(11)
    ro, oo = po.RefScan(24, 16, 10), po.OracleScan(24, 16, 10)
    for i in range(60):
        base = int(rng.integers(30, 200))
        y = (base + rng.integers(-3, 4, (16, 24))).astype(np.uint8)
        if i % 7 == 0:
            y[0, 3] = 255                               # breaks the flat-border test
        u = (128 + rng.integers(-2, 3, (8, 12))).astype(np.uint8)
        v = (128 + rng.integers(-2, 3, (8, 12))).astype(np.uint8)
        y[4:12, 6:18] = np.clip(y[4:12, 6:18].astype(int) + 40, 0, 255).astype(np.uint8)
        assert ro.add_frame(y, u, v) == oo.add_frame(y, u, v)
    assert ro.nframes == oo.nframes
    assert np.array_equal(ro.sums(), oo.sums())
    for clean in (False, True):
        a, b = ro.get_logo(255, clean), oo.get_logo(255, clean)
        assert (a is None) == (b is None)
        if a is not None:
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.skipif(not po.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_port_delogo_and_calcfade2_equal_the_reference_code_live():
    """Round 2: AMTEraseLogo::Delogo and CalcFade2 (LogoScan.hpp:1248-1315) are compiled from the reference's own lines into
    oracle/_ref; the plain-C port (which the GPU erase kernel and the product's amtk_calc_fade2 are tested against) must
    reproduce them exactly: pixel bytes for Delogo (rounding, clamping, per-field pitches), the selected fades for
    CalcFade2 incl. the double-offset quirk (:1273-1275) and the clip-end clamps."""
    if not po.ref_has_erase():
        pytest.skip("prebuilt oracle/_ref predates the Delogo/CalcFade2 extraction")
    rng = np.random.default_rng(11)
    for dtype, maxv in ((np.uint8, 255.0), (np.uint16, 1023.0)):
        for (w, h, lp, ip) in ((64, 64, 64, 96), (32, 16, 64, 200), (7, 5, 7, 7)):       # field passes: logopitch 2w, imgpitch 2*pitch
            img = rng.integers(0, int(maxv) + 1, size=(h, ip)).astype(dtype)
            A = rng.uniform(0.8, 1.6, size=h * lp).astype(np.float32)
            B = rng.uniform(-0.6, 0.1, size=h * lp).astype(np.float32)
            for fade in (0.0, 0.1, 0.3, 0.5, 0.9, 1.0):
                a, b = img.copy(), img.copy()
                po.or_delogo(a, A, B, fade, maxv, logopitch=lp, imgpitch=ip, w=w, h=h)
                po.ref_delogo(b, A, B, fade, maxv, logopitch=lp, imgpitch=ip, w=w, h=h)
                assert np.array_equal(a, b), (dtype, w, h, fade)
                assert fade == 0.0 or not np.array_equal(a, img)
    for N in (1, 5, 8, 9, 23, 64, 101):
        rec = rng.uniform(0.0, 1.0, size=(N, 33)).astype(np.float32)
        # sudden appear / disappear patterns so that both branches of :1295-1314 are taken
        for k in range(N):
            rec[k, : 11] += np.abs(np.arange(11) - (0 if (k // 7) % 2 == 0 else 10)) * np.float32(0.5)
        took = set()
        for n in range(N):
            want = po.ref_calc_fade2(rec, N, n)
            got = po.or_calc_fade2(rec, N, n)
            assert got == want, (N, n, got, want)
            took.add(want[0] == want[1])
        if N >= 23:
            assert took == {True, False}


@pytest.mark.skipif(not po.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_reference_mergefield_is_the_even_odd_row_weave_live():
    """AMTSource::MergeField / Copy1 / Copy2 compiled from the reference's own lines (AMTSource.hpp:291-355): even rows of
    every plane from `top`, odd rows from `bottom`, NV12 chroma de-interleaved -- the statement the GPU weave kernel
    (amtk_weave_frames) is tested against on the device."""
    if not po.ref_has_mergefield():
        pytest.skip("prebuilt oracle/_ref predates the MergeField extraction")
    rng = np.random.default_rng(5)
    for (w, h) in ((16, 8), (208, 72), (64, 36)):      # heights are multiples of 4: Copy1 writes row pairs of the chroma planes too
        ysz, cw, ch = w * h, w // 2, h // 2
        t = rng.integers(0, 256, ysz + 2 * cw * ch).astype(np.uint8)
        b = rng.integers(0, 256, ysz + 2 * cw * ch).astype(np.uint8)
        got = po.ref_merge_field(t, b, w, h)
        for (o, rows, cols) in ((0, h, w), (ysz, ch, cw), (ysz + cw * ch, ch, cw)):
            exp = t[o:o + rows * cols].reshape(rows, cols).copy()
            exp[1::2] = b[o:o + rows * cols].reshape(rows, cols)[1::2]
            assert np.array_equal(got[o:o + rows * cols].reshape(rows, cols), exp)

        def to_nv12(a):
            uv = np.stack([a[ysz:ysz + cw * ch], a[ysz + cw * ch:]], axis=1).reshape(-1)
            return np.concatenate([a[:ysz], uv])
        assert np.array_equal(po.ref_merge_field(to_nv12(t), to_nv12(b), w, h, nv12=True), got)


@pytest.mark.skipif(not po.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_reference_frame_drivers_equal_their_restated_compositions_live():
    """AMTAnalyzeLogo::GetFrameT (LogoScan.hpp:1119-1161) and LogoFrame::ScanFrame (:1543-1568) compiled from the reference's
    own lines: the compositions the parity tests use (ref_analyze_frame / ref_scan_frame: the reference's DeintY, CopyY and
    EvaluateLogo called in the order those functions call them) give the same bits, including the source-frame clamp at the
    clip end (:1133), the |.| of every evaluation, and the (0, -1) result for invalid or wrong-sized logos (:1551-1558)."""
    if not po.ref_has_drivers():
        pytest.skip("prebuilt oracle/_ref predates the GetFrameT/ScanFrame extraction")
    w, h, imgx, imgy, N = 256, 128, 160, 32, 13
    lg = synth.make_logo(64, 64)
    fr = synth.make_frames(40, N, w, h, device="cpu", logo=lg, imgx=imgx, imgy=imgy, logo_period=12).numpy()
    raw = po.RefLogo.create(lg["data"], 64, 64, w, h, imgx, imgy)
    de, top, bot = raw.deint().create_mask(0.35), raw.field(0).create_mask(0.35), raw.field(1).create_mask(0.35)
    Y = fr[:, : w * h].reshape(N, h, w)
    for n in range((N + 7) // 8):
        got = po.ref_analyze_getframe(de, top, bot, fr, w, h, n)
        want = np.stack([po.ref_analyze_frame(de, top, bot, Y[min(N - 1, 8 * n + i)]) for i in range(8)])
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), n
    other = po.RefLogo.create(lg["data"], 64, 64, w + 16, h, imgx, imgy).deint().create_mask(0.35)      # made for another frame size
    for k in (0, 5, N - 1):
        got = po.ref_scan_frame_code([de, None, other], fr[k], w, h)
        assert np.array_equal(got[0].view(np.uint32), po.ref_scan_frame(de, Y[k]).view(np.uint32))
        assert tuple(got[1]) == (0.0, -1.0) and tuple(got[2]) == (0.0, -1.0)


def test_erase_and_weave_golden_from_the_reference_code(tmp_path):
    """Committed golden vectors produced by the reference's own Delogo / CalcFade2 / CalcFade + ReadLogoFrameFile / MergeField
    (tests/golden/gen_golden.py, round 2): the C port, the product's host-side amtk_calc_fade2 and the even/odd-row weave
    statement reproduce them -- also where /root/reference is absent."""
    import amatsukaze_b200 as ab
    g = GOLD["erase"]
    rng = np.random.default_rng(g["seed"])
    it = iter(g["delogo"])
    for dtype, maxv in (("uint8", 255.0), ("uint16", 1023.0)):
        for (w, h, lp, ip) in ((64, 64, 64, 96), (32, 16, 64, 200)):
            img = rng.integers(0, int(maxv) + 1, size=(h, ip)).astype(dtype)
            A = rng.uniform(0.8, 1.6, size=h * lp).astype(np.float32)
            B = rng.uniform(-0.6, 0.1, size=h * lp).astype(np.float32)
            for fade in (0.1, 0.5, 1.0):
                e = next(it)
                assert (e["dtype"], e["w"], e["h"], e["fade"]) == (dtype, w, h, fade)
                a = img.copy()
                po.or_delogo(a, A, B, fade, maxv, logopitch=lp, imgpitch=ip, w=w, h=h)
                assert digest(a) == e["sha"], e
    c2 = g["calc_fade2"]
    N = c2["n"]
    rng = np.random.default_rng(c2["seed"])
    rec = rng.uniform(0.0, 1.0, size=(N, 33)).astype(np.float32)
    for k in range(N):
        rec[k, :11] += np.abs(np.arange(11) - (0 if (k // 7) % 2 == 0 else 10)) * np.float32(0.5)
    want = np.array(c2["fades_bits"], np.uint32).view(np.float32).reshape(N, 2)
    assert bits(np.array([po.or_calc_fade2(rec, N, n) for n in range(N)], np.float32)) == c2["fades_bits"]
    assert bits(np.array([ab.calc_fade2(rec, N, n) for n in range(N)], np.float32)) == c2["fades_bits"]        # product (host code, no GPU)
    # CalcFade with a logoframe file: uniform +-maxfade/2 neighbourhoods take 0 / 1, the rest CalcFade2 (LogoScan.hpp:1317-1341)
    cf = g["calc_fade"]
    state = np.array(cf["state"], np.int32)
    half = cf["maxfade"] >> 1
    exp = np.zeros((N, 2), np.float32)
    for i in range(N):
        win = state[np.clip(np.arange(i - half, i + half + 1), 0, N - 1)]
        exp[i] = ((1.0, 1.0) if state[i] == 2 else (0.0, 0.0)) if np.all(win == win[0]) else want[i]
    assert bits(exp) == cf["fades_bits"]
    assert set(state.tolist()) == {0, 1, 2}
    # MergeField: even rows from top, odd rows from bottom, every plane
    m = GOLD["mergefield"]
    w, h = m["w"], m["h"]
    rng = np.random.default_rng(m["seed"])
    msz = w * h * 3 // 2
    t = rng.integers(0, 256, msz).astype(np.uint8)
    b = rng.integers(0, 256, msz).astype(np.uint8)
    out = t.copy()
    for (o, rows, cols) in ((0, h, w), (w * h, h // 2, w // 2), (w * h + (w // 2) * (h // 2), h // 2, w // 2)):
        out[o:o + rows * cols].reshape(rows, cols)[1::2] = b[o:o + rows * cols].reshape(rows, cols)[1::2]
    assert digest(out) == m["sha"]
