"""Round-2 host pieces on a real device, driven through tests/cpp/test_pipeline.cpp: the multi-pass telecine driver on one
HBM-resident clip, AMTSource's ingest semantics, and the device-frame path of the erase chain."""
import os
import struct
import subprocess

import numpy as np
import pytest

import amatsukaze_b200 as ab
from amatsukaze_b200 import synth, _build
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
W, H, IMGX, IMGY = 256, 128, 160, 32


@pytest.fixture(scope="module")
def exe():
    return _build.build_pipeline_test() if os.path.exists("/usr/bin/g++") else _build.PIPELINE_TEST


def _write_raw1(path, frames, w=W, h=H):
    with open(path, "wb") as f:
        f.write(b"AMTSRAW1" + struct.pack("<6i", w, h, 8, frames.shape[0], 30000, 1001))
        f.write(frames.tobytes())


def test_three_pass_telecine_on_one_resident_clip(exe, tmp_path):
    """KFMVfrScript = Misc.cs:1305-1323: passes 0 and 1 are pre-processes (counters; pulldown decision), pass 2 is the
    output; AMTFilterSource appends AMTDecimate and reads the timecodes.  The clip is opened (uploaded) ONCE."""
    n = 63
    frames = synth.make_frames(0, n, W, H, mode="telecine").numpy()
    _write_raw1(tmp_path / "amts0.dat", frames)
    r = subprocess.run([exe, "passes", str(tmp_path), "-", "vfr"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    line = [l for l in r.stdout.splitlines() if l.startswith("passes:")][0]
    kv = dict(p.split("=") for p in line.split()[1:])
    assert int(kv["preproc"]) == 2 and int(kv["uploads"]) == 1
    cycles = 12                                              # 63 frames = 12 full 5-frame cycles + 3
    assert int(kv["out_frames"]) == n - cycles and int(kv["timecodes"]) == n - cycles + 1 and int(kv["vfrfps"]) == 60
    assert "pass 0: preproc=1 frames=63" in r.stdout and "pass 1: preproc=1 frames=63" in r.stdout
    assert "decimate=1 map: 0 1 2 4 5 6 7 9 10 11 12 14" in r.stdout
    # pass 1's counters are the metric itself
    got = np.loadtxt(tmp_path / "v0-0-0.avstmp.combstat.txt", dtype=np.int64).astype(np.int32)
    Y, U, V = synth.split_planes(frames, W, H)
    assert np.array_equal(got, po.or_comb_clip(Y, U, V, ab.default_comb_params().as_list()))
    dur = [int(x) for x in open(tmp_path / "v0-0-0.avstmp.duration.txt").read().split()]
    assert sum(dur) == n and dur[:5] == [1, 1, 2, 1, 1]
    # the output frames are the source frames the decimation map selects (no eraser configured)
    out = np.fromfile(tmp_path / "out_frames.bin", np.uint8).reshape(-1, W * H * 3 // 2)
    src_of = np.concatenate([[0], np.cumsum(dur)[:-1]])
    for k, i in enumerate(range(0, n - cycles, 5)):
        assert np.array_equal(out[k], frames[src_of[i]]), i
    # constant-frame-rate script: one pre-process pass only (Misc.cs:1311-1312); no duration file -> no AMTDecimate
    for f in tmp_path.glob("v0-0-0.avstmp*"):
        f.unlink()
    r = subprocess.run([exe, "passes", str(tmp_path), "-", "cfr"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "passes: preproc=1 uploads=1 out_frames=63" in r.stdout and "decimate=0" in r.stdout


def test_filter_source_with_logo_eraser_in_make_source(exe, tmp_path):
    """MakeSource() = AMTSource + AMTEraseLogo(AMTAnalyzeLogo(logo), logo, logof, maxfade) (FilteredSource.hpp:441-475): the
    erased frames reach the output of the pass driver, computed on device frames."""
    n = 40
    lg = synth.make_logo(64, 64, seed=1)
    frames = synth.make_frames(35, n, W, H, logo=lg, imgx=IMGX, imgy=IMGY, logo_period=20).numpy()
    _write_raw1(tmp_path / "amts0.dat", frames)
    logo_path = str(tmp_path / "logo.lgd")
    ab.Logo.create(lg["data"], 64, 64, W, H, IMGX, IMGY).save(logo_path)
    # logoframe file (LogoScan.hpp:1818-1819 format): logo fading in around frame 6, out around frame 30
    open(tmp_path / "logof0.txt", "w").write("%6d S 0 ALL %6d %6d\n%6d E 0 ALL %6d %6d\n" % (6, 5, 7, 30, 28, 32))
    r = subprocess.run([exe, "passes", str(tmp_path), logo_path, "cfr"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "passes: preproc=1 uploads=1 out_frames=40" in r.stdout
    out = np.fromfile(tmp_path / "out_frames.bin", np.uint8).reshape(-1, W * H * 3 // 2)
    raw = po.OracleLogo.create(lg["data"], 64, 64, W, H, IMGX, IMGY)
    de, top, bot = raw.deint().create_mask(0.35), raw.field(0).create_mask(0.35), raw.field(1).create_mask(0.35)
    Y, _, _ = synth.split_planes(frames, W, H)
    rec = np.stack([po.or_analyze_frame(de, top, bot, Y[i]) for i in range(n)])
    fr = np.zeros(n, int)                                    # ReadLogoFrameFile (:1421-1461)
    fr[5:8] = 1; fr[7:29] = 2; fr[29:33] = 1
    changed = 0
    for k, i in enumerate(range(0, n, 5)):
        win = [fr[max(0, min(n - 1, i + d))] for d in range(-8, 9)]          # CalcFade (:1317-1341), maxfade 16
        if all(v == win[0] for v in win):
            ft = fb = 1.0 if fr[i] == 2 else 0.0
        else:
            ft, fb = po.or_calc_fade2(rec, n, i)
        Yi, Ui, Vi = [np.ascontiguousarray(p[i]) for p in synth.split_planes(frames.copy(), W, H)]
        po.or_erase_frame(raw, Yi, Ui, Vi, ft, fb)
        exp = np.concatenate([Yi.ravel(), Ui.ravel(), Vi.ravel()])
        assert np.array_equal(out[k], exp), i
        changed += int(not np.array_equal(exp, frames[i]))
    assert changed > 0


@pytest.mark.parametrize("nv12", [0, 1])
def test_amtsource_ingest_semantics(exe, tmp_path, nv12):
    """Picture structures -> output frames (StreamReform.hpp:874-904), half-delay weave = MergeField(prev, cur)
    (AMTSource.hpp:291-366,524-551), NV12 split (Copy2), FrameType property from the picture that supplies the top field."""
    w, h = 96, 64
    pics = [0, 3, 4, 3, 6, 6, 1, 4, 2, 5, 4]                 # FRAME TFF BFF TFF BFF_RFF BFF_RFF DOUBLING BFF TRIPLING TFF_RFF BFF
    ptype = [1, 2, 3, 3, 2, 1, 2, 3, 3, 2, 1]
    nd = len(pics)
    dec = synth.make_frames(5, nd, w, h).numpy()             # planar decoded pictures
    ysz, csz = w * h, (w // 2) * (h // 2)
    store = dec.copy()
    if nv12:                                                 # decoder output with interleaved chroma
        u = dec[:, ysz:ysz + csz].reshape(nd, h // 2, w // 2)
        v = dec[:, ysz + csz:].reshape(nd, h // 2, w // 2)
        uv = np.stack([u, v], axis=-1).reshape(nd, -1)
        store = np.concatenate([dec[:, :ysz], uv], axis=1)
    path = tmp_path / "clip.amtsraw2"
    with open(path, "wb") as f:
        f.write(b"AMTSRAW2" + struct.pack("<6i", w, h, 8, nd, 30000, 1001) + struct.pack("<i", nv12))
        f.write(bytes(np.stack([pics, ptype], axis=1).astype(np.uint8).ravel()))
        f.write(store.tobytes())
    r = subprocess.run([exe, "ingest", str(path), str(tmp_path / "out.bin")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    # expected plan, from the reference's rules
    plan = []                                                # (top, bottom)
    for d, p in enumerate(pics):
        if p in (0, 3, 5): plan.append((d, d))
        elif p == 1: plan += [(d, d)] * 2
        elif p == 2: plan += [(d, d)] * 3
        elif p == 4: plan.append((d - 1, d))
        elif p == 6: plan += [(d - 1, d), (d, d)]
    out = np.fromfile(tmp_path / "out.bin", np.uint8).reshape(-1, w * h * 3 // 2)
    assert out.shape[0] == len(plan) == 16
    for k, (t, b) in enumerate(plan):
        exp = dec[b].copy()
        for (off, ph, pw) in ((0, h, w), (ysz, h // 2, w // 2), (ysz + csz, h // 2, w // 2)):
            e = exp[off:off + ph * pw].reshape(ph, pw)
            e[0::2] = dec[t][off:off + ph * pw].reshape(ph, pw)[0::2]          # even rows from the top picture (Copy1)
        assert np.array_equal(out[k], exp), (k, t, b)
    types = [int(x) for x in r.stdout.split("types:")[1].splitlines()[0].split()]
    assert types == [ptype[t] for t, _ in plan]
    assert "mt=1 parity=1 devtypes=3" in r.stdout            # MT_NICE_FILTER; interlaced; DEV_TYPE_CPU | DEV_TYPE_CUDA


def test_device_frames_equal_cpu_frames(exe, tmp_path):
    """AMTEraseLogo(AMTAnalyzeLogo(src), ...) pulled through IClip::GetFrame: with a CUDA consumer every frame stays in HBM
    (zero-copy source view, device-to-device MakeWritable, in-place Delogo); results equal the CPU-frame path."""
    n = 33
    lg = synth.make_logo(64, 64, seed=1)
    frames = synth.make_frames(35, n, W, H, logo=lg, imgx=IMGX, imgy=IMGY, logo_period=16).numpy()
    _write_raw1(tmp_path / "clip.amtsraw", frames)
    logo_path = str(tmp_path / "logo.lgd")
    ab.Logo.create(lg["data"], 64, 64, W, H, IMGX, IMGY).save(logo_path)
    r = subprocess.run([exe, "devframes", str(tmp_path / "clip.amtsraw"), logo_path, str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "devframes mode=0: device_frames=0 of 33" in r.stdout and "devframes mode=1: device_frames=33 of 33" in r.stdout
    assert "identical=1" in r.stdout
    got = np.fromfile(tmp_path / "erased_chain.bin", np.uint8).reshape(n, -1)
    assert not np.array_equal(got, frames)
