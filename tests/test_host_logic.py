"""CPU tests of the PRODUCT's host-side code (no GPU, no compute kernels): the C-ABI library loads and exports every
declared symbol, host logo tables equal the reference's (golden + live), .lgd I/O, CalcFade2, argument validation.
The oracle is only the checker here."""
import ctypes as C
import hashlib
import json
import os
import re

import numpy as np
import pytest

import amatsukaze_b200 as ab
from amatsukaze_b200 import synth
from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "logo_golden.json")))
W, H, IMGX, IMGY = 256, 128, 160, 32


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32).ravel().tolist()


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_library_exports_every_declared_symbol(native_lib):
    hdr = open(os.path.join(ROOT, "include", "amtk_b200.h")).read()
    declared = set(re.findall(r"AMTK_API\s+[\w\s\*]+?\b(amtk_\w+)\s*\(", hdr))
    assert len(declared) >= 30
    bound = {name for name, _, _ in ab.SIGNATURES}
    assert declared == bound, (declared ^ bound)
    for name in declared:
        assert hasattr(native_lib, name)


def test_no_cpu_fallback_without_gpu(native_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert native_lib.amtk_device_count() == 0
    with pytest.raises(ab.AmtkError, match="no CPU fallback"):
        ab.Context(0)


def test_product_does_not_touch_the_oracle():
    """The product sources must not include, link or import anything under oracle/."""
    pkg = os.path.join(ROOT, "amatsukaze_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(root, f)).read()
                assert "pyoracle" not in txt and "amtk_oracle" not in txt and "libamtk_ref" not in txt, f
    out = os.popen("ldd %s" % ab.LIB_PATH).read()
    assert "oracle" not in out and "amtk_ref" not in out


def test_host_tables_match_reference_golden(native_lib):
    lg = synth.make_logo(64, 64, seed=1)
    raw = ab.Logo.create(lg["data"], 64, 64, W, H, IMGX, IMGY)
    logos = {"deint": raw.deint().create_mask(0.35), "top": raw.field(0).create_mask(0.35),
             "bot": raw.field(1).create_mask(0.35), "deint10": raw.deint().create_mask(0.1)}
    for name, t in GOLD["tables"].items():
        l = logos[name]
        i, tb = l.info(), l.tables()
        ny = i.w * i.h
        assert i.maskpixels == t["maskpixels"] and i.count == t["count"]
        assert bits([i.black_score])[0] == t["black_bits"]
        assert digest(tb["data"][:2 * ny]) == t["ab_sha"]
        assert digest(tb["mask"]) == t["mask_sha"]
        assert digest(tb["kernels"]) == t["kernels_sha"]
        assert digest(tb["scales"]) == t["scales_sha"]


def test_host_tables_quirk_count_below_maskpixels(native_lib):
    """maskratio so large that zero-variance / border pixels get selected: kernels exist only for visited pixels."""
    lg = synth.make_logo(48, 40, seed=5)
    p = ab.Logo.create(lg["data"], 48, 40, 320, 200, 100, 60).deint().create_mask(0.9)
    o = po.OracleLogo.create(lg["data"], 48, 40, 320, 200, 100, 60).deint().create_mask(0.9)
    i = p.info()
    assert i.count == o.s.count < i.maskpixels == o.s.maskpixels
    t = p.tables()
    assert np.array_equal(t["mask"], o.mask())
    assert np.array_equal(t["kernels"].view(np.uint32), o.kernels().view(np.uint32))
    assert np.array_equal(t["scales"].view(np.uint32), o.scales().view(np.uint32))
    assert bits([i.black_score]) == bits([o.s.blackScore])


def test_lgd_roundtrip_and_reference_bytes(native_lib, tmp_path):
    lg = synth.make_logo(64, 64, seed=1)
    raw = ab.Logo.create(lg["data"], 64, 64, W, H, IMGX, IMGY)
    p = str(tmp_path / "a.lgd")
    raw.save(p, "No Name", 410)
    blob = bytearray(open(p, "rb").read())
    g = GOLD["lgd"]
    assert len(blob) == g["size"] == 32 + 48 + 64 * 64 * 12 + 540 + (64 * 64 + 2 * 32 * 32) * 2 * 4
    blob[g["pad_offset"]] = 0
    assert hashlib.sha256(bytes(blob)).hexdigest() == g["sha"]        # byte-identical to LogoData::Save
    back = ab.Logo.load(p)
    i = back.info()
    assert (i.w, i.h, i.imgw, i.imgh, i.imgx, i.imgy) == (64, 64, W, H, IMGX, IMGY)
    assert np.array_equal(back.tables()["data"].view(np.uint32), lg["data"].view(np.uint32))
    hdr = back.header
    assert int(np.frombuffer(hdr[:4].tobytes(), np.int32)[0]) == 0x12345
    with pytest.raises(ab.AmtkError, match="Failed to read logo file"):
        ab.Logo.load(str(tmp_path / "missing.lgd"))


def test_calc_fade2_matches_oracle(native_lib):
    rng = np.random.default_rng(3)
    nrec = 50
    rec = rng.random((nrec, 33)).astype(np.float32)
    # make a sudden logo switch around frame 20: p-minimum index jumps from 0 to 10
    for i in range(nrec):
        rec[i, :11] = np.abs(np.arange(11) - (0 if i < 20 else 10)) * 0.1 + rng.random(11).astype(np.float32) * 0.01
    for n in list(range(0, 8)) + list(range(14, 30)) + list(range(44, 50)):
        assert ab.calc_fade2(rec, nrec, n) == po.or_calc_fade2(rec, nrec, n)
    ft, fb = ab.calc_fade2(rec, nrec, 19)
    assert 0.0 <= ft <= 1.0 and 0.0 <= fb <= 1.0


def test_argument_validation(native_lib):
    lg = synth.make_logo(64, 64)
    with pytest.raises(ab.AmtkError):
        ab.Logo.create(lg["data"], 3, 3, W, H, 0, 0)            # too small for a 5x5 window
    raw = ab.Logo.create(lg["data"], 64, 64, W, H, IMGX, IMGY)
    with pytest.raises(ab.AmtkError, match="maskratio"):
        raw.deint().create_mask(0.0)
    p = ab.default_comb_params()
    assert p.as_list() == [20, 12, 36, 24, 16, 48]


def test_c_abi_header_is_plain_c_and_cxx(tmp_path):
    """include/amtk_b200.h is the drop-in boundary: it must compile as C99 and as C++ with nothing but the standard
    headers (no torch / CUDA types in the signatures)."""
    import subprocess
    hdr = os.path.join(ROOT, "include", "amtk_b200.h")
    for comp, std, lang in (("gcc", "-std=c99", "c"), ("g++", "-std=c++17", "c++")):
        r = subprocess.run([comp, std, "-x", lang, "-fsyntax-only", "-Wall", "-Werror", "-pedantic", hdr],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_host_filter_sources_compile(tmp_path):
    """The C++ mirror of the reference's filter interface compiles against the C ABI alone (no CUDA headers)."""
    import subprocess
    src = os.path.join(ROOT, "tests", "cpp", "test_filters.cpp")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_docs_name_only_declared_entry_points():
    """INTEGRATION.md / DESIGN.md / README.md may only mention amtk_* names that include/amtk_b200.h declares."""
    hdr = open(os.path.join(ROOT, "include", "amtk_b200.h")).read()
    declared = set(re.findall(r"\b(amtk_[a-z0-9_]+)\b", hdr)) | {"amtk_oracle", "amtk_b200", "amtk_internal", "amtk_or", "amtk_comb_avx2"}   # + file stems
    for name in ("INTEGRATION.md", "DESIGN.md", "README.md"):
        used = set(re.findall(r"\b(amtk_[a-z0-9_]+)\b", open(os.path.join(ROOT, name)).read()))
        assert not sorted(u for u in used if u not in declared), name
