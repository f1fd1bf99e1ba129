"""oracle/pyoracle.py -- TEST INFRASTRUCTURE ONLY (ctypes bindings for the checker libraries).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The product package (amatsukaze_b200) never does.

Two libraries:
  * oracle/_build/libamtk_oracle.so -- this repo's plain-C restatement (oracle/amtk_oracle.c)
  * oracle/_ref/libamtk_ref.so      -- the reference's OWN code compiled from /root/reference by
                                       oracle/build_ref.sh (present only where it was built; travels to the GPU box)
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libamtk_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libamtk_ref.so")

c_float_p = C.POINTER(C.c_float)
c_u8_p = C.POINTER(C.c_uint8)
c_u16_p = C.POINTER(C.c_uint16)
c_i32_p = C.POINTER(C.c_int32)
c_f64_p = C.POINTER(C.c_double)


def build_oracle(force=False):
    """Compile oracle/amtk_oracle.c (gcc, no contraction, no fast-math)."""
    src = os.path.join(HERE, "amtk_oracle.c")
    hdr = os.path.join(HERE, "amtk_oracle.h")
    avx = os.path.join(HERE, "amtk_comb_avx2.c")
    if (not force and os.path.exists(ORACLE_SO)
            and os.path.getmtime(ORACLE_SO) >= max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(avx))):
        return ORACLE_SO
    os.makedirs(os.path.dirname(ORACLE_SO), exist_ok=True)
    obj = os.path.join(os.path.dirname(ORACLE_SO), "amtk_comb_avx2.o")
    # the AVX2 form of the combing spec is its own object so that -mavx2 never touches the float restatement
    subprocess.check_call(["gcc", "-std=gnu99", "-O2", "-fPIC", "-mavx2", "-c", avx, "-o", obj])
    cmd = ["gcc", "-std=c99", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-fopenmp",
           "-o", ORACLE_SO, src, obj, "-lm"]
    subprocess.check_call(cmd)
    return ORACLE_SO


def build_ref():
    """Run oracle/build_ref.sh when the reference tree is present (this container only)."""
    if os.path.exists("/root/reference/Amatsukaze/ComputeKernel.cpp"):
        srcs = [os.path.join(HERE, f) for f in ("build_ref.sh", "ref_glue.cpp", "amtk_oracle.c", "amtk_comb_avx2.c", "shim/ref_shim.h")]
        if not os.path.exists(REF_SO) or os.path.getmtime(REF_SO) < max(os.path.getmtime(f) for f in srcs):
            subprocess.check_call(["bash", os.path.join(HERE, "build_ref.sh")])
    return REF_SO if os.path.exists(REF_SO) else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return a.ctypes.data_as(t)


class _OrLogoStruct(C.Structure):
    _fields_ = [("w", C.c_int), ("h", C.c_int), ("logUVx", C.c_int), ("logUVy", C.c_int),
                ("imgw", C.c_int), ("imgh", C.c_int), ("imgx", C.c_int), ("imgy", C.c_int),
                ("data", c_float_p), ("aY", c_float_p), ("bY", c_float_p), ("aU", c_float_p), ("bU", c_float_p),
                ("aV", c_float_p), ("bV", c_float_p),
                ("mask", c_u8_p), ("maskpixels", C.c_int), ("count", C.c_int),
                ("kernels", c_float_p), ("scales", c_float_p), ("blackScore", C.c_float)]


class _OrScanStruct(C.Structure):
    _fields_ = [("scanw", C.c_int), ("scanh", C.c_int), ("logUVx", C.c_int), ("logUVy", C.c_int),
                ("thy", C.c_int), ("nframes", C.c_int), ("sums", c_f64_p)]


_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        build_oracle()
        L = C.CDLL(ORACLE_SO)
        LP = C.POINTER(_OrLogoStruct)
        SP = C.POINTER(_OrScanStruct)
        L.amtk_or_corr5x5.restype = C.c_float
        L.amtk_or_corr5x5.argtypes = [c_float_p, c_float_p, C.c_int, C.c_int, C.c_int, c_float_p]
        L.amtk_or_corr5x5_scalar_order.restype = C.c_float
        L.amtk_or_corr5x5_scalar_order.argtypes = L.amtk_or_corr5x5.argtypes
        for n, t in (("u8", c_u8_p), ("u16", c_u16_p)):
            for f in ("deint_y", "copy_y"):
                fn = getattr(L, "amtk_or_%s_%s" % (f, n))
                fn.restype = None
                fn.argtypes = [c_float_p, t, C.c_int, C.c_int, C.c_int]
            fn = getattr(L, "amtk_or_scan_frame_" + n)
            fn.restype = None
            fn.argtypes = [LP, t, C.c_int, C.c_float, c_float_p]
            fn = getattr(L, "amtk_or_analyze_frame_" + n)
            fn.restype = None
            fn.argtypes = [LP, LP, LP, t, C.c_int, C.c_float, c_float_p]
            fn = getattr(L, "amtk_or_delogo_" + n)
            fn.restype = None
            fn.argtypes = [t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, c_float_p, c_float_p, C.c_float]
            fn = getattr(L, "amtk_or_comb_frame_" + n)
            fn.restype = None
            fn.argtypes = [t] * 6 + [C.c_int] * 6 + [c_i32_p, c_i32_p]
        L.amtk_or_logo_new.restype = LP
        L.amtk_or_logo_new.argtypes = [C.c_int] * 8 + [c_float_p]
        L.amtk_or_logo_free.restype = None
        L.amtk_or_logo_free.argtypes = [LP]
        L.amtk_or_logo_deint.restype = LP
        L.amtk_or_logo_deint.argtypes = [LP]
        L.amtk_or_logo_field.restype = LP
        L.amtk_or_logo_field.argtypes = [LP, C.c_int]
        L.amtk_or_logo_create_mask.restype = None
        L.amtk_or_logo_create_mask.argtypes = [LP, C.c_float]
        L.amtk_or_logo_corr_score.restype = C.c_float
        L.amtk_or_logo_corr_score.argtypes = [LP, c_float_p, C.c_float]
        L.amtk_or_logo_evaluate.restype = C.c_float
        L.amtk_or_logo_evaluate.argtypes = [LP, c_float_p, C.c_float, C.c_float, c_float_p, C.c_int]
        L.amtk_or_erase_frame_u8.restype = None
        L.amtk_or_erase_frame_u8.argtypes = [LP, c_u8_p, c_u8_p, c_u8_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float]
        L.amtk_or_calc_fade2.restype = None
        L.amtk_or_calc_fade2.argtypes = [c_float_p, C.c_int, C.c_int, C.c_int, c_float_p, c_float_p]
        L.amtk_or_scan_new.restype = SP
        L.amtk_or_scan_new.argtypes = [C.c_int] * 5
        L.amtk_or_scan_free.restype = None
        L.amtk_or_scan_free.argtypes = [SP]
        L.amtk_or_scan_add_frame_u8.restype = C.c_int
        L.amtk_or_scan_add_frame_u8.argtypes = [SP, c_u8_p, c_u8_p, c_u8_p, C.c_int, C.c_int]
        L.amtk_or_scan_get_logo.restype = C.c_int
        L.amtk_or_scan_get_logo.argtypes = [SP, C.c_int, C.c_int, c_float_p]
        L.amtk_or_comb_frame_u8_avx2.restype = None
        L.amtk_or_comb_frame_u8_avx2.argtypes = [c_u8_p] * 6 + [C.c_int] * 6 + [c_i32_p, c_i32_p]
        L.amtk_or_comb_have_avx2.restype = C.c_int
        L.amtk_or_bench_run.restype = C.c_double
        L.amtk_or_bench_run.argtypes = [LP, c_u8_p, C.c_int, C.c_int, C.c_int, c_i32_p, C.c_int, C.c_int, C.c_int, c_float_p, c_i32_p]
        L.amtk_or_bench_scan_comb_u8.restype = C.c_double
        L.amtk_or_bench_scan_comb_u8.argtypes = [LP, c_u8_p, C.c_int, C.c_int, C.c_int, c_i32_p, C.c_int, c_float_p, c_i32_p]
        _oracle = L
    return _oracle


def logo_data_size(w, h, logUVx=1, logUVy=1):
    return (w * h + (w >> logUVx) * (h >> logUVy) * 2) * 2


class OracleLogo:
    """amtk_or_logo handle (LogoDataParam restatement)."""

    def __init__(self, ptr):
        self.L = oracle_lib()
        self.ptr = ptr

    @classmethod
    def create(cls, data, w, h, imgw, imgh, imgx, imgy, logUVx=1, logUVy=1):
        L = oracle_lib()
        d = _f32(data)
        assert d.size == logo_data_size(w, h, logUVx, logUVy)
        return cls(L.amtk_or_logo_new(w, h, logUVx, logUVy, imgw, imgh, imgx, imgy, _p(d, c_float_p)))

    def __del__(self):
        try:
            if self.ptr:
                self.L.amtk_or_logo_free(self.ptr)
                self.ptr = None
        except Exception:
            pass

    @property
    def s(self):
        return self.ptr.contents

    def deint(self):
        return OracleLogo(self.L.amtk_or_logo_deint(self.ptr))

    def field(self, bottom):
        return OracleLogo(self.L.amtk_or_logo_field(self.ptr, int(bottom)))

    def create_mask(self, maskratio):
        self.L.amtk_or_logo_create_mask(self.ptr, C.c_float(maskratio))
        return self

    def data(self):
        s = self.s
        n = logo_data_size(s.w, s.h, s.logUVx, s.logUVy)
        return np.ctypeslib.as_array(s.data, shape=(n,)).copy()

    def mask(self):
        s = self.s
        return np.ctypeslib.as_array(s.mask, shape=(s.h, s.w)).copy()

    def kernels(self):
        s = self.s
        return np.ctypeslib.as_array(s.kernels, shape=(s.count, 25)).copy()

    def scales(self):
        s = self.s
        return np.ctypeslib.as_array(s.scales, shape=(s.count, 32, 2)).copy()

    def evaluate(self, src, maxv, fade, stride=-1):
        s = self.s
        src = _f32(src)
        work = np.zeros(s.w * s.h + 8, np.float32)
        return float(self.L.amtk_or_logo_evaluate(self.ptr, _p(src, c_float_p), C.c_float(maxv), C.c_float(fade),
                                                  _p(work, c_float_p), stride))

    def scan_frame(self, planeY, pitch=None, maxv=None):
        a = np.ascontiguousarray(planeY)
        pitch = a.shape[1] if pitch is None else pitch
        out = np.zeros(2, np.float32)
        if a.dtype == np.uint8:
            self.L.amtk_or_scan_frame_u8(self.ptr, _p(a, c_u8_p), pitch, C.c_float(255.0 if maxv is None else maxv), _p(out, c_float_p))
        else:
            assert a.dtype == np.uint16
            self.L.amtk_or_scan_frame_u16(self.ptr, _p(a, c_u16_p), pitch, C.c_float(1023.0 if maxv is None else maxv), _p(out, c_float_p))
        return out


def or_analyze_frame(dl, ft, fb, planeY, maxv=None, pitch=None):
    L = oracle_lib()
    a = np.ascontiguousarray(planeY)
    pitch = a.shape[1] if pitch is None else pitch
    out = np.zeros(33, np.float32)
    if a.dtype == np.uint8:
        L.amtk_or_analyze_frame_u8(dl.ptr, ft.ptr, fb.ptr, _p(a, c_u8_p), pitch, C.c_float(255.0 if maxv is None else maxv), _p(out, c_float_p))
    else:
        L.amtk_or_analyze_frame_u16(dl.ptr, ft.ptr, fb.ptr, _p(a, c_u16_p), pitch, C.c_float(1023.0 if maxv is None else maxv), _p(out, c_float_p))
    return out


def or_comb_frame(cur, prev, th6, impl="scalar"):
    """cur/prev = (Y,U,V) arrays (2-D, contiguous rows = pitch).  Returns int32[12].
    impl: "scalar" = the normative spec loop, "avx2" = its vectorised form (8-bit only)."""
    L = oracle_lib()
    cy, cu, cv = [np.ascontiguousarray(p) for p in cur]
    py, pu, pv = [np.ascontiguousarray(p) for p in prev]
    h, w = cy.shape
    logx = 0 if cu.shape[1] == w else 1
    logy = 0 if cu.shape[0] == h else 1
    th = np.asarray(th6, np.int32)
    out = np.zeros(12, np.int32)
    if cy.dtype == np.uint8:
        t = c_u8_p
        fn = L.amtk_or_comb_frame_u8_avx2 if impl == "avx2" else L.amtk_or_comb_frame_u8
    else:
        t = c_u16_p
        fn = L.amtk_or_comb_frame_u16
    fn(_p(cy, t), _p(cu, t), _p(cv, t), _p(py, t), _p(pu, t), _p(pv, t), w, h, cy.shape[1], cu.shape[1], logx, logy,
       _p(th, c_i32_p), _p(out, c_i32_p))
    return out


def or_comb_clip(Y, U, V, th6):
    """Y:(N,H,W) U,V:(N,H/2,W/2) -> int32 (N,12) with prev(0)=frame 0."""
    n = Y.shape[0]
    out = np.zeros((n, 12), np.int32)
    for i in range(n):
        j = max(i - 1, 0)
        out[i] = or_comb_frame((Y[i], U[i], V[i]), (Y[j], U[j], V[j]), th6)
    return out


def or_delogo(dst, A, B, fade, maxv, logopitch=None, imgpitch=None, w=None, h=None):
    L = oracle_lib()
    a = dst
    assert a.flags["C_CONTIGUOUS"]
    A = _f32(A)
    B = _f32(B)
    h = a.shape[0] if h is None else h
    w = a.shape[1] if w is None else w
    imgpitch = a.shape[1] if imgpitch is None else imgpitch
    logopitch = w if logopitch is None else logopitch
    if a.dtype == np.uint8:
        L.amtk_or_delogo_u8(_p(a, c_u8_p), w, h, logopitch, imgpitch, C.c_float(maxv), _p(A, c_float_p), _p(B, c_float_p), C.c_float(fade))
    else:
        L.amtk_or_delogo_u16(_p(a, c_u16_p), w, h, logopitch, imgpitch, C.c_float(maxv), _p(A, c_float_p), _p(B, c_float_p), C.c_float(fade))
    return a


def or_erase_frame(logo, Y, U, V, fadeT, fadeB, maxv=255.0):
    L = oracle_lib()
    for p in (Y, U, V):
        assert p.flags["C_CONTIGUOUS"] and p.dtype == np.uint8
    L.amtk_or_erase_frame_u8(logo.ptr, _p(Y, c_u8_p), _p(U, c_u8_p), _p(V, c_u8_p), Y.shape[1], U.shape[1],
                             C.c_float(maxv), C.c_float(fadeT), C.c_float(fadeB))


def or_calc_fade2(records, num_frames, n):
    L = oracle_lib()
    r = _f32(records).reshape(-1, 33)
    ft = C.c_float()
    fb = C.c_float()
    L.amtk_or_calc_fade2(_p(r, c_float_p), r.shape[0], num_frames, n, C.byref(ft), C.byref(fb))
    return ft.value, fb.value


class OracleScan:
    def __init__(self, scanw, scanh, thy, logUVx=1, logUVy=1):
        self.L = oracle_lib()
        self.ptr = self.L.amtk_or_scan_new(scanw, scanh, logUVx, logUVy, thy)
        self.ny = scanw * scanh
        self.nc = (scanw >> logUVx) * (scanh >> logUVy)
        self.n = logo_data_size(scanw, scanh, logUVx, logUVy)

    def __del__(self):
        try:
            self.L.amtk_or_scan_free(self.ptr)
        except Exception:
            pass

    def add_frame(self, y, u, v, pitchY=None, pitchUV=None):
        y, u, v = [np.ascontiguousarray(p, np.uint8) for p in (y, u, v)]
        return self.L.amtk_or_scan_add_frame_u8(self.ptr, _p(y, c_u8_p), _p(u, c_u8_p), _p(v, c_u8_p),
                                                y.shape[1] if pitchY is None else pitchY,
                                                u.shape[1] if pitchUV is None else pitchUV)

    @property
    def nframes(self):
        return self.ptr.contents.nframes

    def sums(self):
        return np.ctypeslib.as_array(self.ptr.contents.sums, shape=(self.ny + 2 * self.nc, 5)).copy()

    def set_sums(self, sums, nframes):
        s = np.ascontiguousarray(sums, np.float64).reshape(-1)
        C.memmove(self.ptr.contents.sums, s.ctypes.data, s.nbytes)
        self.ptr.contents.nframes = nframes

    def get_logo(self, maxv=255, clean=False):
        out = np.zeros(self.n, np.float32)
        ok = self.L.amtk_or_scan_get_logo(self.ptr, maxv, int(clean), _p(out, c_float_p))
        return out if ok else None


# ----------------------------------------------------------------------------------------------------
# The reference's own code (oracle/_ref/libamtk_ref.so)
# ----------------------------------------------------------------------------------------------------
_ref = None


def ref_available():
    return os.path.exists(REF_SO)


def ref_lib():
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            raise RuntimeError("oracle/_ref/libamtk_ref.so missing: run oracle/build_ref.sh where /root/reference exists")
        R = C.CDLL(REF_SO)
        V = C.c_void_p
        R.ref_corr5x5_avx.restype = C.c_float
        R.ref_corr5x5_avx.argtypes = [c_float_p, c_float_p, C.c_int, C.c_int, C.c_int, c_float_p]
        R.ref_corr5x5_scalar.restype = C.c_float
        R.ref_corr5x5_scalar.argtypes = R.ref_corr5x5_avx.argtypes
        R.ref_is_avx.restype = C.c_int
        R.ref_logo_create.restype = V
        R.ref_logo_create.argtypes = [C.c_int] * 8 + [c_float_p]
        R.ref_logo_free.argtypes = [V]
        R.ref_logo_deint.restype = V
        R.ref_logo_deint.argtypes = [V]
        R.ref_logo_field.restype = V
        R.ref_logo_field.argtypes = [V, C.c_int]
        R.ref_logo_create_mask.argtypes = [V, C.c_float]
        R.ref_logo_dims.argtypes = [V, C.POINTER(C.c_int)]
        R.ref_logo_get_data.argtypes = [V, c_float_p]
        R.ref_logo_black_score.restype = C.c_float
        R.ref_logo_black_score.argtypes = [V]
        R.ref_logo_get_mask.argtypes = [V, c_u8_p]
        R.ref_logo_get_kernels.argtypes = [V, c_float_p, C.c_int]
        R.ref_logo_get_scales.argtypes = [V, c_float_p, C.c_int]
        R.ref_logo_evaluate.restype = C.c_float
        R.ref_logo_evaluate.argtypes = [V, c_float_p, C.c_float, C.c_float, c_float_p, C.c_int]
        R.ref_logo_corr_score.restype = C.c_float
        R.ref_logo_corr_score.argtypes = [V, c_float_p, C.c_float]
        R.ref_logo_save.restype = C.c_int
        R.ref_logo_save.argtypes = [V, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
        R.ref_logo_load.restype = V
        R.ref_logo_load.argtypes = [C.c_char_p, C.c_void_p]
        R.ref_sizeof.restype = C.c_int
        R.ref_sizeof.argtypes = [C.c_int]
        for n, t in (("u8", c_u8_p), ("u16", c_u16_p)):
            for f in ("deint_y", "copy_y"):
                fn = getattr(R, "ref_%s_%s" % (f, n))
                fn.restype = None
                fn.argtypes = [c_float_p, t, C.c_int, C.c_int, C.c_int]
            fn = getattr(R, "ref_scan_add_frame_" + n)
            fn.restype = C.c_int
            fn.argtypes = [V, t, t, t, C.c_int, C.c_int]
        R.ref_scan_create.restype = V
        R.ref_scan_create.argtypes = [C.c_int] * 5
        R.ref_scan_free.argtypes = [V]
        R.ref_scan_nframes.restype = C.c_int
        R.ref_scan_nframes.argtypes = [V]
        R.ref_scan_get_sums.argtypes = [V, c_f64_p]
        R.ref_scan_set_sums.argtypes = [V, c_f64_p, C.c_int]
        R.ref_scan_normalize.argtypes = [V, C.c_int]
        R.ref_scan_get_logo.restype = C.c_int
        R.ref_scan_get_logo.argtypes = [V, C.c_int, c_float_p]
        R.ref_logoframe_write.restype = C.c_int
        R.ref_logoframe_write.argtypes = [c_float_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_int), c_float_p]
        if hasattr(R, "ref_delogo_u8"):                         # round 2: AMTEraseLogo::Delogo / CalcFade2 (LogoScan.hpp:1248-1315)
            R.ref_delogo_u8.restype = None
            R.ref_delogo_u8.argtypes = [c_u8_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, c_float_p, c_float_p, C.c_float]
            R.ref_delogo_u16.restype = None
            R.ref_delogo_u16.argtypes = [c_u16_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, c_float_p, c_float_p, C.c_float]
            R.ref_calc_fade2.restype = None
            R.ref_calc_fade2.argtypes = [c_float_p, C.c_int, C.c_int, C.c_int, c_float_p, c_float_p]
        if hasattr(R, "ref_merge_field_u8"):                    # round 2: AMTSource::MergeField (AMTSource.hpp:291-355)
            R.ref_merge_field_u8.restype = None
            R.ref_merge_field_u8.argtypes = [c_u8_p, c_u8_p, c_u8_p, C.c_int, C.c_int, c_u8_p, c_u8_p, c_u8_p, C.c_int, C.c_int,
                                             c_u8_p, c_u8_p, c_u8_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        if hasattr(R, "ref_analyze_getframe"):                  # round 2: AMTAnalyzeLogo::GetFrameT / LogoFrame::ScanFrame themselves
            R.ref_analyze_getframe.restype = None
            R.ref_analyze_getframe.argtypes = [V, V, V, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_float_p]
            R.ref_scan_frame.restype = None
            R.ref_scan_frame.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, c_float_p, c_float_p, c_float_p]
        if hasattr(R, "ref_read_timecode"):                     # round 2: FilteredSource.hpp side-file readers
            R.ref_read_timecode.restype = C.c_int
            R.ref_read_timecode.argtypes = [C.c_char_p, c_f64_p, C.c_int, C.POINTER(C.c_int)]
            R.ref_decimate_map.restype = C.c_int
            R.ref_decimate_map.argtypes = [C.c_char_p, C.c_int, c_i32_p, C.c_int, C.c_char_p, C.c_int]
        if hasattr(R, "ref_erase_fades"):
            R.ref_erase_fades.restype = C.c_int
            R.ref_erase_fades.argtypes = [c_float_p, C.c_int, C.c_int, C.c_char_p, C.c_int, c_float_p, c_i32_p, C.c_char_p, C.c_int]
        R.ref_bench_create.restype = V
        R.ref_bench_create.argtypes = [V, C.c_int, C.c_int, C.c_int]
        R.ref_bench_free.argtypes = [V]
        R.ref_bench_run.restype = C.c_double
        R.ref_bench_run.argtypes = [V, c_u8_p, C.c_int, c_i32_p, C.c_int, C.c_int, c_float_p, c_i32_p]
        R.ref_bench_scan_comb_u8.restype = C.c_double
        R.ref_bench_scan_comb_u8.argtypes = [V, c_u8_p, C.c_int, C.c_int, C.c_int, c_i32_p, C.c_int, c_float_p, c_i32_p]
        _ref = R
    return _ref


def ref_has_erase():
    return ref_available() and hasattr(ref_lib(), "ref_delogo_u8")


def ref_delogo(dst, A, B, fade, maxv, logopitch=None, imgpitch=None, w=None, h=None):
    """The reference's own AMTEraseLogo::Delogo (LogoScan.hpp:1248-1261), in place on a C-contiguous u8/u16 array."""
    a = dst
    assert a.flags["C_CONTIGUOUS"]
    A, B = _f32(A), _f32(B)
    h = a.shape[0] if h is None else h
    w = a.shape[1] if w is None else w
    imgpitch = a.shape[1] if imgpitch is None else imgpitch
    logopitch = w if logopitch is None else logopitch
    if a.dtype == np.uint8:
        ref_lib().ref_delogo_u8(_p(a, c_u8_p), w, h, logopitch, imgpitch, C.c_float(maxv), _p(A, c_float_p), _p(B, c_float_p), C.c_float(fade))
    else:
        ref_lib().ref_delogo_u16(_p(a, c_u16_p), w, h, logopitch, imgpitch, C.c_float(maxv), _p(A, c_float_p), _p(B, c_float_p), C.c_float(fade))
    return a


def ref_calc_fade2(records, num_frames, n):
    """The reference's own AMTEraseLogo::CalcFade2 (LogoScan.hpp:1263-1315) over an analyze clip built from `records`
    ((num_frames, 33) floats) the way AMTAnalyzeLogo lays it out: frame k = records of source frames 8k..8k+7, clamped to
    the last source frame (:1133); GetFrame(n) outside the clip is clamped to its range (AviSynth's contract)."""
    r = _f32(records).reshape(-1, 33)
    N = r.shape[0]
    nblocks = (N + 7) // 8
    idx = np.minimum(np.arange(nblocks * 8), N - 1)
    blocks = np.ascontiguousarray(r[idx])
    ft, fb = C.c_float(), C.c_float()
    ref_lib().ref_calc_fade2(_p(blocks, c_float_p), nblocks, int(num_frames), int(n), C.byref(ft), C.byref(fb))
    return ft.value, fb.value


def _analyze_blocks(records):
    r = _f32(records).reshape(-1, 33)
    N = r.shape[0]
    nblocks = (N + 7) // 8
    return np.ascontiguousarray(r[np.minimum(np.arange(nblocks * 8), N - 1)]), nblocks


def ref_erase_fades(records, num_frames, logof_path=None, max_fade_length=16):
    """The reference's own AMTEraseLogo fade selection for every frame: ReadLogoFrameFile (LogoScan.hpp:1421-1461) when a
    logoframe file is given, then CalcFade (:1317-1341, which falls back to CalcFade2).  Returns (fades (N,2) float32,
    frameResult (N,) int32 or None); raises RuntimeError with the reference's ThrowError text."""
    blocks, nblocks = _analyze_blocks(records)
    out = np.zeros((num_frames, 2), np.float32)
    fr = np.zeros(num_frames, np.int32)
    err = C.create_string_buffer(512)
    ok = ref_lib().ref_erase_fades(_p(blocks, c_float_p), nblocks, int(num_frames), str(logof_path).encode() if logof_path else None,
                                   int(max_fade_length), _p(out, c_float_p), _p(fr, c_i32_p), err, 512)
    if not ok:
        raise RuntimeError(err.value.decode("utf-8", "replace"))
    return out, (fr if logof_path else None)


def ref_has_drivers():
    return ref_available() and hasattr(ref_lib(), "ref_analyze_getframe")


def ref_analyze_getframe(dl, ft, fb, frames, w, h, n, bits=8):
    """The reference's own AMTAnalyzeLogo::GetFrameT (LogoScan.hpp:1119-1161) for analyze frame n: the 8 LogoAnalyzeFrame
    records (8, 33) of source frames 8n..8n+7 (clamped to the last one, :1133).  frames: (N, w*h*3/2) packed planar 4:2:0."""
    fr = np.ascontiguousarray(frames)
    d = dl.dims()
    out = np.zeros((8, 33), np.float32)
    ref_lib().ref_analyze_getframe(dl.ptr, ft.ptr, fb.ptr, d["imgx"], d["imgy"], fr.ctypes.data, fr.shape[0], w, h, bits, int(n), _p(out, c_float_p))
    return out


def ref_scan_frame_code(logos, frame, w, h, bits=8):
    """The reference's own LogoFrame::ScanFrame (LogoScan.hpp:1543-1568) on one packed frame; logos: deint RefLogo objects
    or None (an invalid logo -> corr0 = 0, corr1 = -1).  Returns (len(logos), 2) float32."""
    fr = np.ascontiguousarray(frame)
    arr = (C.c_void_p * len(logos))(*[(lg.ptr if lg is not None else None) for lg in logos])
    big = max([lg.dims()["w"] * lg.dims()["h"] for lg in logos if lg is not None] + [1])
    mem_d, mem_w = np.zeros(big + 8, np.float32), np.zeros(big + 8, np.float32)
    out = np.zeros((len(logos), 2), np.float32)
    ref_lib().ref_scan_frame(arr, len(logos), fr.ctypes.data, w, h, bits, _p(mem_d, c_float_p), _p(mem_w, c_float_p), _p(out, c_float_p))
    return out


def ref_has_sidefiles():
    return ref_available() and hasattr(ref_lib(), "ref_read_timecode")


def ref_read_timecode(path):
    """The reference's own AMTFilterSource::readTimecodeFile + base-fps estimate (FilteredSource.hpp:163-188,197-210).
    Returns (timeCodes list, vfrTimingFps) or None when the file cannot be opened."""
    out = np.zeros(1 << 16, np.float64)
    fps = C.c_int(0)
    n = ref_lib().ref_read_timecode(str(path).encode(), _p(out, c_f64_p), out.size, C.byref(fps))
    return None if n < 0 else (out[:n].tolist(), fps.value)


def ref_decimate_map(duration_path, num_source_frames):
    """The reference's own AMTDecimate (FilteredSource.hpp:645-660,663-666): source frame of every output frame.  Raises
    RuntimeError with the ThrowError text on a frame-count mismatch, IOError when the file cannot be opened."""
    m = np.zeros(1 << 16, np.int32)
    err = C.create_string_buffer(512)
    n = ref_lib().ref_decimate_map(str(duration_path).encode(), int(num_source_frames), _p(m, c_i32_p), m.size, err, 512)
    if n == -1:
        raise IOError("cannot open " + str(duration_path))
    if n == -2:
        raise RuntimeError(err.value.decode("utf-8", "replace"))
    return m[:n].tolist()


def ref_has_mergefield():
    return ref_available() and hasattr(ref_lib(), "ref_merge_field_u8")


def ref_merge_field(top, bottom, w, h, nv12=False):
    """The reference's own AMTSource::MergeField (AMTSource.hpp:291-355) on two packed 8-bit 4:2:0 frames (1-D uint8 arrays:
    Y then U, V planar -- or Y then interleaved UV when nv12): even rows from `top`, odd rows from `bottom`; returns the packed
    planar YV12 frame."""
    t = np.ascontiguousarray(top, np.uint8)
    b = np.ascontiguousarray(bottom, np.uint8)
    ysz, cw, ch = w * h, w // 2, h // 2
    out = np.zeros(ysz + 2 * cw * ch, np.uint8)

    def at(a, off):
        return C.cast(a.ctypes.data + off, c_u8_p)
    if nv12:
        tu, tv, bu, bv, spuv = at(t, ysz), at(t, ysz), at(b, ysz), at(b, ysz), w
    else:
        tu, tv, bu, bv, spuv = at(t, ysz), at(t, ysz + cw * ch), at(b, ysz), at(b, ysz + cw * ch), cw
    ref_lib().ref_merge_field_u8(at(out, 0), at(out, ysz), at(out, ysz + cw * ch), w, cw,
                                 at(t, 0), tu, tv, w, spuv, at(b, 0), bu, bv, w, spuv, w, h, 1 if nv12 else 0)
    return out


def ref_logoframe(eval_results, frames_per_sec, outpath=None, num_candidates=-1):
    """The reference's LogoFrame::selectLogo (+ writeResult when outpath is given) on an (N, L, 2) score array.
    Returns (bestLogo, logoRatio)."""
    e = _f32(eval_results)
    n, nl = e.shape[0], e.shape[1]
    best, ratio = C.c_int(), C.c_float()
    ok = ref_lib().ref_logoframe_write(_p(e, c_float_p), n, nl, frames_per_sec, num_candidates,
                                       outpath.encode() if outpath else None, C.byref(best), C.byref(ratio))
    assert ok
    return best.value, ratio.value


def usable_cpu_threads():
    """Host threads this process may really use: the affinity mask capped by the cgroup CPU quota (os.cpu_count()
    ignores both, which oversubscribed the round-1 CPU arm)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", ):
        try:
            quota, period = open(path).read().split()[:2]
            if quota != "max":
                n = min(n, max(1, int(float(quota) / float(period) + 0.5)))
        except Exception:
            pass
    try:
        q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
        per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
        if q > 0 and per > 0:
            n = min(n, max(1, int(q / per + 0.5)))
    except Exception:
        pass
    return max(1, n)


class CpuBench:
    """Bounded CPU runs of the fused hot path (ScanFrame + comb) over packed-YV12 frames (numpy uint8 (n, w*h*3/2)).
    Logo half: the reference's own compiled code when oracle/_ref exists (kind "reference"), else the plain-C port
    ("port").  Combing half: this repo's spec, scalar or AVX2 (comb_impl) -- not Amatsukaze code either way.
    The thread team and all scratch are created here, outside any timed region."""

    def __init__(self, w, h, logo_data, imgx, imgy, nthreads, maskratio=0.35, logo_w=64, logo_h=64):
        self.w, self.h, self.nthreads = w, h, int(nthreads)
        self.kind = "reference" if ref_available() else "port"
        if self.kind == "reference":
            self.logo = RefLogo.create(logo_data, logo_w, logo_h, w, h, imgx, imgy).deint().create_mask(maskratio)
            self.hb = C.c_void_p(ref_lib().ref_bench_create(self.logo.ptr, w, h, self.nthreads))
        else:
            self.logo = OracleLogo.create(logo_data, logo_w, logo_h, w, h, imgx, imgy).deint().create_mask(maskratio)
            self.hb = None
        self.comb_avx2 = bool(oracle_lib().amtk_or_comb_have_avx2())

    def close(self):
        if self.hb:
            ref_lib().ref_bench_free(self.hb)
            self.hb = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, frames, th6, mode=3, comb_impl="avx2"):
        """One pass.  mode: 1 logo only, 2 comb only, 3 both.  Returns (seconds, scores (n,2), counts (n,12))."""
        fr = frames if (isinstance(frames, np.ndarray) and frames.dtype == np.uint8 and frames.flags["C_CONTIGUOUS"]) \
            else np.ascontiguousarray(frames, np.uint8)
        n = fr.shape[0]
        scores = np.zeros((n, 2), np.float32)
        counts = np.zeros((n, 12), np.int32)
        th = np.asarray(th6, np.int32)
        ci = 1 if comb_impl == "avx2" else 0
        if self.kind == "reference":
            sec = ref_lib().ref_bench_run(self.hb, _p(fr, c_u8_p), n, _p(th, c_i32_p), mode, ci, _p(scores, c_float_p), _p(counts, c_i32_p))
        else:
            sec = oracle_lib().amtk_or_bench_run(self.logo.ptr, _p(fr, c_u8_p), n, self.w, self.h, _p(th, c_i32_p), self.nthreads,
                                                 mode, ci, _p(scores, c_float_p), _p(counts, c_i32_p))
        return sec, scores, counts


def cpu_scan_comb(frames, w, h, logo_data, imgx, imgy, th6, nthreads, maskratio=0.35, logo_w=64, logo_h=64, comb_impl="scalar"):
    """One-shot form: returns (seconds, scores (n,2), counts (n,12), kind)."""
    b = CpuBench(w, h, logo_data, imgx, imgy, nthreads, maskratio, logo_w, logo_h)
    sec, scores, counts = b.run(frames, th6, 3, comb_impl)
    b.close()
    return sec, scores, counts, b.kind


class RefLogo:
    """The reference's LogoDataParam (compiled from /root/reference)."""

    def __init__(self, ptr):
        self.R = ref_lib()
        self.ptr = C.c_void_p(ptr)

    @classmethod
    def create(cls, data, w, h, imgw, imgh, imgx, imgy, logUVx=1, logUVy=1):
        R = ref_lib()
        d = _f32(data)
        assert d.size == logo_data_size(w, h, logUVx, logUVy)
        return cls(R.ref_logo_create(w, h, logUVx, logUVy, imgw, imgh, imgx, imgy, _p(d, c_float_p)))

    @classmethod
    def load(cls, path):
        R = ref_lib()
        hdr = np.zeros(540, np.uint8)
        p = R.ref_logo_load(path.encode(), hdr.ctypes.data_as(C.c_void_p))
        if not p:
            return None, None
        return cls(p), hdr

    def save(self, path, imgw, imgh, imgx, imgy, name="No Name", service_id=0):
        return self.R.ref_logo_save(self.ptr, path.encode(), imgw, imgh, imgx, imgy, name.encode(), service_id)

    def __del__(self):
        try:
            self.R.ref_logo_free(self.ptr)
        except Exception:
            pass

    def dims(self):
        v = (C.c_int * 10)()
        self.R.ref_logo_dims(self.ptr, v)
        return dict(zip(("w", "h", "logUVx", "logUVy", "imgw", "imgh", "imgx", "imgy", "maskpixels"), list(v)[:9]))

    def deint(self):
        return RefLogo(self.R.ref_logo_deint(self.ptr))

    def field(self, bottom):
        return RefLogo(self.R.ref_logo_field(self.ptr, int(bottom)))

    def create_mask(self, maskratio):
        self.R.ref_logo_create_mask(self.ptr, C.c_float(maskratio))
        return self

    def data(self):
        d = self.dims()
        out = np.zeros(logo_data_size(d["w"], d["h"], d["logUVx"], d["logUVy"]), np.float32)
        self.R.ref_logo_get_data(self.ptr, _p(out, c_float_p))
        return out

    def black_score(self):
        return float(self.R.ref_logo_black_score(self.ptr))

    def mask(self):
        d = self.dims()
        out = np.zeros((d["h"], d["w"]), np.uint8)
        self.R.ref_logo_get_mask(self.ptr, _p(out, c_u8_p))
        return out

    def visited_count(self):
        m = self.mask()
        return int(m[2:-2, 2:-2].sum())

    def kernels(self):
        n = self.visited_count()
        out = np.zeros((n, 25), np.float32)
        self.R.ref_logo_get_kernels(self.ptr, _p(out, c_float_p), n)
        return out

    def scales(self):
        n = self.visited_count()
        out = np.zeros((n, 32, 2), np.float32)
        self.R.ref_logo_get_scales(self.ptr, _p(out, c_float_p), n)
        return out

    def evaluate(self, src, maxv, fade, stride=-1):
        d = self.dims()
        src = _f32(src)
        work = np.zeros(d["w"] * d["h"] + 8, np.float32)
        return float(self.R.ref_logo_evaluate(self.ptr, _p(src, c_float_p), C.c_float(maxv), C.c_float(fade),
                                              _p(work, c_float_p), stride))


def ref_deint_y(plane, w, h, off=0, pitch=None):
    R = ref_lib()
    a = np.ascontiguousarray(plane)
    pitch = a.shape[1] if pitch is None else pitch
    out = np.zeros(w * h + 8, np.float32)
    flat = a.reshape(-1)[off:]
    if a.dtype == np.uint8:
        R.ref_deint_y_u8(_p(out, c_float_p), flat.ctypes.data_as(c_u8_p), pitch, w, h)
    else:
        R.ref_deint_y_u16(_p(out, c_float_p), flat.ctypes.data_as(c_u16_p), pitch, w, h)
    return out


def ref_copy_y(plane, w, h, off=0, pitch=None):
    R = ref_lib()
    a = np.ascontiguousarray(plane)
    pitch = a.shape[1] if pitch is None else pitch
    out = np.zeros(w * h + 8, np.float32)
    flat = a.reshape(-1)[off:]
    if a.dtype == np.uint8:
        R.ref_copy_y_u8(_p(out, c_float_p), flat.ctypes.data_as(c_u8_p), pitch, w, h)
    else:
        R.ref_copy_y_u16(_p(out, c_float_p), flat.ctypes.data_as(c_u16_p), pitch, w, h)
    return out


def ref_scan_frame(deint_logo, planeY, maxv=None, pitch=None):
    """LogoFrame::ScanFrame (LogoScan.hpp:1559-1566) composed from the reference's own DeintY + EvaluateLogo."""
    d = deint_logo.dims()
    a = np.ascontiguousarray(planeY)
    pitch = a.shape[1] if pitch is None else pitch
    maxv = (255.0 if a.dtype == np.uint8 else 1023.0) if maxv is None else maxv
    de = ref_deint_y(a, d["w"], d["h"], off=d["imgx"] + d["imgy"] * pitch, pitch=pitch)
    return np.array([deint_logo.evaluate(de, maxv, 0.0), deint_logo.evaluate(de, maxv, 1.0)], np.float32)


def ref_analyze_frame(dl, ft, fb, planeY, maxv=None, pitch=None):
    """AMTAnalyzeLogo::GetFrameT per source frame (LogoScan.hpp:1146-1155) from the reference's own pieces."""
    d = dl.dims()
    a = np.ascontiguousarray(planeY)
    pitch = a.shape[1] if pitch is None else pitch
    maxv = (255.0 if a.dtype == np.uint8 else 1023.0) if maxv is None else maxv
    off = d["imgx"] + d["imgy"] * pitch
    w, h = d["w"], d["h"]
    cp = ref_copy_y(a, w, h, off=off, pitch=pitch)
    de = ref_deint_y(a, w, h, off=off, pitch=pitch)
    out = np.zeros(33, np.float32)
    for f in range(11):
        fade = np.float32(f) / np.float32(10.0)
        out[f] = abs(np.float32(dl.evaluate(de, maxv, fade)))
        out[11 + f] = abs(np.float32(ft.evaluate(cp, maxv, fade, stride=2 * w)))
        out[22 + f] = abs(np.float32(fb.evaluate(cp[w:], maxv, fade, stride=2 * w)))
    return out


class RefScan:
    def __init__(self, scanw, scanh, thy, logUVx=1, logUVy=1):
        self.R = ref_lib()
        self.ptr = C.c_void_p(self.R.ref_scan_create(scanw, scanh, logUVx, logUVy, thy))
        self.ny = scanw * scanh
        self.nc = (scanw >> logUVx) * (scanh >> logUVy)
        self.n = logo_data_size(scanw, scanh, logUVx, logUVy)

    def __del__(self):
        try:
            self.R.ref_scan_free(self.ptr)
        except Exception:
            pass

    def add_frame(self, y, u, v, pitchY=None, pitchUV=None):
        y, u, v = [np.ascontiguousarray(p, np.uint8) for p in (y, u, v)]
        return self.R.ref_scan_add_frame_u8(self.ptr, _p(y, c_u8_p), _p(u, c_u8_p), _p(v, c_u8_p),
                                            y.shape[1] if pitchY is None else pitchY,
                                            u.shape[1] if pitchUV is None else pitchUV)

    @property
    def nframes(self):
        return self.R.ref_scan_nframes(self.ptr)

    def sums(self):
        out = np.zeros((self.ny + 2 * self.nc, 5), np.float64)
        self.R.ref_scan_get_sums(self.ptr, _p(out, c_f64_p))
        return out

    def set_sums(self, sums, nframes):
        s = np.ascontiguousarray(sums, np.float64)
        self.R.ref_scan_set_sums(self.ptr, _p(s, c_f64_p), nframes)

    def get_logo(self, maxv=255, clean=False):
        """Normalize(maxv) then GetLogo(clean) on a COPY of the accumulators (Normalize is destructive)."""
        saved = self.sums()
        n = self.nframes
        self.R.ref_scan_normalize(self.ptr, maxv)
        out = np.zeros(self.n, np.float32)
        ok = self.R.ref_scan_get_logo(self.ptr, int(clean), _p(out, c_float_p))
        self.set_sums(saved, n)
        return out if ok else None
