"""The combing / field-difference metric is specified by THIS repo (parity with Amatsukaze is
unpinned because the reference delegates it to an external plugin).  These CPU tests pin the C spec oracle two ways:
an independent numpy restatement of the normative text, and hand-computed known answers."""
import numpy as np
import pytest

from oracle import pyoracle as po


def numpy_spec(planes, prev, th6):
    """DESIGN.md section 4, written from the text (vectorised; int64 arithmetic)."""
    out = np.zeros(12, np.int64)
    for cls, idx in ((0, (0,)), (1, (1, 2))):
        thM, thS, thL = th6[3 * cls:3 * cls + 3]
        for k in idx:
            p = planes[k].astype(np.int64)
            q = prev[k].astype(np.int64)
            H = p.shape[0]
            move = np.abs(p - q) >= thM
            comb = np.zeros_like(p)
            valid = np.zeros(H, bool)
            if H >= 5:                               # planes lower than 5 rows have no row with 2 <= y < H-2
                comb[2:H - 2] = np.abs(p[0:H - 4] + 4 * p[2:H - 2] + p[4:H] - 3 * (p[1:H - 3] + p[3:H - 1]))
                valid[2:H - 2] = True
            shima = (comb >= thS) & valid[:, None]
            lshima = (comb >= thL) & valid[:, None]
            for f in (0, 1):
                rows = (np.arange(H) & 1) == f
                out[cls * 6 + f * 3 + 0] += move[rows].sum()
                out[cls * 6 + f * 3 + 1] += shima[rows].sum()
                out[cls * 6 + f * 3 + 2] += lshima[rows].sum()
    return out.astype(np.int32)


@pytest.mark.parametrize("dtype,maxv", [(np.uint8, 255), (np.uint16, 1023), (np.uint16, 65535)])
def test_c_oracle_equals_numpy_restatement(dtype, maxv):
    rng = np.random.default_rng(7)
    for (W, H) in ((32, 16), (48, 10), (16, 6), (64, 34)):
        cur = [rng.integers(0, maxv + 1, (H >> s, W >> s), dtype=np.int64).astype(dtype) for s in (0, 1, 1)]
        prv = [rng.integers(0, maxv + 1, (H >> s, W >> s), dtype=np.int64).astype(dtype) for s in (0, 1, 1)]
        sc = 1 if maxv == 255 else (maxv + 1) // 256
        th6 = [20 * sc, 12 * sc, 36 * sc, 24 * sc, 16 * sc, 48 * sc]
        got = po.or_comb_frame(tuple(cur), tuple(prv), th6)
        assert np.array_equal(np.asarray(got, np.int32), numpy_spec(cur, prv, th6)), (dtype, W, H)


def test_known_answers():
    # a frame of alternating lines 0 / 100 ("perfect combing"): comb = |0+0+0-3*(200)| = 600 on even rows,
    # |100+400+100-0| = 600 on odd rows, for 2 <= y < H-2
    W, H = 16, 12
    Y = np.zeros((H, W), np.uint8)
    Y[1::2] = 100
    U = np.full((H // 2, W // 2), 128, np.uint8)
    V = U.copy()
    th6 = [20, 12, 36, 24, 16, 48]
    c = np.asarray(po.or_comb_frame((Y, U, V), (Y, U, V), th6))
    rows_per_field = (H - 4) // 2
    assert list(c[:6]) == [0, rows_per_field * W, rows_per_field * W, 0, rows_per_field * W, rows_per_field * W]
    assert not c[6:].any()                       # flat chroma, identical previous frame
    # move counts every row (no edge exclusion) and compares with >=
    P = Y.copy()
    P[0, :3] += 20                               # |diff| == thM -> counted (top field)
    P[1, :5] -= 19                               # below threshold -> not counted
    P[H - 1, :7] -= 50                           # last row, bottom field
    c2 = np.asarray(po.or_comb_frame((Y, U, V), (P, U, V), th6))
    assert c2[0] == 3 and c2[3] == 7
    # a flat frame has no combing whatever the thresholds
    F = np.full((H, W), 77, np.uint8)
    assert not np.asarray(po.or_comb_frame((F, U, V), (F, U, V), [1, 1, 1, 1, 1, 1])).any()


def test_first_frame_compares_with_itself_and_fields_partition_rows():
    rng = np.random.default_rng(3)
    N, W, H = 4, 32, 20
    Y = rng.integers(0, 256, (N, H, W), dtype=np.uint8)
    U = rng.integers(0, 256, (N, H // 2, W // 2), dtype=np.uint8)
    V = rng.integers(0, 256, (N, H // 2, W // 2), dtype=np.uint8)
    th6 = [20, 12, 36, 24, 16, 48]
    out = po.or_comb_clip(Y, U, V, th6)
    assert out.shape == (N, 12)
    assert out[0, 0] == 0 and out[0, 3] == 0 and out[0, 6] == 0 and out[0, 9] == 0      # move of frame 0 is zero
    # lshima is a subset of shima when thL >= thS
    assert np.all(out[:, [2, 5, 8, 11]] <= out[:, [1, 4, 7, 10]])
    # thresholds of 1 count every pixel with a non-zero response: top + bottom rows = H-4 rows of luma
    all1 = po.or_comb_clip(Y, U, V, [1, 1, 1, 1, 1, 1])
    assert np.all(all1[:, 1] + all1[:, 4] <= (H - 4) * W)


def test_avx2_form_equals_scalar_spec():
    """oracle/amtk_comb_avx2.c (the vectorised CPU baseline of bench.py; NOT Amatsukaze code) against the scalar
    normative spec and the numpy restatement: ragged widths (vector tails), tiny planes, extreme thresholds."""
    rng = np.random.default_rng(11)
    shapes = ((1920, 1080), (1440, 1080), (352, 270), (96, 36), (50, 22), (34, 6), (16, 4), (70, 10))
    ths = ([20, 12, 36, 24, 16, 48], [1, 1, 1, 1, 1, 1], [128, 2047, 300, 255, 1530, 1531], [0, 0, -5, 300, 40000, 70000],
           [255, 1, 1530, 256, 2, 3])
    for (W, H) in shapes:
        cur = [rng.integers(0, 256, (H >> s, W >> s), dtype=np.int64).astype(np.uint8) for s in (0, 1, 1)]
        # previous frame: mostly close to the current one so that the move threshold splits the pixels
        prv = [np.clip(c.astype(np.int64) + rng.integers(-40, 41, c.shape), 0, 255).astype(np.uint8) for c in cur]
        for th6 in ths:
            a = po.or_comb_frame(tuple(cur), tuple(prv), th6)
            b = po.or_comb_frame(tuple(cur), tuple(prv), th6, "avx2")
            assert np.array_equal(a, b), (W, H, th6, a, b)
        if W <= 352:
            assert np.array_equal(np.asarray(po.or_comb_frame(tuple(cur), tuple(prv), ths[0], "avx2"), np.int32),
                                  numpy_spec(cur, prv, ths[0]))
    # saturated input: every comb response at its maximum
    H, W = 40, 64
    stripes = np.zeros((H, W), np.uint8); stripes[0::2] = 255
    cur = [stripes, stripes[:H // 2, :W // 2].copy(), stripes[:H // 2, :W // 2].copy()]
    prv = [255 - c for c in cur]
    for th6 in ([255, 1530, 1530, 255, 1530, 1530], [255, 1531, 1531, 255, 1531, 1531]):
        assert np.array_equal(po.or_comb_frame(tuple(cur), tuple(prv), th6), po.or_comb_frame(tuple(cur), tuple(prv), th6, "avx2"))


def test_usable_cpu_threads_respects_affinity():
    import os
    n = po.usable_cpu_threads()
    assert 1 <= n <= len(os.sched_getaffinity(0))
