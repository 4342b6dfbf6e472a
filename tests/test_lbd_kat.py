"""Known-answer tests for the LBD descriptor (BinaryDescriptor::computeLBD + binaryConversion of OpenCV-contrib
line_descriptor, called by LineSegment::ExtractLineSegment, reference include/ExtractLineSegment.h:38).

No executable reference of LBD exists in this environment ("parity unpinned").  What CAN be derived
by hand is the descriptor of a horizontal line on a horizontal step edge: the Sobel response is zero everywhere except on
the two rows next to the edge, where dy = 4 (b - a) and dx = 0, so of the 63 rows of the line support region exactly two
carry a row sum, len * 4 |b - a| * gaussG[row], in the "positive (or negative) gradient across the line" statistic.  The
72 band statistics, the normalisation and the 256 comparisons follow from that in closed form (below, in float32).  The
oracle (CPU test) and the CUDA kernel k_lbd (GPU test, through plslam_lines_compute_lbd) must reproduce it bit for bit.
"""
import numpy as np
import pytest

f32 = np.float32
LBD_W, BANDS, ROWS = 7, 9, 63
COMB = [(0, 1), (0, 2), (0, 3), (0, 4), (0, 5), (0, 6), (1, 2), (1, 3), (1, 4), (1, 5), (1, 6), (2, 3), (2, 4), (2, 5), (2, 6),
        (2, 7), (2, 8), (3, 4), (3, 5), (3, 6), (3, 7), (3, 8), (4, 5), (4, 6), (4, 7), (4, 8), (5, 6), (5, 7), (5, 8), (6, 7),
        (6, 8), (7, 8)]


def gauss_tables():
    u, sg = (LBD_W * 3 - 1) // 2, (LBD_W * 2 + 1) // 2  # 10, 7 (integer divisions of the BinaryDescriptor constructor)
    gl = np.array([np.exp((i - u) ** 2 * (-1.0 / (2.0 * sg * sg))) for i in range(LBD_W * 3)]).astype(f32)
    u = (BANDS * LBD_W - 1) // 2  # 31
    gg = np.array([np.exp((i - u) ** 2 * (-1.0 / (2.0 * u * u))) for i in range(ROWS)]).astype(f32)
    return gl, gg


def expected_lbd(length, amp, rows, positive):
    """72 floats + 32 bytes of a line whose support region has gradient `amp` across the line on `rows` only."""
    gl, gg = gauss_tables()
    rowsum = np.zeros((ROWS, 4), f32)  # pgdL, ngdL, pgdO, ngdO
    s = f32(0)
    for _ in range(length):
        s = f32(s + f32(amp))
    for h in rows:
        rowsum[h][2 if positive else 3] = f32(gg[h] * s)
    des = np.zeros(72, f32)
    inv2, inv3 = f32(1.0 / (LBD_W * 2.0)), f32(1.0 / (LBD_W * 3.0))
    for b in range(BANDS):
        acc = np.zeros(8, f32)  # q: 0 pgdL, 1 ngdL, 2 pgdL^2, 3 ngdL^2, 4 pgdO, 5 ngdO, 6 pgdO^2, 7 ngdO^2
        for h in range(max(0, (b - 1) * LBD_W), min(ROWS, (b + 2) * LBD_W)):
            b0, m = divmod(h, LBD_W)
            coef = gl[m + LBD_W] if b == b0 else (gl[m + 2 * LBD_W] if b == b0 - 1 else gl[m])
            for q in range(8):
                v = rowsum[h][(q & 1) + (2 if q & 4 else 0)]
                if q & 2:
                    acc[q] = f32(acc[q] + f32(f32(coef * coef) * f32(v * v)))
                else:
                    acc[q] = f32(acc[q] + f32(coef * v))
        invn = inv2 if b in (0, BANDS - 1) else inv3
        for k, (lin, sq) in enumerate(((0, 2), (1, 3), (4, 6), (5, 7))):
            mean = f32(acc[lin] * invn)
            des[b * 8 + k] = mean
            des[b * 8 + 4 + k] = np.sqrt(f32(f32(acc[sq] * invn) - f32(mean * mean)))
    tm = ts = f32(0)
    for b in range(BANDS):
        for q in range(4):
            tm = f32(tm + f32(des[b * 8 + q] * des[b * 8 + q]))
        for q in range(4, 8):
            ts = f32(ts + f32(des[b * 8 + q] * des[b * 8 + q]))
    tm, ts = f32(f32(1) / np.sqrt(tm)), f32(f32(1) / np.sqrt(ts))
    for b in range(BANDS):
        for q in range(4):
            des[b * 8 + q] = f32(des[b * 8 + q] * tm)
        for q in range(4, 8):
            des[b * 8 + q] = f32(des[b * 8 + q] * ts)
    des = np.where(des.astype(np.float64) > 0.4, f32(0.4), des).astype(f32)
    t = f32(0)
    for v in des:
        t = f32(t + f32(v * v))
    t = f32(f32(1) / np.sqrt(t))
    des = (des * t).astype(f32)
    out = np.zeros(32, np.uint8)
    for c, (x, y) in enumerate(COMB):
        out[c] = sum(1 << i for i in range(8) if des[x * 8 + i] > des[y * 8 + i])
    return des, out


def step_image(H, W, edge_row, a, b):
    img = np.full((H, W), a, np.uint8)
    img[edge_row:] = b
    return img


def keyline(dtype, x0, x1, y):
    kl = np.zeros(1, dtype)
    kl["startPointX"] = kl["sPointInOctaveX"] = x0
    kl["endPointX"] = kl["ePointInOctaveX"] = x1
    kl["startPointY"] = kl["sPointInOctaveY"] = kl["endPointY"] = kl["ePointInOctaveY"] = y
    kl["angle"] = 0.0
    kl["numOfPixels"] = abs(x1 - x0) + 1
    kl["lineLength"] = abs(x1 - x0)
    kl["pt_x"], kl["pt_y"] = (x0 + x1) / 2.0, y
    return kl


# (edge row, line row, first x, last x, grey above, grey below)
CASES = [(240, 240, 100, 160, 50, 200), (240, 240, 300, 420, 200, 50), (200, 196, 40, 100, 10, 250), (300, 310, 500, 560, 90, 130),
         (240, 252, 100, 180, 0, 255), (100, 79, 20, 60, 255, 0)]


def _expect(case):
    e, y, x0, x1, a, b = case
    rows = [e - 1 - (y - 31), e - (y - 31)]  # support-region rows of image rows e - 1 and e
    assert all(0 <= r < ROWS for r in rows)
    return expected_lbd(x1 - x0 + 1, 4 * abs(b - a), rows, b > a)


@pytest.mark.parametrize("case", CASES)
def test_oracle_lbd_equals_the_hand_derived_descriptor(oracle, case):
    e, y, x0, x1, a, b = case
    img = step_image(480, 640, e, a, b)
    d32, d72 = oracle.lbd(img, keyline(oracle.KEYLINE_DTYPE, x0, x1, y))
    des, out = _expect(case)
    assert np.array_equal(d72[0].view(np.uint32), des.view(np.uint32)), "72 band statistics"
    assert np.array_equal(d32[0], out), "32 descriptor bytes"


def test_the_descriptor_tells_the_cases_apart(oracle):
    outs = [_expect(c)[1].tobytes() for c in CASES]
    assert len(set(outs)) >= 4  # polarity and the edge's offset from the line change the bytes


@pytest.mark.gpu
def test_k_lbd_equals_the_hand_derived_descriptor(oracle):
    import plslam_b200 as pl
    ls = pl.LineSegment()
    for case in CASES:
        e, y, x0, x1, a, b = case
        img = step_image(480, 640, e, a, b)
        got = ls.compute_lbd(img, keyline(pl.KEYLINE_DTYPE, x0, x1, y))
        assert np.array_equal(got[0], _expect(case)[1]), case


@pytest.mark.gpu
def test_k_lbd_on_given_keylines_equals_the_oracle(oracle):
    """plslam_lines_compute_lbd on hand-made key lines of every orientation, some crossing the image border (reflected
    Sobel taps, clamped sampling positions), against the oracle."""
    import plslam_b200 as pl
    from plslam_b200.synth import synth_frame
    img = synth_frame(5)
    rng = np.random.default_rng(1)
    kls = np.zeros(64, pl.KEYLINE_DTYPE)
    for i in range(len(kls)):
        x0, y0 = rng.uniform(-5, 645), rng.uniform(-5, 485)
        ang = rng.uniform(-np.pi, np.pi)
        ln = rng.uniform(8, 200)
        x1, y1 = x0 + ln * np.cos(ang), y0 + ln * np.sin(ang)
        x0, x1 = np.clip([x0, x1], 0, 639)
        y0, y1 = np.clip([y0, y1], 0, 479)
        kls[i]["startPointX"] = kls[i]["sPointInOctaveX"] = x0
        kls[i]["startPointY"] = kls[i]["sPointInOctaveY"] = y0
        kls[i]["endPointX"] = kls[i]["ePointInOctaveX"] = x1
        kls[i]["endPointY"] = kls[i]["ePointInOctaveY"] = y1
        kls[i]["angle"] = np.arctan2(np.float32(y1) - np.float32(y0), np.float32(x1) - np.float32(x0))
        kls[i]["numOfPixels"] = max(abs(int(round(float(np.float32(x1)))) - int(round(float(np.float32(x0))))),
                                    abs(int(round(float(np.float32(y1)))) - int(round(float(np.float32(y0)))))) + 1
    want, _ = oracle.lbd(img, kls.view(oracle.KEYLINE_DTYPE))
    got = pl.LineSegment().compute_lbd(img, kls)
    assert np.array_equal(got, want)
