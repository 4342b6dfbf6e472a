"""CPU tests pinning the oracle (oracle/*.cc) against the only executable copy of the OpenCV
primitives in this environment, cv2 4.13, and against the golden data the
reference binary holds (constructor tables, rBRIEF pattern).  No GPU needed."""
import hashlib
import math

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


@pytest.fixture(scope="module")
def frames():
    from plslam_b200.synth import synth_frame
    return [synth_frame(s) for s in range(3)]


def test_constructor_tables_match_reference_binary(oracle):
    # quotas and umax recovered from lib/libORB_SLAM2.so@0x73050
    t = oracle.OrbOracle(1000, 1.2, 8, 20, 7).tables()
    assert t["quota"].tolist() == [217, 181, 151, 126, 105, 87, 73, 60]
    assert t["umax"].tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert oracle.OrbOracle(2000, 1.2, 8, 20, 7).tables()["quota"].tolist() == [434, 362, 302, 251, 209, 175, 145, 122]
    assert oracle.OrbOracle(8000, 1.2, 8, 20, 7).tables()["quota"].tolist() == [1737, 1448, 1207, 1005, 838, 698, 582, 485]
    assert np.allclose(t["scale"], [1, 1.2, 1.44, 1.728, 2.0736, 2.48832, 2.985984, 3.5831816], rtol=1e-6)


def test_pattern_table_hash(oracle):
    # bit_pattern_31_ from lib/libORB_SLAM2.so@0x341c40 (sha256 in SURVEY section 0)
    p = oracle.orb_pattern().astype("<i4")
    assert hashlib.sha256(p.tobytes()).hexdigest() == "7e645581387b82784797e8adddb9b6f0c12611859fda09ca8a9bec96d767a05f"
    assert p[:8].tolist() == [8, -3, 9, 5, 4, 2, 7, -12]


def test_resize_chain_bit_exact(oracle, frames):
    inv = oracle.OrbOracle().tables()["inv_scale"]
    for img in frames[:2] + [np.random.default_rng(1).integers(0, 256, (720, 1280)).astype(np.uint8)]:
        H, W = img.shape
        cur = img
        for l in range(1, 8):
            w = int(np.rint(np.float32(W) * inv[l])); h = int(np.rint(np.float32(H) * inv[l]))
            ref = cv2.resize(cur, (w, h), interpolation=cv2.INTER_LINEAR)
            assert np.array_equal(ref, oracle.resize_linear(cur, w, h))
            cur = ref


def test_blur_bit_exact(oracle, frames):
    noise = np.random.default_rng(2).integers(0, 256, (333, 517)).astype(np.uint8)
    for img in frames + [noise]:
        ref = cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
        assert np.array_equal(ref, oracle.blur7(img))


def test_fast_bit_exact(oracle, frames):
    img = frames[0]
    for th in (20, 7):
        det = cv2.FastFeatureDetector_create(th, True)
        for sub in (img, np.ascontiguousarray(img[16:54, 100:137]), np.ascontiguousarray(img[200:238, 300:337])):
            ref = np.array([[int(p.pt[0]), int(p.pt[1]), int(p.response)] for p in det.detect(sub)], np.int32).reshape(-1, 3)
            assert np.array_equal(ref, oracle.fast9(sub, th))


def test_fast_atan2_bit_exact(oracle):
    rng = np.random.default_rng(3)
    ys = rng.integers(-60000, 60000, 5000); xs = rng.integers(-60000, 60000, 5000)
    ref = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in zip(ys, xs)], np.float32)
    mine = np.array([oracle.fast_atan2(y, x) for y, x in zip(ys, xs)], np.float32)
    assert np.array_equal(ref, mine)
    assert oracle.fast_atan2(0, 0) == 0.0


def test_pinned_trig_is_correctly_rounded_almost_everywhere(oracle):
    xs = np.linspace(0, 6.2831855, 50001).astype(np.float32)
    sc = np.array([oracle.sincos(x) for x in xs], np.float32)
    assert np.array_equal(sc[:, 0], np.sin(xs.astype(np.float64)).astype(np.float32))
    assert np.array_equal(sc[:, 1], np.cos(xs.astype(np.float64)).astype(np.float32))
    rng = np.random.default_rng(4)
    y = rng.normal(0, 100, 20000).astype(np.float32); x = rng.normal(0, 100, 20000).astype(np.float32)
    mine = np.array([oracle.pl_atan2f(a, b) for a, b in zip(y, x)], np.float32)
    ref = np.arctan2(y.astype(np.float64), x.astype(np.float64)).astype(np.float32)
    assert np.mean(mine != ref) < 1e-3 and np.max(np.abs(mine - ref)) < 5e-7


def test_lsd_front_half_bit_exact(oracle):
    for sigma, ks in ((0.75, 7), (2.0, 7), (1.1, 9)):
        k = oracle.gauss_table_u8(sigma, ks)
        assert k.sum() == 256
        img = np.random.default_rng(5).integers(0, 256, (97, 131)).astype(np.uint8)
        assert np.array_equal(cv2.GaussianBlur(img, (ks, ks), sigma), oracle.gauss_blur_u8(img, k))
    assert oracle.gauss_table_u8(0.75, 7).tolist() == [0, 4, 56, 136, 56, 4, 0]
    for (W, H) in ((640, 480), (333, 250), (641, 481), (100, 77)):
        g = np.random.default_rng(6).integers(0, 256, (H, W)).astype(np.uint8)
        ref = cv2.resize(g, None, fx=0.8, fy=0.8, interpolation=cv2.INTER_LINEAR_EXACT)
        assert np.array_equal(ref, oracle.resize_linear_exact(g, ref.shape[1], ref.shape[0], 0.8))


@pytest.mark.parametrize("size", [(640, 480), (1280, 720), (333, 250)])
def test_lsd_identical_to_cv2(oracle, size):
    """COMPAT mode (libm trig) reproduces cv2 4.13 LSD_REFINE_ADV segments, widths, precisions and NFA."""
    from plslam_b200.synth import synth_frame
    W, H = size
    det = cv2.createLineSegmentDetector(cv2.LSD_REFINE_ADV)
    for seed in range(2):
        img = synth_frame(seed, W, H)
        lines, width, prec, nfa = det.detect(img)
        mine, _ = oracle.lsd_detect(img, compat=1)
        assert len(mine) == len(lines) and len(mine) > 20
        assert np.array_equal(lines.reshape(-1, 4), mine[:, :4].astype(np.float32))
        assert np.array_equal(width.ravel(), mine[:, 4])
        assert np.array_equal(prec.ravel(), mine[:, 5])
        assert np.allclose(nfa.ravel(), mine[:, 6], rtol=0, atol=1e-9)


def test_lsd_pinned_mode_agrees_with_cv2_on_nearly_all_lines(oracle, frames):
    """PINNED mode (shared sin/cos definition, what the CUDA path reproduces) vs cv2: measured agreement."""
    det = cv2.createLineSegmentDetector(cv2.LSD_REFINE_ADV)
    tot = same = 0
    for img in frames:
        ref = det.detect(img)[0].reshape(-1, 4)
        mine, _ = oracle.lsd_detect(img, compat=0)
        s = {tuple(np.float32(r[:4])) for r in mine}
        tot += len(ref); same += sum(tuple(r) in s for r in ref)
    assert same / tot > 0.99


def test_extract_lines_shapes_and_selection(oracle, frames):
    kl, desc, funcs, nd = oracle.extract_lines(frames[0], max_lines=40)
    assert nd > 40 and len(kl) == 40 and desc.shape == (40, 32) and funcs.shape == (40, 3)
    assert np.all(np.diff(kl["response"]) <= 0) and kl["class_id"].tolist() == list(range(40))
    allk, _, _, _ = oracle.extract_lines(frames[0], max_lines=0)
    assert len(allk) == nd
    # the kept lines are the 40 largest responses, ties in detection order
    order = np.argsort(-allk["response"], kind="stable")[:40]
    assert np.array_equal(allk["startPointX"][order], kl["startPointX"])
    # line functions: unit normal, passes through both end points
    for i in range(40):
        l = funcs[i]
        assert abs(math.hypot(l[0], l[1]) - 1) < 1e-12
        assert abs(l @ [kl["startPointX"][i], kl["startPointY"][i], 1.0]) < 1e-9
