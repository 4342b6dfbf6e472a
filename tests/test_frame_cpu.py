"""CPU pins of the Frame post-extraction oracle (oracle/frame_oracle.cc): undistortion against cv2 4.13's
undistortPoints (bit for bit, the reference's TUM1 calibration), RGB-D stereo coordinates and grid assignment
against an independent numpy restatement of the disassembled semantics, image bounds of the product's
host helper against the oracle."""
import numpy as np

from matchdata import GRID_COLS, GRID_ROWS


def _cv2_undistort(cal, xy):
    import cv2
    K = np.array([[cal["fx"], 0, cal["cx"]], [0, cal["fy"], cal["cy"]], [0, 0, 1]], np.float32)
    D = np.array([cal[k] for k in ("k1", "k2", "p1", "p2", "k3")], np.float32).reshape(5, 1)
    return cv2.undistortPoints(np.ascontiguousarray(xy, np.float32).reshape(-1, 1, 2), K, D, None, K).reshape(-1, 2)


def test_undistort_equals_cv2(oracle):
    import plslam_b200 as pl
    rng = np.random.default_rng(0)
    xy = np.stack([rng.uniform(-5, 645, 30000), rng.uniform(-5, 485, 30000)], 1).astype(np.float32)
    for cal in (pl.TUM1_CALIB, dict(pl.TUM1_CALIB, k3=0.0), dict(pl.TUM1_CALIB, k1=-0.28, k2=0.07, p1=1e-4, p2=-2e-4, k3=0.0)):
        assert np.array_equal(oracle.undistort_points(cal, xy), _cv2_undistort(cal, xy))
    # "not distorted" (mDistCoef(0) == 0): UndistortKeyPoints copies the keypoints
    cal0 = dict(pl.TUM1_CALIB, k1=0.0)
    assert np.array_equal(oracle.undistort_points(cal0, xy), xy)


def test_image_bounds(oracle):
    import plslam_b200 as pl
    for cal in (pl.TUM1_CALIB, dict(pl.TUM1_CALIB, k1=0.0)):
        b = oracle.image_bounds(cal, 640, 480)
        assert np.array_equal(b, pl.frame_image_bounds(cal, 640, 480))  # product host helper == oracle
        if cal["k1"] != 0:
            m = _cv2_undistort(cal, np.array([[0, 0], [640, 0], [0, 480], [640, 480]], np.float32))
            ref = np.array([min(m[0, 0], m[2, 0]), max(m[1, 0], m[3, 0]), min(m[0, 1], m[1, 1]), max(m[2, 1], m[3, 1])], np.float32)
            assert np.array_equal(b, ref)
        else:
            assert b.tolist() == [0.0, 640.0, 0.0, 480.0]


def test_stereo_and_grid_semantics(oracle):
    import plslam_b200 as pl
    from plslam_b200.synth import synth_depth
    cal = pl.TUM1_CALIB
    rng = np.random.default_rng(3)
    n = 1200
    xy = np.stack([rng.uniform(0, 639.9, n), rng.uniform(0, 479.9, n)], 1).astype(np.float32)
    depth = synth_depth(1).astype(np.float32) * np.float32(1.0 / 5000.0)
    depth[rng.random(depth.shape) < 0.2] = 0.0  # holes
    bounds = oracle.image_bounds(cal, 640, 480)
    o = oracle.frame_post(cal, bounds, xy, depth)
    un = _cv2_undistort(cal, xy)
    assert np.array_equal(o["un_xy"], un)
    d = depth[xy[:, 1].astype(np.int32), xy[:, 0].astype(np.int32)]  # truncating casts on the distorted keypoint
    assert np.array_equal(o["depth"], np.where(d > 0, d, np.float32(-1)))
    ur = np.where(d > 0, un[:, 0] - np.float32(cal["bf"]) / np.where(d > 0, d, np.float32(1)), np.float32(-1)).astype(np.float32)
    assert np.array_equal(o["uright"], ur)
    wInv = np.float32(GRID_COLS) / (bounds[1] - bounds[0])
    hInv = np.float32(GRID_ROWS) / (bounds[3] - bounds[2])
    fx = (un[:, 0] - bounds[0]) * wInv
    fy = (un[:, 1] - bounds[2]) * hInv
    px = (np.sign(fx) * np.floor(np.abs(fx) + np.float32(0.5))).astype(np.int64)  # roundf: half away from zero
    py = (np.sign(fy) * np.floor(np.abs(fy) + np.float32(0.5))).astype(np.int64)
    ok = (px >= 0) & (px < GRID_COLS) & (py >= 0) & (py < GRID_ROWS)
    cell = px * GRID_ROWS + py
    order = np.argsort(np.where(ok, cell, 1 << 30), kind="stable")[:ok.sum()]
    assert np.array_equal(o["grid_items"], order.astype(np.int32))
    counts = np.bincount(cell[ok], minlength=GRID_COLS * GRID_ROWS)
    assert np.array_equal(o["grid_start"][1:], np.cumsum(counts))
    assert 0 < (~ok).sum() < n // 4  # cells 64 / 48 are reached by rounding at the far edges: those keypoints are dropped
