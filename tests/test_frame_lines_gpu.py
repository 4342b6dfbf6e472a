"""Line analogues of the Frame steps (reference include/Frame.h:267 UndistortKeyLines, :116 GetLinesInArea, :107
isInFrustum(MapLine*, float)): CUDA path against the oracle.  PARITY UNPINNED - the reference ships no definition of these in any
form; the oracle restates the header contracts and the public fork family (oracle/frame_oracle.cc), so these tests establish
oracle <-> CUDA consistency only."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _keylines(seed, n=400, W=640, H=480):
    rng = np.random.default_rng(seed)
    s = np.stack([rng.uniform(0, W, n), rng.uniform(0, H, n)], 1)
    e = s + rng.normal(0, 60, (n, 2))
    e[:, 0] = np.clip(e[:, 0], 0, W - 1); e[:, 1] = np.clip(e[:, 1], 0, H - 1)
    return np.hstack([s, e]).astype(np.float32)


def test_undistort_keylines_matches_oracle(oracle):
    import plslam_b200 as pl
    xy4 = _keylines(1)
    for calib in (pl.TUM1_CALIB, dict(pl.TUM1_CALIB, k1=0.0)):
        c10 = np.array([calib[k] for k in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2", "k3")] + [40.0], np.float32)
        want = oracle.undistort_keylines(c10, xy4)
        got = pl.undistort_keylines_host(calib, xy4)
        assert np.array_equal(got, want)
    assert np.array_equal(got, xy4)                       # k1 == 0 copies
    assert pl.undistort_keylines_host(pl.TUM1_CALIB, np.zeros((0, 4), np.float32)).shape == (0, 4)


def test_lines_in_area_matches_oracle(oracle):
    import plslam_b200 as pl
    rng = np.random.default_rng(2)
    xy4 = _keylines(3, n=333)
    mid = 0.5 * (xy4[:, :2] + xy4[:, 2:])
    ang = np.arctan2(xy4[:, 3] - xy4[:, 1], xy4[:, 2] - xy4[:, 0])
    lines4 = np.stack([mid[:, 0], mid[:, 1], ang, rng.integers(0, 3, len(xy4))], 1).astype(np.float32)
    q = []
    for _ in range(200):
        i = int(rng.integers(0, len(xy4)))
        j = rng.normal(0, 4, 4)
        q.append([xy4[i, 0] + j[0], xy4[i, 1] + j[1], xy4[i, 2] + j[2], xy4[i, 3] + j[3], float(rng.uniform(5, 120)),
                  float(rng.integers(-1, 3)), float(rng.integers(-1, 3))])
    q = np.array(q, np.float32)
    got = pl.lines_in_area_host(q, lines4)
    total = 0
    for k in range(len(q)):
        want = oracle.get_lines_in_area(q[k], lines4)
        assert np.array_equal(got[k], want), k
        total += len(want)
    assert total > 500
    assert all(len(x) == 0 for x in pl.lines_in_area_host(q[:3], np.zeros((0, 4), np.float32)))


def test_line_in_frustum_matches_oracle(oracle):
    import plslam_b200 as pl
    from matchdata import frustum_case
    seen = 0
    for seed, lim in ((1, 0.5), (2, 0.8)):
        c = frustum_case(2000, seed=seed, motion=0.1)
        rng = np.random.default_rng(seed)
        sp = c["xyz"]
        ep = (sp + rng.normal(0, 0.15, sp.shape)).astype(np.float32)
        sp_ep = np.hstack([sp, ep]).astype(np.float32)
        want = oracle.line_in_frustum(sp_ep, c["normal"], c["dist_range"], c["cam8"], c["tcw"], c["ow"], c["mbf"], c["log_sf"], c["n_levels"], lim)
        got = pl.line_in_frustum_host(sp_ep, c["normal"], c["dist_range"], c["cam8"], c["tcw"], c["ow"], c["mbf"], c["log_sf"], c["n_levels"], lim)
        for name in ("in_view", "proj", "level", "viewcos"):
            assert np.array_equal(got[name], want[name]), (seed, name)
        seen += int(want["in_view"].sum())
    assert seen > 300
