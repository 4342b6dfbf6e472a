"""On-disk formats either side of the path: the TUM association list (LoadImages,
Examples/RGB-D/rgbd_tum.cc:151-176) and the trajectory line of System::SaveTrajectoryTUM (System.h:104, @0x3df90).
Host-only code: the oracle (iostream, as the reference writes it) is pinned against the reference's own association file, cv2's
gemm and scipy's quaternions; the product (byte-level parser, printf-style writer) must agree with the oracle byte for byte."""
import hashlib
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "tum_io.json")))
REF_ASSOC = "/root/reference/Examples/RGB-D/associations"


def assoc_digest(ts, rgb, dep):
    h = hashlib.sha256()
    for t, a, b in zip(ts, rgb, dep):
        h.update(("%s|%s|%s\n" % (float(t).hex(), a, b)).encode())
    return h.hexdigest()


TRICKY = ("1305031453.359684 rgb/a.png 1305031453.374112 depth/a.png\n"
          "\n"
          "1305031453.391690\trgb/b.png   1305031453.404816 depth/b.png  trailing words\n"
          "   \n"                                            # white space only: an entry of zeros / empty names
          "1305031453.423683 rgb/c.png\n"                     # short line
          "12abc rgb/d.png 3 depth/d.png\r\n"                 # number glued to text, CRLF
          "not-a-number rgb/e.png 4 depth/e.png\n"
          "-1.5e3 rgb/f.png 5 depth/f.png")                   # no newline at the end


def test_loader_equals_oracle_on_tricky_lines(oracle, tmp_path):
    import plslam_b200 as pl
    p = tmp_path / "assoc.txt"
    p.write_bytes(TRICKY.encode())
    ots, orgb, odep = oracle.tum_load_associations(str(p))
    rgb, dep, ts = pl.LoadImages(str(p))
    assert len(ots) == 7 == len(ts)
    assert np.array_equal(ts, ots) and rgb == orgb and dep == odep
    assert ts[0] == 1305031453.359684 and rgb[1] == "rgb/b.png" and dep[1] == "depth/b.png"
    assert (ts[2], rgb[2], dep[2]) == (0.0, "", "")
    assert (ts[3], rgb[3], dep[3]) == (1305031453.423683, "rgb/c.png", "")
    assert (ts[4], rgb[4]) == (12.0, "abc")
    assert (ts[5], rgb[5], dep[5]) == (0.0, "", "")
    assert (ts[6], dep[6]) == (-1500.0, "depth/f.png")
    empty = tmp_path / "empty.txt"
    empty.write_bytes(b"")
    assert len(pl.LoadImages(str(empty))[2]) == 0 == len(oracle.tum_load_associations(str(empty))[0])
    with pytest.raises(pl.PlslamError):
        pl.LoadImages(str(tmp_path / "missing.txt"))


def test_loader_on_a_full_size_list(oracle, tmp_path):
    """573 entries in fr1_desk's layout (the real list is not shipped: see the golden test below)."""
    import plslam_b200 as pl
    rng = np.random.default_rng(5)
    t = 1305031453.359684 + np.cumsum(rng.uniform(0.02, 0.04, 573))
    lines = ["%.6f rgb/%.6f.png %.6f depth/%.6f.png" % (a, a, a + 0.0144, a + 0.0144) for a in t]
    p = tmp_path / "fr1_like.txt"
    p.write_text("\n".join(lines) + "\n")
    rgb, dep, ts = pl.LoadImages(str(p))
    ots, orgb, odep = oracle.tum_load_associations(str(p))
    assert len(ts) == 573 and np.array_equal(ts, ots) and rgb == orgb and dep == odep
    assert np.array_equal(ts, np.array([float("%.6f" % a) for a in t]))


@pytest.mark.skipif(not os.path.isdir(REF_ASSOC), reason="the reference tree is only present in the build container")
@pytest.mark.parametrize("name", sorted(GOLD["associations"]))
def test_reference_association_files_against_golden(oracle, name):
    """Every association list the reference ships: oracle and product agree with the committed digests
    (tests/golden/make_golden.py wrote them from the oracle's parse, cross-checked there against str.split)."""
    import plslam_b200 as pl
    g = GOLD["associations"][name]
    path = os.path.join(REF_ASSOC, name)
    ots, orgb, odep = oracle.tum_load_associations(path)
    assert (len(ots), assoc_digest(ots, orgb, odep)) == (g["n"], g["sha256"])
    rgb, dep, ts = pl.LoadImages(path)
    assert (len(ts), assoc_digest(ts, rgb, dep)) == (g["n"], g["sha256"])
    assert [float(ts[0]).hex(), rgb[0], dep[0]] == g["first"] and [float(ts[-1]).hex(), rgb[-1], dep[-1]] == g["last"]


def random_poses(n, seed):
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(seed)
    R = Rotation.random(n, random_state=seed).as_matrix()
    # rotations by ~180 degrees about each axis exercise the three trace <= 0 branches
    for k, axis in enumerate(np.eye(3)):
        R[k] = Rotation.from_rotvec(axis * (np.pi - 1e-3 * (k + 1))).as_matrix()
    T = np.zeros((n, 4, 4), np.float32)
    T[:, :3, :3] = R
    T[:, :3, 3] = rng.uniform(-3, 3, (n, 3))
    T[:, 3, 3] = 1
    return T


def test_pose_arithmetic_pinned_against_cv2_and_scipy(oracle):
    cv2 = pytest.importorskip("cv2")
    from scipy.spatial.transform import Rotation
    T = random_poses(200, 3)
    for Tcw in T:
        v = oracle.tum_pose(Tcw)
        Rwc = np.ascontiguousarray(Tcw[:3, :3].T)
        twc = cv2.gemm(Rwc, np.ascontiguousarray(Tcw[:3, 3:4]), -1.0, None, 0.0).ravel()  # -Rwc * tcw as cv::Mat evaluates it
        assert np.array_equal(v[:3], twc)
        q = Rotation.from_matrix(Rwc.astype(np.float64)).as_quat()
        if np.dot(q, v[3:].astype(np.float64)) < 0:
            q = -q
        assert np.abs(q - v[3:]).max() < 2e-6  # scipy re-orthonormalises the float matrix; Eigen does not


def test_trajectory_lines_equal_oracle(oracle, tmp_path):
    import plslam_b200 as pl
    T = random_poses(300, 4)
    ts = 1305031453.359684 + 0.033 * np.arange(len(T))
    want = [oracle.tum_pose_line(t, Tcw) for t, Tcw in zip(ts, T)]
    assert [pl.trajectory_line(t, Tcw) for t, Tcw in zip(ts, T)] == want
    assert want[0].count(" ") == 7 and want[0].endswith("\n") and want[0].startswith("1305031453.359684 ")
    assert all(len(x.split()[1].split(".")[1]) == 9 for x in want[:5])
    out = tmp_path / "CameraTrajectory.txt"
    pl.SaveTrajectoryTUM(str(out), ts, T)
    assert out.read_text() == "".join(want)
    assert pl.trajectory_line(0.0, np.eye(4, dtype=np.float32)) == GOLD["identity_line"] == oracle.tum_pose_line(0.0, np.eye(4))
