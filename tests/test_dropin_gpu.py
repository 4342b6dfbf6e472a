"""The C++ drop-in classes (rgbd-pl-slam_b200/host: ORB_SLAM2::ORBextractor, LineSegment, LSDmatcher, ORBmatcher)
called the way Frame::ExtractORB / Frame::ExtractLSD call the reference's, compared with the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "rgbd-pl-slam_b200", "host", "dropin_demo")


def test_cpp_classes_match_oracle(oracle, tmp_path):
    from plslam_b200.synth import synth_pair
    assert os.path.exists(DEMO), "build the veneer: make -C rgbd-pl-slam_b200/host"
    a, b = synth_pair(11)
    fa, fb, out = tmp_path / "a.raw", tmp_path / "b.raw", tmp_path / "out.bin"
    a.tofile(fa); b.tofile(fb)
    r = subprocess.run([DEMO, "640", "480", str(fa), str(fb), str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    buf = open(out, "rb").read()
    hdr = struct.unpack("<8i", buf[:32])
    nA, nB, lA, lB, nl, d01, levels, nknn = hdr
    off = 32
    kps = np.frombuffer(buf, oracle.KP_DTYPE, nA, off); off += 28 * nA
    desc = np.frombuffer(buf, np.uint8, nA * 32, off).reshape(nA, 32); off += 32 * nA
    kls = np.frombuffer(buf, oracle.KEYLINE_DTYPE, lA, off); off += 68 * lA
    ldesc = np.frombuffer(buf, np.uint8, lA * 32, off).reshape(lA, 32); off += 32 * lA
    funcs = np.frombuffer(buf, np.float64, lA * 3, off).reshape(lA, 3); off += 24 * lA
    lmatch = np.frombuffer(buf, np.int32, lA, off); off += 4 * lA
    mads = np.frombuffer(buf, np.float64, 2, off)
    orc = oracle.OrbOracle()
    okA, odA = orc.extract(a)
    okB, odB = orc.extract(b)
    assert (nA, nB, levels) == (len(okA), len(okB), 8)
    for f in okA.dtype.names:
        assert np.array_equal(kps[f], okA[f]), f
    assert np.array_equal(desc, odA)
    lkA, ldA, lfA, _ = oracle.extract_lines(a, 40)
    lkB, ldB, lfB, _ = oracle.extract_lines(b, 40)
    assert (lA, lB) == (len(lkA), len(lkB))
    for f in lkA.dtype.names:
        assert np.array_equal(kls[f], lkA[f]), f
    assert np.array_equal(ldesc, ldA) and np.array_equal(funcs, lfA)
    assert d01 == oracle.descriptor_distance(odA[0], odB[0])
    knn = oracle.knn2(ldA, ldB)
    exp = np.where((knn[:, 1] <= 100) & ((knn[:, 2] < 0) | (knn[:, 1].astype(np.float32) < np.float32(0.8) * knn[:, 3].astype(np.float32))),
                   knn[:, 0], -1)
    assert np.array_equal(lmatch, exp) and nl == int((exp >= 0).sum()) and nknn == lA

    def mad(x):  # include/auxiliar.h:92-106
        x = np.sort(np.asarray(x, np.float64)); m = x[len(x) // 2]
        return 1.4826 * np.sort(np.abs(x - m))[len(x) // 2]
    assert mads[0] == mad(knn[:, 1]) and mads[1] == mad(knn[:, 3] - knn[:, 1])
