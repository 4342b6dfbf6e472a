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


def test_matchers_through_the_reference_signatures(oracle, tmp_path):
    """ORB_SLAM2::ORBmatcher called with the reference's own signatures (include/ORBmatcher.h:61, :78, :104, :111) on mock
    Frame / KeyFrame / MapPoint classes carrying the reference's member names (host/dropin_matchers_demo.cc): what the
    calls leave in the objects (mvpMapPoints, vpMapPointMatches, vMatchedPairs, return values) must equal the flattened
    C-ABI calls on the same arrays — which tests/test_golden_gpu.py pins against the reference's machine code."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import plslam_b200 as pl
    from matchdata import fake_feature_vector, local_points_case, projection_case, triangulation_case
    from plslam_b200.synth import synth_frame, synth_pair
    demo = os.path.join(ROOT, "rgbd-pl-slam_b200", "host", "dropin_matchers_demo")
    assert os.path.exists(demo), "build the veneer: make -C rgbd-pl-slam_b200/host"
    d = tmp_path

    def put(name, a, dtype=None):
        np.ascontiguousarray(a, dtype).tofile(d / name)

    orc = oracle.OrbOracle()
    sf = orc.tables()["scale"]
    a, b = synth_pair(21)
    (ka, da), (kb, db) = orc.extract(a), orc.extract(b)
    # TrackWithMotionModel
    last, cur, cam, _, tc, tl = projection_case(ka, da, kb, db, sf, seed=21, motion=0.02)
    for k, v in last.items():
        put("pj_last_" + k, v)
    for k, v in cur.items():
        put("pj_cur_" + k, v)
    put("pj_cam", cam, np.float32); put("pj_sf", sf, np.float32); put("pj_tc", tc, np.float32); put("pj_tl", tl, np.float32)
    put("pj_par", [7.0, 0.0], np.float32)
    # SearchLocalPoints
    mp, fr, cam4 = local_points_case(ka, da, kb, db, seed=5, jitter=3.0)
    for k, v in mp.items():
        put("lp_mp_" + k, v)
    for k, v in fr.items():
        put("lp_fr_" + k, v)
    put("lp_cam4", cam4, np.float32); put("lp_sf", sf, np.float32); put("lp_par", [3.0, 0.8], np.float32)
    # TrackReferenceKeyFrame
    rng = np.random.default_rng(3)
    kf = dict(desc=da, angle=np.ascontiguousarray(ka["angle"]), valid=(rng.random(len(da)) < 0.85).astype(np.uint8))
    kf["nodes"], kf["start"], kf["idx"] = fake_feature_vector(da, seed=7)
    f = dict(desc=db, angle=np.ascontiguousarray(kb["angle"]))
    f["nodes"], f["start"], f["idx"] = fake_feature_vector(db, seed=7)
    for k, v in kf.items():
        put("bw_kf_" + k, v)
    for k, v in f.items():
        put("bw_f_" + k, v)
    put("bw_par", [0.7, 1.0], np.float32)
    # LoopClosing::ComputeSim3
    kfb = dict(f, valid=(rng.random(len(db)) < 0.85).astype(np.uint8))
    for k, v in kf.items():
        put("bk_kf1_" + k, v)
    for k, v in kfb.items():
        put("bk_kf2_" + k, v)
    put("bk_par", [0.75, 1.0], np.float32)
    # Relocalization
    from matchdata import relocalisation_case
    rkf, rcur, rcam, rsf, rlsf, rtc = relocalisation_case(ka, da, kb, db, sf, seed=21, motion=0.03)
    for k, v in rkf.items():
        put("rk_kf_" + k, v)
    for k, v in rcur.items():
        put("rk_cur_" + k, v)
    put("rk_cam", rcam, np.float32); put("rk_sf", rsf, np.float32); put("rk_tc", rtc, np.float32)
    put("rk_par", [10.0, 100.0, float(rlsf)], np.float32)
    # LoopClosing::ComputeSim3 (projection with a similarity)
    from matchdata import loop_projection_case
    lkf, lmp, lscw, lmi = loop_projection_case(ka, da, kb, db, sf, seed=21, motion=0.02, scale=1.3)
    for k, v in lkf.items():
        if k not in ("gwi", "ghi", "log_sf"):
            put("lc_kf_" + k, v)
    for k, v in lmp.items():
        put("lc_mp_" + k, v)
    put("lc_scw", lscw, np.float32); put("lc_matched_in", lmi, np.int32)
    put("lc_par", [10.0, float(lkf["gwi"]), float(lkf["ghi"]), float(lkf["log_sf"])], np.float32)
    # LocalMapping::SearchInNeighbors / LoopClosing::SearchAndFuse
    from matchdata import fuse_case
    fcases = {"fu_": fuse_case(ka, da, kb, db, sf, seed=21, motion=0.02), "fs_": fuse_case(ka, da, kb, db, sf, seed=22, motion=0.02, scale=1.4, sim3=True)}
    for pfx, (fkf, fmp, fkp) in fcases.items():
        for k, v in fkf.items():
            if k not in ("gwi", "ghi", "log_sf", "mbf"):
                put(pfx + "kf_" + k, v)
        for k, v in fmp.items():
            put(pfx + "mp_" + k, v)
        for k in ("has", "nobs", "bad"):
            put(pfx + "kp_" + k, fkp[k])
        put(pfx + "par", [3.0 if pfx == "fu_" else 4.0, float(fkf["gwi"]), float(fkf["ghi"]), float(fkf["log_sf"]), float(fkf["mbf"])], np.float32)
    # LoopClosing::ComputeSim3 (SearchBySim3)
    from matchdata import sim3_case
    skf1, skf2, smp1, smp2, ss12, sR12, st12, smi = sim3_case(ka, da, kb, db, sf, seed=21, s12=1.25)
    for pfx, dd in (("s3_kf1_", skf1), ("s3_kf2_", skf2), ("s3_mp1_", smp1), ("s3_mp2_", smp2)):
        for k, v in dd.items():
            if k not in ("gwi", "ghi", "log_sf"):
                put(pfx + k, v)
    put("s3_R12", sR12, np.float32); put("s3_t12", st12, np.float32); put("s3_matched_in", smi, np.int32)
    put("s3_par", [7.5, float(skf1["gwi"]), float(skf1["ghi"]), float(skf1["log_sf"]), float(ss12)], np.float32)
    # CreateNewMapPoints
    kps, desc = orc.extract(synth_frame(33))
    kf1, kf2, F12, pose, camt, sft, sg = triangulation_case(kps, desc, seed=33, stereo_fraction=0.5)
    for k, v in kf1.items():
        put("tr_kf1_" + k, v)
    for k, v in kf2.items():
        put("tr_kf2_" + k, v)
    put("tr_F12", F12, np.float32); put("tr_R2w", pose[0], np.float32); put("tr_t2w", pose[1], np.float32); put("tr_Cw", pose[2], np.float32)
    put("tr_sf", sft, np.float32); put("tr_sg", sg, np.float32); put("tr_par", list(camt) + [0.0, 1.0], np.float32)

    r = subprocess.run([demo, str(d)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr

    # expectations: the flattened calls on the same arrays
    m, n = pl.search_by_projection_host(last, cur, cam, sf, tc, tl, 7.0, False, True, report_removed=True)
    out = np.fromfile(d / "pj_out", np.int32)
    had = (np.arange(len(m)) % 7 == 0) & (cur["taken"] == 0)  # the demo's unobserved map points
    want = np.where(m >= 0, m, np.where((m == -2) & had, -2, -1))
    assert out[-1] == n > 50 and np.array_equal(out[:-1], want)
    m, n = pl.search_local_points_host(mp, fr, cam4, sf, 3.0, 0.8)
    out = np.fromfile(d / "lp_out", np.int32)
    assert out[-1] == n > 100 and np.array_equal(out[:-1], m)
    em, en = oracle.search_by_bow(kf, f, 0.7, True)  # (the kernel equals the oracle: tests/test_match_gpu.py)
    out = np.fromfile(d / "bw_out", np.int32)
    assert out[-1] == en > 50 and np.array_equal(out[:-1], em)
    em, en = oracle.search_by_bow_kfkf(kf, kfb, 0.75, True)  # (kernel vs reference code: tests/test_golden_gpu.py, bk*)
    out = np.fromfile(d / "bk_out", np.int32)
    assert out[-1] == en > 50 and np.array_equal(out[:-1], em)
    m, n = pl.search_by_projection_kf_host(rkf, rcur, rcam, rsf, rlsf, rtc, 10.0, 100, True)  # (pinned: test_golden_gpu.py, rk*)
    out = np.fromfile(d / "rk_out", np.int32)
    assert out[-1] == n > 100 and np.array_equal(out[:-1], m)
    m, n = pl.search_by_projection_sim3_host(lkf, lmp, lscw, lmi, 10)  # (pinned: test_golden_gpu.py, lc*)
    out = np.fromfile(d / "lc_out", np.int32)
    assert out[-1] == n > 100 and np.array_equal(out[:-1], m)
    fkf, fmp, fkp = fcases["fu_"]   # (pinned: test_golden_gpu.py, fu* / fs*)
    nf, log = oracle.fuse_replay(pl.fuse_search_host(fkf, fmp, 3.0), fmp, fkp)
    assert int(np.fromfile(d / "fu_out", np.int32)[-1]) == nf > 50
    assert np.array_equal(np.fromfile(d / "fu_log", np.int32).reshape(-1, 3), np.array(log, np.int32).reshape(-1, 3))
    fkf, fmp, fkp = fcases["fs_"]
    nf, log, rep = oracle.fuse_replay_sim3(pl.fuse_search_host(fkf, fmp, 4.0, scw=fkf["scw"]), fmp, fkp)
    out = np.fromfile(d / "fs_out", np.int32)
    assert out[-1] == nf > 50 and np.array_equal(out[:-1], rep)
    assert np.array_equal(np.fromfile(d / "fs_log", np.int32).reshape(-1, 3), np.array(log, np.int32).reshape(-1, 3))
    m, n = pl.search_by_sim3_host(skf1, skf2, smp1, smp2, ss12, sR12, st12, 7.5, smi)   # (pinned: test_golden_gpu.py, s3*)
    out = np.fromfile(d / "s3_out", np.int32)
    assert out[-1] == n > 50 and np.array_equal(out[:-1], m)
    ex, ey = pl.epipole(*pose, *camt)
    m12, n12, pairs = pl.search_for_triangulation_host(kf1, kf2, F12, ex, ey, sft, sg, False, True)
    out = np.fromfile(d / "tr_out", np.int32)
    assert out[-1] == n12 > 50
    assert np.array_equal(out[:-1].reshape(-1, 2), np.asarray(pairs, np.int32).reshape(-1, 2))
