#!/usr/bin/env python3
"""Generates the committed golden vectors under tests/golden/ (run in the BUILD container, where cv2 4.13, the
reference tree /root/reference and oracle/_ref exist; the GPU box only reads the .npz/.json files).

Sources of truth, none of them this repository's own oracle:
  cv2_primitives.npz   cv2 4.13: resize INTER_LINEAR, GaussianBlur 7x7 sigma 2, FAST-9/16 + NMS, fastAtan2,
                       undistortPoints (the OpenCV calls of ORBextractor / Frame: SURVEY.md Appendix B, section 8f-2)
  cv2_lsd.npz          cv2 4.13 createLineSegmentDetector(LSD_REFINE_ADV) on two small synthetic frames
  dbow2_ref.npz        the reference's OWN Thirdparty/DBoW2 sources (compiled into oracle/_ref/libdbow2_ref.so):
                       FORB::distance and ORBVocabulary::transform on a small vocabulary
  reference_binary.json  constants read from /root/reference/lib/libORB_SLAM2.so (rBRIEF pattern, matcher thresholds)
  reference_code.npz    outputs of the reference's OWN MACHINE CODE for four leaf functions of the matcher path
                       (RadiusByViewingCos, CheckDistEpipolarLine, ComputeThreeMaxima, DescriptorDistance), executed from
                       lib/libORB_SLAM2.so by tests/golden/reference_code.py  (`python tests/golden/make_golden.py refcode`)
  reference_library.npz  the reference library ITSELF, dlopen'ed over generated stub dependencies (reference_code.py:
                       RefLibrary): constructor tables of ORB_SLAM2::ORBextractor for four parameter sets, and the outputs of
                       ORBextractor::DistributeOctTree (run under a monotonic operator new, which fixes its pointer-valued
                       tie-break to allocation order) for real FAST candidate lists and for lattices full of ties
                       (`python tests/golden/make_golden.py reflib`, in a fresh process)
  tum_io.json          the reference's own association lists (Examples/RGB-D/associations/*.txt) parsed with str.split /
                       float(): entry count, digest, first and last entry of each; the trajectory line of the identity pose
                       from cv2.gemm + Python's "%.9f"  (`python tests/golden/make_golden.py tum` writes only this file)

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "rgbd-pl-slam_b200")):
    sys.path.insert(0, p)


def refcode():
    sys.path.insert(0, HERE)
    from reference_code import RefCode, SO
    r = RefCode()
    rng = np.random.default_rng(77)
    out = {"so_sha256": np.array(hashlib.sha256(open(SO, "rb").read()).hexdigest())}
    # RadiusByViewingCos: the floats around the threshold, and a spread
    t = np.float32(0.998)
    vals = [t]
    for _ in range(4):
        vals.append(np.nextafter(vals[-1], np.float32(2)))
    lo = t
    for _ in range(4):
        lo = np.nextafter(lo, np.float32(-2)); vals.append(lo)
    vals += list(rng.uniform(-1, 1, 64).astype(np.float32)) + [np.float32(1.0), np.float32(0.0)]
    out["radius_in"] = np.array(vals, np.float32)
    out["radius_out"] = np.array([r.radius_by_viewing_cos(float(v)) for v in vals], np.float32)
    # DescriptorDistance
    a = rng.integers(0, 256, (300, 32)).astype(np.uint8)
    b = rng.integers(0, 256, (300, 32)).astype(np.uint8)
    a[0] = 0; b[0] = 255; b[1] = a[1]; a[2] = 0; b[2] = 0; b[2, 31] = 0x80
    out["dd_a"], out["dd_b"] = a, b
    out["dd_out"] = np.array([r.descriptor_distance(x, y) for x, y in zip(a, b)], np.int32)
    # ComputeThreeMaxima: random populations, ties, and the 0.1 * max thresholds
    hists = [rng.integers(0, 40, 30) for _ in range(150)] + [rng.integers(0, 3, 30) for _ in range(50)]
    for m1 in (10, 20, 30, 50, 100, 101):
        for m2 in (m1 // 10 - 1, m1 // 10, m1 // 10 + 1, m1):
            for m3 in (0, m1 // 10 - 1, m1 // 10, m1 // 10 + 1):
                h = np.zeros(30, np.int64); pos = rng.permutation(30)[:3]
                h[pos[0]], h[pos[1]], h[pos[2]] = m1, max(m2, 0), max(m3, 0)
                hists.append(h)
    hists.append(np.zeros(30, np.int64))
    hists = np.array(hists, np.int32)
    out["tm_in"] = hists
    res = []
    for h in hists:
        i = r.compute_three_maxima([int(x) for x in h])
        res.append([-1 if x == -7 else x for x in i])  # untouched outputs keep the callers' initial -1
    out["tm_out"] = np.array(res, np.int32)
    # CheckDistEpipolarLine: for random geometry, bisect kp2.y in float32 until the reference's answer flips between two
    # ADJACENT floats, and keep both: any difference in the rounding sequence (the binary's FMA pattern) moves that boundary.
    sig = np.array([1.2 ** (2 * l) for l in range(8)], np.float32)
    cases = []
    while len(cases) < 1500:
        F = rng.standard_normal((3, 3)).astype(np.float32) * np.float32(10.0 ** rng.uniform(-4, 0))
        if len(cases) % 50 == 0:
            F[:, :2] = 0  # a = b = 0: den == 0 -> false
        kp1 = rng.uniform(0, 640, 2).astype(np.float32)
        x2 = np.float32(rng.uniform(0, 640)); octv = int(rng.integers(0, 8))
        f = lambda y: r.check_dist_epipolar_line(kp1, (x2, np.float32(y)), octv, F, sig)
        ys = np.linspace(-2000, 2000, 81).astype(np.float32)
        v = [f(y) for y in ys]
        flips = [i for i in range(80) if v[i] != v[i + 1]]
        if not flips:
            cases.append((F, kp1, x2, np.float32(rng.uniform(0, 480)), octv)); continue
        for i in flips[:2]:
            lo, hi = ys[i], ys[i + 1]
            flo = v[i]
            while np.nextafter(lo, hi) != hi:
                mid = np.float32((np.float64(lo) + np.float64(hi)) / 2)
                if mid == lo or mid == hi:
                    break
                if f(mid) == flo:
                    lo = mid
                else:
                    hi = mid
            cases.append((F, kp1, x2, lo, octv)); cases.append((F, kp1, x2, hi, octv))
    out["ep_F"] = np.array([c[0] for c in cases], np.float32)
    out["ep_kp1"] = np.array([c[1] for c in cases], np.float32)
    out["ep_kp2"] = np.array([[c[2], c[3]] for c in cases], np.float32)
    out["ep_oct"] = np.array([c[4] for c in cases], np.int32)
    out["ep_sigma2"] = sig
    out["ep_out"] = np.array([r.check_dist_epipolar_line(c[1], (c[2], c[3]), c[4], c[0], sig) for c in cases], np.uint8)
    np.savez_compressed(os.path.join(HERE, "reference_code.npz"), **out)
    print("reference_code.npz: %d radius, %d distances, %d histograms, %d epipolar cases (%d true)"
          % (len(vals), len(a), len(hists), len(cases), int(out["ep_out"].sum())))


def reflib():
    sys.path.insert(0, HERE)
    from reference_code import RefLibrary, SO
    # Every synthetic frame is rendered BEFORE the reference library is loaded: its stub / shim objects define cv:: symbols
    # with RTLD_GLOBAL, and plslam_b200.synth warps the second frame of a pair with cv2, whose lazily bound calls would then
    # land in the shims (observed: a different second frame on every call, eventually a crash).
    import plslam_b200.synth as _synth
    _frames = {("f", a): _synth.synth_frame(*a) for a in [(0,), (7,), (2,), (21,), (22,), (24,), (3, 320, 240), (11, 400, 304),
                                                            (12, 256, 200), (13, 333, 250), (0, 640, 480), (5, 640, 480),
                                                            (40, 1280, 720)]}
    _frames.update({("p", s): _synth.synth_pair(s) for s in (1, 2)})
    synth_frame = lambda *a: _frames[("f", a)]
    synth_pair = lambda s: _frames[("p", s)]
    R = RefLibrary()  # the monotonic operator new must precede any global libstdc++
    from oracle import bindings as orb  # only to produce realistic inputs (FAST candidates); outputs come from the reference
    out = {"so_sha256": np.array(hashlib.sha256(open(SO, "rb").read()).hexdigest())}
    params = [(1000, 1.2, 8, 20, 7), (2000, 1.2, 8, 20, 7), (8000, 1.2, 8, 20, 7), (500, 1.5, 4, 12, 5)]
    out["ctor_params"] = np.array(params, np.float64)
    objs = []
    for k, pr in enumerate(params):
        obj, t = R.extractor(*pr)
        objs.append(obj)
        for name in ("quota", "umax", "scale", "inv_scale", "sigma2", "inv_sigma2", "pattern"):
            out["ctor%d_%s" % (k, name)] = t[name]
    rng = np.random.default_rng(5)
    cases = []
    o = orb.OrbOracle()
    quota = o.tables()["quota"]
    for seed in (0, 7):
        o.extract(synth_frame(seed))
        for l in (0, 1, 3, 5, 7):
            h, w = o.level(l).shape
            cases.append((o.candidates(l), 16, w - 16, 16, h - 16, int(quota[l])))
    c0 = cases[0][0]
    cases.append((c0, 16, 640 - 16, 16, 480 - 16, 50))       # far fewer features than candidates
    cases.append((c0[:150], 16, 640 - 16, 16, 480 - 16, 217))  # more features than candidates
    for step, N in ((8, 200), (16, 150), (5, 400)):           # lattices with equal responses: every node size ties
        xs, ys = np.meshgrid(np.arange(0, 600, step), np.arange(0, 440, step))
        pts = np.stack([xs.ravel(), ys.ravel(), np.full(xs.size, 30)], 1).astype(np.int32)
        cases.append((pts[rng.permutation(len(pts))], 16, 640 - 16, 16, 480 - 16, N))
    pts = np.stack([rng.integers(0, 300, 3000), rng.integers(0, 440, 3000), rng.integers(7, 60, 3000)], 1).astype(np.int32)
    cases.append((pts, 16, 640 - 16, 16, 480 - 16, 300))      # everything in the left half
    out["qt_n"] = np.array(len(cases))
    for k, (c, x0, x1, y0, y1, N) in enumerate(cases):
        idx, _ = R.distribute(objs[0], c, x0, x1, y0, y1, N, 0)
        out["qt%d_in" % k] = np.ascontiguousarray(c, np.int32)
        out["qt%d_args" % k] = np.array([x0, x1, y0, y1, N], np.int32)
        out["qt%d_out" % k] = idx
    # ORBextractor::ComputeKeyPointsOctTree + computeOrientation, run from the reference library on the oracle's pyramid levels
    # of small frames (cv::FAST / cv::fastAtan2 supplied by the cv2-pinned shims of reference_code.py).  The fixture keeps the
    # input FRAME (the pyramid is rebuilt by the oracle at test time) and the reference's per-level keypoints.
    ck = []
    for k, (seed, W, H, nf) in enumerate([(3, 320, 240, 500), (11, 400, 304, 1000), (12, 256, 200, 300)]):
        img = synth_frame(seed, W, H)
        oo = orb.OrbOracle(nf)
        oo.extract(img)
        obj, t = R.extractor(nf)
        ref = R.compute_keypoints_oct_tree(obj, [oo.level(l) for l in range(8)])
        out["ck%d_img" % k] = img
        out["ck%d_nf" % k] = np.array(nf)
        out["ck%d_counts" % k] = np.array([len(r) for r in ref], np.int32)
        out["ck%d_kps" % k] = np.concatenate(ref)
        ck.append(int(sum(len(r) for r in ref)))
    out["ck_n"] = np.array(len(ck))
    # The whole ORBextractor::operator() (pyramid, key points, blur, rBRIEF, scaling), the reference's own code end to end.
    # Small frames are stored; the standard sizes are named by their synth_frame arguments (plslam_b200/synth.py).
    ex = []
    for k, (seed, W, H, nf, store) in enumerate([(3, 320, 240, 500, True), (13, 333, 250, 1000, True), (12, 256, 200, 300, True),
                                                 (0, 640, 480, 1000, False), (5, 640, 480, 1000, False), (40, 1280, 720, 2000, False)]):
        img = synth_frame(seed, W, H)
        obj, t = R.extractor(nf)
        kps, desc = R.extract(obj, img)
        out["ex%d_args" % k] = np.array([seed, W, H, nf], np.int32)
        if store:
            out["ex%d_img" % k] = img
        out["ex%d_img_sha256" % k] = np.array(hashlib.sha256(img.tobytes()).hexdigest())
        out["ex%d_kps" % k], out["ex%d_desc" % k] = kps, desc
        ex.append(len(kps))
    out["ex_n"] = np.array(len(ex))
    # ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th) — the local-map search — on faked Frame / MapPoint
    # objects; inputs are rebuilt at test time from the same seeds (tests/matchdata.py: local_points_case)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from matchdata import local_points_case
    oo = orb.OrbOracle()
    sf = oo.tables()["scale"]
    lp = []
    for seed in (1, 2):
        a, b = synth_pair(seed)
        (ka, da), (kb, db) = oo.extract(a), oo.extract(b)
        for th, nnr, jit in ((1.0, 0.8, 2.0), (3.0, 0.8, 6.0), (1.0, 0.6, 1.0), (5.0, 0.9, 10.0)):
            mpd, frd, cam4 = local_points_case(ka, da, kb, db, seed=10 * seed + int(th), jitter=jit)
            m, n = R.search_local_points(mpd, frd, cam4, sf, th, nnr)
            k = len(lp)
            out["lp%d_args" % k] = np.array([seed, th, nnr, jit], np.float64)
            out["lp%d_match" % k], out["lp%d_n" % k] = m, np.array(n)
            lp.append(n)
    out["lp_n"] = np.array(len(lp))
    print("SearchByProjection(Frame&, MapPoints, th):", lp)
    # ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) on faked Frame / MapPoint objects, cv::Mat expressions
    # (-Rcw.t()*tcw, Rlw*twc+tlw, Rcw*x3Dw+tcw) evaluated by the expression shims with cv::gemm's arithmetic
    from matchdata import projection_case
    pj = []
    for seed in (1, 2):
        a, b = synth_pair(seed)
        (ka, da), (kb, db) = oo.extract(a), oo.extract(b)
        for motion in (0.02, 0.3, -0.3):
            last, cur, cam, sf2, tc, tl = projection_case(ka, da, kb, db, sf, seed=seed, motion=motion)
            for th, mono in ((7.0, 0), (15.0, 0), (15.0, 1)):
                m, n = R.search_by_projection(last, cur, cam, sf, tc, tl, th, bool(mono), True)
                k = len(pj)
                out["pj%d_args" % k] = np.array([seed, motion, th, mono], np.float64)
                out["pj%d_match" % k], out["pj%d_n" % k] = m, np.array(n)
                pj.append(n)
    out["pj_n"] = np.array(len(pj))
    print("SearchByProjection(Frame&, Frame&):", pj)
    # ORBmatcher::SearchForInitialization (oracle only so far: the GPU kernel is a next-round item)
    from matchdata import frame_grid
    si = []
    for seed in (1, 2):
        a, b = synth_pair(seed)
        (ka, da), (kb, db) = oo.extract(a), oo.extract(b)
        mk = lambda k, d: dict(xy=np.stack([k["x"], k["y"]], 1).astype(np.float32), octave=k["octave"].astype(np.int32),
                               angle=k["angle"].astype(np.float32), desc=d)
        f1, f2 = mk(ka, da), mk(kb, db)
        gs, gi, (mnx, mxx, mny, mxy, gwi, ghi) = frame_grid(f2["xy"], 640, 480)
        f2["grid_start"], f2["grid_items"] = gs, gi
        cam4 = np.array([mnx, mny, gwi, ghi], np.float32)
        for win, nnr, ori in ((100, 0.9, 1), (30, 0.9, 1), (100, 0.6, 0)):
            m, n, prev = R.search_for_initialization(f1, f2, cam4, f1["xy"].copy(), win, nnr, bool(ori))
            k = len(si)
            out["si%d_args" % k] = np.array([seed, win, nnr, ori], np.float64)
            out["si%d_match" % k], out["si%d_n" % k], out["si%d_prev" % k] = m, np.array(n), prev
            si.append(n)
    out["si_n"] = np.array(len(si))
    print("SearchForInitialization:", si)
    # Frame::UndistortKeyPoints / ComputeImageBounds on a faked Frame (cv::undistortPoints = the cv2-pinned shim)
    r4 = np.random.default_rng(4)
    cal = [dict(fx=517.306408, fy=516.469215, cx=318.643040, cy=255.313989, k1=0.262383, k2=-0.953104, p1=-0.005358, p2=0.002628, k3=1.163314),
           dict(fx=535.4, fy=539.2, cx=320.1, cy=247.6, k1=0.0, k2=0.0, p1=0.0, p2=0.0, k3=0.0),
           dict(fx=520.9, fy=521.0, cx=325.1, cy=249.7, k1=0.2312, k2=-0.7849, p1=-0.0033, p2=-0.0001, k3=0.9172)]
    for k, c in enumerate(cal):
        xyu = np.stack([r4.uniform(0, 640, 1500), r4.uniform(0, 480, 1500)], 1).astype(np.float32)
        un, bnd = R.undistort_and_bounds(xyu, c, 640, 480)
        out["un%d_calib" % k] = np.array([c[n] for n in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2", "k3")], np.float64)
        out["un%d_xy" % k], out["un%d_out" % k] = xyu, np.stack([un["x"], un["y"]], 1)
        out["un%d_bounds" % k] = np.array(bnd, np.float32)
    out["un_n"] = np.array(len(cal))
    print("UndistortKeyPoints / ComputeImageBounds: %d calibrations" % len(cal))
    # Frame::ComputeStereoFromRGBD on a faked Frame: distorted / undistorted keypoints, a float depth map with holes and negatives
    r3 = np.random.default_rng(3)
    nst = 3000
    st_xy = np.stack([r3.uniform(0, 639.99, nst), r3.uniform(0, 479.99, nst)], 1).astype(np.float32)
    st_un = (st_xy + r3.normal(0, 2.0, st_xy.shape)).astype(np.float32)
    st_depth = (r3.uniform(0.3, 6, (480, 640)) * (r3.random((480, 640)) > 0.2)).astype(np.float32)
    st_depth[r3.random((480, 640)) < 0.02] = -1.0
    ur, dz = R.compute_stereo_from_rgbd(st_xy, st_un, st_depth, 40.0)
    out["st_xy"], out["st_un"], out["st_depth"] = st_xy, st_un, st_depth.astype(np.float16).astype(np.float32)
    ur, dz = R.compute_stereo_from_rgbd(st_xy, st_un, out["st_depth"], 40.0)  # on the stored (half-precision valued) map
    out["st_uright"], out["st_z"] = ur, dz
    out["st_depth"] = out["st_depth"].astype(np.float16)
    print("ComputeStereoFromRGBD: %d of %d keypoints with depth" % (int((dz > 0).sum()), nst))
    # ORBmatcher::SearchForTriangulation on faked KeyFrame objects: the reference computes its own epipole through
    # KeyFrame::GetCameraCenter / GetRotation / GetTranslation and the gemm shims, walks real std::map feature vectors
    from matchdata import triangulation_case
    tr = []
    tri_cases = [(21, {}), (22, dict(stereo_fraction=0.0, t21=(0.005, 0.002, 0.15))), (24, dict(nbits=1))]
    for seed, kw in tri_cases:
        kps, desc = orb.OrbOracle().extract(synth_frame(seed))
        kf1, kf2, F12, pose, cam, sft, sgt = triangulation_case(kps, desc, seed=seed, **kw)
        for only, ori in ((0, 1), (0, 0), (1, 1)):
            m, n, pairs = R.search_for_triangulation(kf1, kf2, F12, pose, cam, sft, sgt, bool(only), bool(ori))
            k = len(tr)
            out["tr%d_args" % k] = np.array([seed, only, ori, kw.get("stereo_fraction", 0.5), kw.get("nbits", 3)] + list(kw.get("t21", (0.12, 0.01, 0.03))), np.float64)
            out["tr%d_match" % k], out["tr%d_n" % k] = m, np.array(n)
            tr.append(n)
    out["tr_n"] = np.array(len(tr))
    print("SearchForTriangulation:", tr)
    # ORBmatcher::SearchByBoW(KeyFrame*, Frame&, matches) on faked KeyFrame / Frame objects (real std::map feature vectors)
    from matchdata import fake_feature_vector
    bw = []
    for seed in (1, 2):
        a, b = synth_pair(seed)
        (ka, da), (kb, db) = oo.extract(a), oo.extract(b)
        r2 = np.random.default_rng(seed)
        for nnr, ori, nbits in ((0.7, 1, 6), (0.9, 0, 6), (0.7, 1, 2), (0.6, 1, 4)):
            kf = dict(desc=da, angle=np.ascontiguousarray(ka["angle"]), valid=(r2.random(len(da)) < 0.85).astype(np.uint8))
            kf["nodes"], kf["start"], kf["idx"] = fake_feature_vector(da, nbits, seed=7)
            ff = dict(desc=db, angle=np.ascontiguousarray(kb["angle"]))
            ff["nodes"], ff["start"], ff["idx"] = fake_feature_vector(db, nbits, seed=7)
            m, n = R.search_by_bow(kf, ff, nnr, bool(ori))
            k = len(bw)
            out["bw%d_args" % k] = np.array([seed, nnr, ori, nbits], np.float64)
            out["bw%d_valid" % k] = kf["valid"]
            out["bw%d_match" % k], out["bw%d_n" % k] = m, np.array(n)
            bw.append(n)
    out["bw_n"] = np.array(len(bw))
    print("SearchByBoW(KeyFrame*, Frame&):", bw)
    # Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea on a faked Frame (the reference's own grid and window search)
    kps, _ = orb.OrbOracle().extract(synth_frame(2))
    kps = kps.copy()
    kps["x"] += rng.uniform(-45, 45, len(kps)).astype(np.float32)   # some keypoints fall outside the grid bounds
    kps["y"] += rng.uniform(-45, 45, len(kps)).astype(np.float32)
    bounds = (-11.5, 651.25, -9.75, 489.5)                           # undistorted-image bounds need not be integers
    q = [(rng.uniform(-40, 700), rng.uniform(-40, 520), rng.uniform(1, 90), int(rng.integers(-1, 8)), int(rng.integers(-1, 8)))
         for _ in range(600)]
    q += [(x, y, r, -1, -1) for x in (-11.5, 0, 10.35, 651.25, 700) for y in (-9.75, 0, 489.5) for r in (0.5, 10.4, 100)]
    start, items, res = R.frame_grid_queries(kps, bounds, q)
    out["fg_kps"], out["fg_bounds"], out["fg_queries"] = kps, np.array(bounds, np.float32), np.array(q, np.float64)
    out["fg_start"], out["fg_items"] = start, items
    out["fg_res_len"] = np.array([len(r) for r in res], np.int32)
    out["fg_res"] = np.concatenate(res).astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "reference_library.npz"), **out)
    print("reference_library.npz: %d constructor tables, %d quad-tree cases, outputs of sizes %s; ComputeKeyPointsOctTree on %d frames: %s keypoints; operator() on %d frames: %s keypoints; grid of %d keypoints (%d inside), %d window queries with %d candidates"
          % (len(params), len(cases), [len(out["qt%d_out" % k]) for k in range(len(cases))], len(ck), ck, len(ex), ex, len(kps),
             len(items), len(q), len(out["fg_res"])))


def reflib2():
    """Round-2 additions, kept in their own file (reference_library2.npz) so that reference_library.npz stays byte-identical:
    further ORBmatcher overloads executed from lib/libORB_SLAM2.so on faked objects."""
    sys.path.insert(0, HERE)
    from reference_code import RefLibrary, SO
    import plslam_b200.synth as _synth
    _pairs = {s: _synth.synth_pair(s) for s in (1, 2)}   # rendered before the library is loaded (see reflib)
    R = RefLibrary()
    from oracle import bindings as orb
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from matchdata import fake_feature_vector
    out = {"so_sha256": np.array(hashlib.sha256(open(SO, "rb").read()).hexdigest())}
    oo = orb.OrbOracle()
    feats = {s: (oo.extract(_pairs[s][0]), oo.extract(_pairs[s][1])) for s in (1, 2)}
    # ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) (loop closing) on faked KeyFrames
    bk = []
    for seed in (1, 2):
        (ka, da), (kb, db) = feats[seed]
        r2 = np.random.default_rng(100 + seed)
        for nnr, ori, nbits in ((0.75, 1, 6), (0.9, 0, 6), (0.75, 1, 2), (0.6, 1, 4)):
            kf1 = dict(desc=da, angle=np.ascontiguousarray(ka["angle"]), valid=(r2.random(len(da)) < 0.85).astype(np.uint8))
            kf1["nodes"], kf1["start"], kf1["idx"] = fake_feature_vector(da, nbits, seed=7)
            kf2 = dict(desc=db, angle=np.ascontiguousarray(kb["angle"]), valid=(r2.random(len(db)) < 0.85).astype(np.uint8))
            kf2["nodes"], kf2["start"], kf2["idx"] = fake_feature_vector(db, nbits, seed=7)
            m, n = R.search_by_bow_kfkf(kf1, kf2, nnr, bool(ori))
            k = len(bk)
            out["bk%d_args" % k] = np.array([seed, nnr, ori, nbits], np.float64)
            out["bk%d_valid1" % k], out["bk%d_valid2" % k] = kf1["valid"], kf2["valid"]
            out["bk%d_match" % k], out["bk%d_n" % k] = m, np.array(n)
            bk.append(n)
    out["bk_n"] = np.array(len(bk))
    print("SearchByBoW(KeyFrame*, KeyFrame*):", bk)
    # ORBmatcher::SearchByProjection(Frame&, KeyFrame*, set<MapPoint*>&, th, ORBdist) (relocalisation) on a faked KeyFrame /
    # Frame / MapPoints and a real std::set; MapPoint::PredictScale, isBad, GetWorldPos ... run from the library as well
    from matchdata import relocalisation_case
    sf = oo.tables()["scale"]
    rk = []
    for seed in (1, 2):
        (ka, da), (kb, db) = feats[seed]
        for motion in (0.03, 0.25, -0.25):
            kf, cur, cam, sf2, lsf, tcw = relocalisation_case(ka, da, kb, db, sf, seed=seed, motion=motion)
            for th, od, ori in ((10.0, 100, 1), (3.0, 64, 1), (10.0, 100, 0)):
                m, n = R.search_by_projection_kf(kf, cur, cam, sf, lsf, tcw, th, od, bool(ori))
                k = len(rk)
                out["rk%d_args" % k] = np.array([seed, motion, th, od, ori], np.float64)
                out["rk%d_match" % k], out["rk%d_n" % k] = m, np.array(n)
                rk.append(n)
    out["rk_n"] = np.array(len(rk))
    print("SearchByProjection(Frame&, KeyFrame*, set, th, ORBdist):", rk)
    # ORBmatcher::SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) (loop closing) on a faked KeyFrame (its grid a real
    # vector<vector<vector<size_t>>>), faked MapPoints and a similarity matrix with scale != 1
    from matchdata import loop_projection_case
    lc = []
    for seed in (1, 2):
        (ka, da), (kb, db) = feats[seed]
        for motion, scale, th in ((0.02, 1.0, 10), (0.02, 1.7, 10), (-0.15, 0.6, 10), (0.02, 1.0, 3)):
            kf, mp, scw, mi = loop_projection_case(ka, da, kb, db, sf, seed=seed, motion=motion, scale=scale)
            m_out, n = R.search_by_projection_sim3(kf, mp, scw, mi, th)
            k = len(lc)
            out["lc%d_args" % k] = np.array([seed, motion, scale, th], np.float64)
            newly = np.where((mi < 0) & (m_out >= 0), m_out, -1).astype(np.int32)
            assert np.array_equal(m_out[mi >= 0], mi[mi >= 0])          # entries matched on entry are left alone
            out["lc%d_match" % k], out["lc%d_n" % k] = newly, np.array(n)
            lc.append(n)
    out["lc_n"] = np.array(len(lc))
    print("SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th):", lc)
    # ORBmatcher::Fuse(pKF, vpMapPoints, th): the library's function with its four map-graph callees (AddObservation, AddMapPoint,
    # Replace, IsInKeyFrame) replaced by logging stand-ins; the fixture is the call log
    from matchdata import fuse_case
    fu = []
    for seed in (1, 2):
        (ka, da), (kb, db) = feats[seed]
        for motion, th in ((0.02, 3.0), (0.02, 1.5), (-0.1, 4.0)):
            kf, mp, kp_ = fuse_case(ka, da, kb, db, sf, seed=seed, motion=motion)
            nf, log = R.fuse(kf, mp, kp_, th)
            k = len(fu)
            out["fu%d_args" % k] = np.array([seed, motion, th], np.float64)
            out["fu%d_log" % k], out["fu%d_n" % k] = np.array(log, np.int32).reshape(-1, 3), np.array(nf)
            fu.append((nf, len(log)))
    # Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) (loop closing): same stand-ins; the fixture is the call log and vpReplacePoint
    fs = []
    for seed in (1, 2):
        (ka, da), (kb, db) = feats[seed]
        for motion, scale, th in ((0.02, 1.0, 4.0), (0.02, 1.6, 4.0), (-0.1, 0.7, 2.0)):
            kf, mp, kp_ = fuse_case(ka, da, kb, db, sf, seed=seed, motion=motion, scale=scale, sim3=True)
            nf, log, rep_ = R.fuse(kf, mp, kp_, th, scw=kf["scw"])
            k = len(fs)
            out["fs%d_args" % k] = np.array([seed, motion, scale, th], np.float64)
            out["fs%d_log" % k], out["fs%d_n" % k], out["fs%d_replace" % k] = np.array(log, np.int32).reshape(-1, 3), np.array(nf), rep_
            fs.append((nf, len(log), int((rep_ >= 0).sum())))
    out["fs_n"] = np.array(len(fs))
    print("Fuse(pKF, Scw, vpPoints, th, vpReplacePoint): (nFused, logged calls, replacements)", fs)
    out["fu_n"] = np.array(len(fu))
    print("Fuse(pKF, vpMapPoints, th): (nFused, logged calls)", fu)
    # ORBmatcher::SearchBySim3 on two faked KeyFrames with their own faked map points
    from matchdata import sim3_case
    s3 = []
    for seed in (1, 2):
        (ka, da), (kb, db) = feats[seed]
        for s12, th in ((1.0, 7.5), (1.3, 7.5), (0.8, 3.0)):
            kf1, kf2, mp1, mp2, sv, R12, t12, mi = sim3_case(ka, da, kb, db, sf, seed=seed, s12=s12)
            m_out, n = R.search_by_sim3(kf1, kf2, mp1, mp2, sv, R12, t12, th, mi)
            k = len(s3)
            out["s3%d_args" % k] = np.array([seed, s12, th], np.float64)
            assert np.array_equal(m_out[mi >= 0], mi[mi >= 0])
            out["s3%d_match" % k], out["s3%d_n" % k] = np.where(mi < 0, m_out, -1).astype(np.int32), np.array(n)
            s3.append(n)
    out["s3_n"] = np.array(len(s3))
    print("SearchBySim3:", s3)
    # Frame::isInFrustum on a faked Frame and faked MapPoints (GetWorldPos, GetNormal, the invariance range and PredictScale are
    # the library's own); the outputs are what the function leaves in the map points
    from matchdata import frustum_case
    fz = []
    for k, (seed, motion, lim) in enumerate(((1, 0.1, 0.5), (2, -0.3, 0.5), (3, 0.0, 0.8))):
        c = frustum_case(3000, seed=seed, motion=motion)
        r = R.is_in_frustum(c["xyz"], c["normal"], c["dist_range"], c["cam8"], c["tcw"], c["ow"], c["mbf"], c["log_sf"], c["n_levels"], lim)
        out["fz%d_args" % k] = np.array([seed, motion, lim], np.float64)
        out["fz%d_ow" % k] = c["ow"]
        for name, v in r.items():
            out["fz%d_%s" % (k, name)] = v
        fz.append(int(r["in_view"].sum()))
    out["fz_n"] = np.array(len(fz))
    print("Frame::isInFrustum: in view", fz, "of 3000")
    np.savez_compressed(os.path.join(HERE, "reference_library2.npz"), **out)


def tum_io():
    import cv2
    d = "/root/reference/Examples/RGB-D/associations"
    out = {"associations": {}}
    for name in sorted(os.listdir(d)):
        ts, rgb, dep = [], [], []
        for line in open(os.path.join(d, name), "rb").read().decode("latin-1").split("\n"):
            if line == "":
                continue
            tok = line.split()
            assert len(tok) == 4, (name, line)  # well-formed lists only: the split rule then equals the reference's extraction
            ts.append(float(tok[0])); rgb.append(tok[1]); dep.append(tok[3])
        h = hashlib.sha256()
        for t, a, b in zip(ts, rgb, dep):
            h.update(("%s|%s|%s\n" % (float(t).hex(), a, b)).encode())
        out["associations"][name] = {"n": len(ts), "sha256": h.hexdigest(), "first": [ts[0].hex(), rgb[0], dep[0]],
                                     "last": [ts[-1].hex(), rgb[-1], dep[-1]]}
    twc = cv2.gemm(np.eye(3, dtype=np.float32), np.zeros((3, 1), np.float32), -1.0, None, 0.0).ravel()
    out["identity_line"] = "%.6f %.9f %.9f %.9f %.9f %.9f %.9f %.9f\n" % ((0.0,) + tuple(float(x) for x in twc) + (0.0, 0.0, 0.0, 1.0))
    json.dump(out, open(os.path.join(HERE, "tum_io.json"), "w"), indent=1)
    print("tum_io.json:", {k: v["n"] for k, v in out["associations"].items()}, repr(out["identity_line"]))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "tum":
        return tum_io()
    if len(sys.argv) > 1 and sys.argv[1] == "refcode":
        return refcode()
    if len(sys.argv) > 1 and sys.argv[1] == "reflib":
        return reflib()
    if len(sys.argv) > 1 and sys.argv[1] == "reflib2":
        return reflib2()
    import cv2
    from plslam_b200.synth import synth_frame
    assert cv2.__version__.startswith("4.13"), cv2.__version__
    rng = np.random.default_rng(2026)

    # ---- OpenCV primitives ----
    img = np.ascontiguousarray(synth_frame(0)[100:228, 200:360])  # 128 x 160
    noise = rng.integers(0, 256, (61, 83)).astype(np.uint8)
    out = {"img": img, "noise": noise}
    out["resize_img_133x107"] = cv2.resize(img, (133, 107), interpolation=cv2.INTER_LINEAR)
    out["resize_noise_69x51"] = cv2.resize(noise, (69, 51), interpolation=cv2.INTER_LINEAR)
    out["blur_img"] = cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    out["blur_noise"] = cv2.GaussianBlur(noise, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    for th in (20, 7):
        det = cv2.FastFeatureDetector_create(th, True)
        for name, sub in (("img", img), ("cell", np.ascontiguousarray(img[16:54, 100:137]))):
            out["fast%d_%s" % (th, name)] = np.array([[int(p.pt[0]), int(p.pt[1]), int(p.response)] for p in det.detect(sub)],
                                                     np.int32).reshape(-1, 3)
    yx = rng.integers(-60000, 60000, (3000, 2)).astype(np.float32)
    yx[:4] = [[0, 0], [0, 5], [5, 0], [-3, -3]]
    out["atan2_yx"] = yx
    out["atan2_deg"] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in yx], np.float32)
    cal = dict(fx=517.306408, fy=516.469215, cx=318.643040, cy=255.313989, k1=0.262383, k2=-0.953104, p1=-0.005358,
               p2=0.002628, k3=1.163314)  # /root/reference/Examples/RGB-D/TUM1.yaml
    K = np.array([[cal["fx"], 0, cal["cx"]], [0, cal["fy"], cal["cy"]], [0, 0, 1]], np.float32)
    D = np.array([cal[k] for k in ("k1", "k2", "p1", "p2", "k3")], np.float32).reshape(5, 1)
    pts = np.stack([rng.uniform(0, 640, 2000), rng.uniform(0, 480, 2000)], 1).astype(np.float32)
    pts[:4] = [[0, 0], [640, 0], [0, 480], [640, 480]]
    out["undist_in"] = pts
    out["undist_out"] = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, D, None, K).reshape(-1, 2)
    out["undist_calib"] = np.array([cal[k] for k in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2", "k3")], np.float32)
    # the LSD front half: sigma 0.75 blur + x0.8 INTER_LINEAR_EXACT
    out["lsd_blur_noise"] = cv2.GaussianBlur(noise, (7, 7), 0.75)
    out["lsd_scaled_img"] = cv2.resize(cv2.GaussianBlur(img, (7, 7), 0.75), None, fx=0.8, fy=0.8, interpolation=cv2.INTER_LINEAR_EXACT)
    np.savez_compressed(os.path.join(HERE, "cv2_primitives.npz"), **out)

    # ---- LSD ----
    det = cv2.createLineSegmentDetector(cv2.LSD_REFINE_ADV)
    lsd = {}
    for i, (seed, W, H) in enumerate(((3, 320, 240), (4, 250, 333))):
        f = synth_frame(seed, W, H)
        lines, width, prec, nfa = det.detect(f)
        lsd["img%d" % i] = f
        lsd["lines%d" % i] = lines.reshape(-1, 4)
        lsd["width%d" % i] = width.ravel()
        lsd["prec%d" % i] = prec.ravel()
        lsd["nfa%d" % i] = nfa.ravel()
    np.savez_compressed(os.path.join(HERE, "cv2_lsd.npz"), **lsd)

    # ---- the reference's DBoW2 ----
    from oracle import bindings as ob
    ob.build()
    assert ob.dbow2_ref() is not None, "oracle/_ref/libdbow2_ref.so missing: run `make -C oracle` where /root/reference exists"
    voc_path = os.path.join(HERE, "voc_k6_L3.txt")
    ob.write_vocabulary_text(voc_path, 6, 3, seed=21)
    desc = rng.integers(0, 256, (400, 32)).astype(np.uint8)
    _, od = ob.OrbOracle().extract(synth_frame(5))
    desc[:200] = od[:200]
    ref = ob.VocReference(voc_path)
    bow = {"desc": desc}
    for lu in (0, 1, 2):
        t = ref.transform(desc, lu)
        for k, v in t.items():
            bow["lu%d_%s" % (lu, k)] = v
    a, b = desc[:200], desc[200:]
    bow["forb_distance"] = np.array([ob.VocReference.forb_distance(x, y) for x, y in zip(a, b)], np.int32)
    np.savez_compressed(os.path.join(HERE, "dbow2_ref.npz"), **bow)

    # ---- constants of the reference binary ----
    so = open("/root/reference/lib/libORB_SLAM2.so", "rb").read()
    pat = struct.unpack("<1024i", so[0x141c40:0x142c40])
    histo, th_low, th_high = struct.unpack("<3i", so[0x1269e0:0x1269ec])
    json.dump({"bit_pattern_31_sha256": hashlib.sha256(struct.pack("<1024i", *pat)).hexdigest(), "bit_pattern_31_first16": pat[:16],
               "HISTO_LENGTH": histo, "TH_LOW": th_low, "TH_HIGH": th_high,
               "source": "lib/libORB_SLAM2.so .data offset 0x141c40 (VA 0x341c40), .rodata 0x1269e0"},
              open(os.path.join(HERE, "reference_binary.json"), "w"), indent=1)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
    if len(sys.argv) == 1:
        tum_io()
        refcode()
        import subprocess
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "reflib"])  # needs a fresh process
