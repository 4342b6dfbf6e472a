"""Builds array-form matcher inputs (what Frame / KeyFrame hold) from oracle ORB features of synthetic pairs."""
import numpy as np

GRID_COLS, GRID_ROWS = 64, 48


def frame_grid(xy, W, H):
    """Frame::AssignFeaturesToGrid / PosInGrid (reference lib/libORB_SLAM2.so@0xf5fa0-0xf600f): CSR in [ix][iy] order."""
    mnMinX, mnMaxX, mnMinY, mnMaxY = np.float32(0), np.float32(W), np.float32(0), np.float32(H)
    gwi = np.float32(GRID_COLS) / (mnMaxX - mnMinX)
    ghi = np.float32(GRID_ROWS) / (mnMaxY - mnMinY)
    cells = [[] for _ in range(GRID_COLS * GRID_ROWS)]
    for i, (x, y) in enumerate(xy):
        px = int(np.round((np.float32(x) - mnMinX) * gwi))  # roundf: half away from zero; coordinates are >= 0
        py = int(np.round((np.float32(y) - mnMinY) * ghi))
        px = int(np.floor(float((np.float32(x) - mnMinX) * gwi) + 0.5))
        py = int(np.floor(float((np.float32(y) - mnMinY) * ghi) + 0.5))
        if 0 <= px < GRID_COLS and 0 <= py < GRID_ROWS:
            cells[px * GRID_ROWS + py].append(i)
    start = np.zeros(GRID_COLS * GRID_ROWS + 1, np.int32)
    start[1:] = np.cumsum([len(c) for c in cells])
    items = np.array([i for c in cells for i in c], np.int32)
    if len(items) == 0:
        items = np.zeros(1, np.int32)
    return start, items, (mnMinX, mnMaxX, mnMinY, mnMaxY, gwi, ghi)


def fake_feature_vector(desc, nbits=6, seed=0):
    """Stand-in for DBoW2 FeatureVector: node id = a hash of a few descriptor bits, so that matching
    descriptors usually share a node.  CSR sorted by node id, indices ascending inside a node."""
    rng = np.random.default_rng(seed)
    bits = rng.choice(256, nbits, replace=False)
    b = np.unpackbits(desc, axis=1, bitorder="little")[:, bits]
    node = (b * (1 << np.arange(nbits))).sum(1).astype(np.int32) * 7 + 3
    order = np.argsort(node, kind="stable")
    nodes, counts = np.unique(node, return_counts=True)
    start = np.zeros(len(nodes) + 1, np.int32)
    start[1:] = np.cumsum(counts)
    return nodes.astype(np.int32), start, order.astype(np.int32)


def projection_case(kps_last, desc_last, kps_cur, desc_cur, scale_factors, W=640, H=480, seed=0, motion=0.02):
    """TrackWithMotionModel-like inputs: last-frame keypoints back-projected at synthetic depths, a small motion."""
    rng = np.random.default_rng(seed)
    fx = fy = np.float32(520.0); cx = np.float32(W / 2 - 0.5); cy = np.float32(H / 2 - 0.5)
    bf = np.float32(40.0); mb = np.float32(bf / fx)
    n1, n2 = len(kps_last), len(kps_cur)
    z = (1.5 + rng.random(n1) * 2.0).astype(np.float32)
    xyz = np.stack([(kps_last["x"] - cx) * z / fx, (kps_last["y"] - cy) * z / fy, z], 1).astype(np.float32)
    tcw_last = np.hstack([np.eye(3), np.zeros((3, 1))]).astype(np.float32)
    ang = 0.01
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    tcw_cur = np.hstack([R, np.array([[motion], [0.005], [-motion * 2]])]).astype(np.float32)
    xy = np.stack([kps_cur["x"], kps_cur["y"]], 1).astype(np.float32)
    gs, gi, (mnx, mxx, mny, mxy, gwi, ghi) = frame_grid(xy, W, H)
    uright = np.where(rng.random(n2) < 0.7, kps_cur["x"] - bf / (1.5 + rng.random(n2) * 2.0), -1).astype(np.float32)
    last = dict(valid=(rng.random(n1) < 0.8).astype(np.uint8), xyz=np.ascontiguousarray(xyz),
                desc=np.ascontiguousarray(desc_last), octave=np.ascontiguousarray(kps_last["octave"], np.int32),
                angle=np.ascontiguousarray(kps_last["angle"], np.float32), obs=(rng.random(n1) < 0.9).astype(np.uint8))
    cur = dict(xy=np.ascontiguousarray(xy), octave=np.ascontiguousarray(kps_cur["octave"], np.int32),
               angle=np.ascontiguousarray(kps_cur["angle"], np.float32), desc=np.ascontiguousarray(desc_cur),
               uright=uright, taken=(rng.random(n2) < 0.05).astype(np.uint8), grid_start=gs, grid_items=gi)
    cam = np.array([fx, fy, cx, cy, bf, mb, mnx, mxx, mny, mxy, gwi, ghi], np.float32)
    return last, cur, cam, np.ascontiguousarray(scale_factors, np.float32), tcw_cur, tcw_last


def local_points_case(kps_map, desc_map, kps_cur, desc_cur, W=640, H=480, seed=0, jitter=3.0):
    """Tracking::SearchLocalPoints-like inputs: "map points" = the other frame's features projected near their true
    position in the current frame (what Frame::isInFrustum would have stored in mTrackProjX/Y/XR, mnTrackScaleLevel,
    mTrackViewCos), current frame = keypoints + grid."""
    rng = np.random.default_rng(seed)
    m, n = len(kps_map), len(kps_cur)
    bf = np.float32(40.0)
    z = (1.5 + rng.random(m) * 2.0).astype(np.float32)
    px = (kps_map["x"] + 3.0 + rng.normal(0, jitter, m)).astype(np.float32)
    py = (kps_map["y"] + 3.0 + rng.normal(0, jitter, m)).astype(np.float32)
    proj = np.stack([px, py, px - bf / z], 1).astype(np.float32)
    level = np.clip(kps_map["octave"] + rng.integers(-1, 2, m), 0, 7).astype(np.int32)
    mp = dict(valid=(rng.random(m) < 0.85).astype(np.uint8), proj=np.ascontiguousarray(proj), level=level,
              viewcos=(0.99 + 0.01 * rng.random(m)).astype(np.float32), desc=np.ascontiguousarray(desc_map),
              obs=(rng.random(m) < 0.9).astype(np.uint8))
    xy = np.stack([kps_cur["x"], kps_cur["y"]], 1).astype(np.float32)
    gs, gi, (mnx, mxx, mny, mxy, gwi, ghi) = frame_grid(xy, W, H)
    zc = (1.5 + rng.random(n) * 2.0).astype(np.float32)
    fr = dict(xy=np.ascontiguousarray(xy), octave=np.ascontiguousarray(kps_cur["octave"], np.int32),
              desc=np.ascontiguousarray(desc_cur),
              uright=np.where(rng.random(n) < 0.7, kps_cur["x"] - bf / zc, -1).astype(np.float32),
              taken=(rng.random(n) < 0.05).astype(np.uint8), grid_start=gs, grid_items=gi)
    cam4 = np.array([mnx, mny, gwi, ghi], np.float32)
    return mp, fr, cam4


def triangulation_case(kps, desc, seed=0, stereo_fraction=0.5, nbits=3, W=640, H=480, t21=(0.12, 0.01, 0.03)):
    """LocalMapping::CreateNewMapPoints-like inputs for ORBmatcher::SearchForTriangulation: key frame 1 = the given features at
    synthetic depths with the camera at the origin, key frame 2 = their projections after a small motion (noisy positions,
    descriptors with a few flipped bits, shuffled) plus distractors.  F12 as LocalMapping::ComputeF12 builds it
    (K^-T [t12]x R12 K^-1), the epipole from the same geometry.  Returns kf1, kf2, F12, (R2w, t2w, Cw), cam, scale_factors,
    level_sigma2."""
    rng = np.random.default_rng(seed)
    n1 = len(kps)
    fx = fy = np.float32(520.0); cx = np.float32(W / 2 - 0.5); cy = np.float32(H / 2 - 0.5)
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float64)
    sf = (1.2 ** np.arange(8)).astype(np.float32)
    z = 1.5 + 2.0 * rng.random(n1)
    X1 = np.stack([(kps["x"] - cx) * z / fx, (kps["y"] - cy) * z / fy, z], 1).astype(np.float64)
    ang = 0.03
    R21 = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    t21 = np.array(t21, np.float64)  # (0.005, 0.002, 0.15) puts the epipole inside the image
    X2 = X1 @ R21.T + t21
    oct1 = np.asarray(kps["octave"], np.int32)
    x2 = np.stack([fx * X2[:, 0] / X2[:, 2] + cx, fy * X2[:, 1] / X2[:, 2] + cy], 1)
    x2 += rng.normal(0, 0.6, x2.shape) * sf[oct1][:, None]
    inside = (x2[:, 0] > 0) & (x2[:, 0] < W) & (x2[:, 1] > 0) & (x2[:, 1] < H) & (rng.random(n1) < 0.85)
    src = np.flatnonzero(inside)
    d2 = desc[src].copy()
    for r in range(len(d2)):  # flip 0..40 random bits: distances on both sides of TH_LOW = 50
        for bit in rng.choice(256, int(rng.integers(0, 41)), replace=False):
            d2[r, bit >> 3] ^= np.uint8(1 << (bit & 7))
    ndis = 150
    dsrc = rng.integers(0, n1, ndis)
    dd = desc[dsrc].copy()
    for r in range(ndis):
        for bit in rng.choice(256, int(rng.integers(10, 60)), replace=False):
            dd[r, bit >> 3] ^= np.uint8(1 << (bit & 7))
    xy2 = np.vstack([x2[src], np.stack([rng.uniform(0, W, ndis), rng.uniform(0, H, ndis)], 1)]).astype(np.float32)
    desc2 = np.vstack([d2, dd])
    oct2 = np.concatenate([oct1[src], rng.integers(0, 8, ndis)]).astype(np.int32)
    ang2 = np.concatenate([np.asarray(kps["angle"], np.float32)[src] + rng.normal(0, 3, len(src)), rng.uniform(0, 360, ndis)])
    wild = rng.random(len(ang2)) < 0.1
    ang2 = np.mod(np.where(wild, rng.uniform(0, 360, len(ang2)), ang2), 360).astype(np.float32)
    perm = rng.permutation(len(xy2))
    xy2, desc2, oct2, ang2 = xy2[perm], np.ascontiguousarray(desc2[perm]), oct2[perm], ang2[perm]
    n2 = len(xy2)
    xy1 = np.stack([kps["x"], kps["y"]], 1).astype(np.float32)
    ur = lambda x, n: np.where(rng.random(n) < stereo_fraction, x - 40.0 / (1.5 + 2 * rng.random(n)), -1).astype(np.float32)
    nodes1, start1, idx1 = fake_feature_vector(desc, nbits, seed=seed + 100)
    nodes2, start2, idx2 = fake_feature_vector(desc2, nbits, seed=seed + 100)
    kf1 = dict(desc=np.ascontiguousarray(desc), xy=xy1, angle=np.asarray(kps["angle"], np.float32), uright=ur(xy1[:, 0], n1),
               has_mp=(rng.random(n1) < 0.2).astype(np.uint8), nodes=nodes1, start=start1, idx=idx1)
    kf2 = dict(desc=desc2, xy=xy2, angle=ang2, octave=oct2, uright=ur(xy2[:, 0], n2), has_mp=(rng.random(n2) < 0.2).astype(np.uint8),
               nodes=nodes2, start=start2, idx=idx2)
    R12, t12 = R21.T, -R21.T @ t21
    tx = np.array([[0, -t12[2], t12[1]], [t12[2], 0, -t12[0]], [-t12[1], t12[0], 0]])
    F12 = (np.linalg.inv(K).T @ tx @ R12 @ np.linalg.inv(K)).astype(np.float32)
    pose = (R21.astype(np.float32), t21.astype(np.float32), np.zeros(3, np.float32))
    return kf1, kf2, F12, pose, (fx, fy, cx, cy), sf, (sf * sf).astype(np.float32)


def relocalisation_case(kps_kf, desc_kf, kps_cur, desc_cur, scale_factors, W=640, H=480, seed=0, motion=0.05):
    """Tracking::Relocalization-like inputs for ORBmatcher::SearchByProjection(Frame&, KeyFrame*, set<MapPoint*>&, th, ORBdist):
    the key frame's map points = its keypoints back-projected at synthetic depths (key frame at the origin), with the scale
    invariance range MapPoint::UpdateNormalAndDepth would give them (some pushed out of range); the current frame a small
    motion away, some of its features already holding a map point."""
    rng = np.random.default_rng(seed)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    fx = fy = np.float32(520.0); cx = np.float32(W / 2 - 0.5); cy = np.float32(H / 2 - 0.5)
    m, n2 = len(kps_kf), len(kps_cur)
    z = (1.5 + rng.random(m) * 2.0).astype(np.float32)
    xyz = np.stack([(kps_kf["x"] - cx) * z / fx, (kps_kf["y"] - cy) * z / fy, z], 1).astype(np.float32)
    dist = np.linalg.norm(xyz.astype(np.float64), axis=1)
    dmax = (dist * sf[kps_kf["octave"]] * rng.uniform(0.6, 1.5, m)).astype(np.float32)
    dmin = (dmax / sf[-1]).astype(np.float32)
    ang = 0.015
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    tcw = np.hstack([R, np.array([[motion], [0.004], [-motion * 1.5]])]).astype(np.float32)
    xy = np.stack([kps_cur["x"], kps_cur["y"]], 1).astype(np.float32)
    gs, gi, (mnx, mxx, mny, mxy, gwi, ghi) = frame_grid(xy, W, H)
    state = rng.choice(np.array([0, 1, 2, 3], np.uint8), m, p=[0.1, 0.75, 0.05, 0.1]).astype(np.uint8)
    kf = dict(state=state, valid=(state == 1).astype(np.uint8), xyz=np.ascontiguousarray(xyz), desc=np.ascontiguousarray(desc_kf),
              dist_range=np.ascontiguousarray(np.stack([dmin, dmax], 1)), angle=np.ascontiguousarray(kps_kf["angle"], np.float32))
    cur = dict(xy=np.ascontiguousarray(xy), octave=np.ascontiguousarray(kps_cur["octave"], np.int32),
               angle=np.ascontiguousarray(kps_cur["angle"], np.float32), desc=np.ascontiguousarray(desc_cur),
               taken=(rng.random(n2) < 0.05).astype(np.uint8), grid_start=gs, grid_items=gi)
    cam = np.array([fx, fy, cx, cy, mnx, mxx, mny, mxy, gwi, ghi], np.float32)
    log_sf = np.float32(np.log(np.float32(1.2)))
    return kf, cur, cam, sf, log_sf, tcw


def frustum_case(n=3000, W=640, H=480, seed=0, motion=0.1):
    """Tracking::SearchLocalPoints-like inputs for Frame::isInFrustum: local map points scattered in and around the view frustum
    (some behind the camera, outside the image, out of their scale-invariance range or seen from too oblique an angle)."""
    rng = np.random.default_rng(seed)
    fx = fy = np.float32(520.0); cx = np.float32(W / 2 - 0.5); cy = np.float32(H / 2 - 0.5)
    ang = 0.05
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    t = np.array([[motion], [0.01], [-motion]])
    tcw = np.hstack([R, t]).astype(np.float32)
    Rf, tf = tcw[:, :3].astype(np.float64), tcw[:, 3].astype(np.float64)
    ow = (-(Rf.T @ tf)).astype(np.float32)          # (the fixture stores it: Frame::mOw is an input of isInFrustum)
    z = rng.uniform(-1.0, 6.0, n)
    x = rng.uniform(-1.3, 1.3, n) * np.abs(z) * (W / 2) / fx
    y = rng.uniform(-1.3, 1.3, n) * np.abs(z) * (H / 2) / fy
    xyz = np.stack([x, y, z], 1).astype(np.float32)
    d = np.linalg.norm(xyz.astype(np.float64) - ow.astype(np.float64), axis=1)
    dmax = (d * rng.uniform(0.5, 4.0, n)).astype(np.float32)
    dmin = (dmax / np.float32(1.2) ** 7 * rng.uniform(0.5, 1.5, n)).astype(np.float32)
    po = xyz.astype(np.float64) - ow.astype(np.float64)
    nrm = po / np.maximum(np.linalg.norm(po, axis=1, keepdims=True), 1e-9)
    nrm = nrm + rng.normal(0, 0.7, nrm.shape)
    nrm = (nrm / np.linalg.norm(nrm, axis=1, keepdims=True)).astype(np.float32)
    cam8 = np.array([fx, fy, cx, cy, 0, W, 0, H], np.float32)
    return dict(xyz=np.ascontiguousarray(xyz), normal=np.ascontiguousarray(nrm), dist_range=np.ascontiguousarray(np.stack([dmin, dmax], 1)),
                cam8=cam8, tcw=tcw, ow=ow, mbf=np.float32(40.0), log_sf=np.float32(np.log(np.float32(1.2))), n_levels=8)


def loop_projection_case(kps_kf, desc_kf, kps_src, desc_src, scale_factors, W=640, H=480, seed=0, motion=0.05, scale=1.0):
    """LoopClosing-like inputs for ORBmatcher::SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th): the map points are the
    OTHER frame's features placed at synthetic depths so that they project near their true position in the key frame (jitter),
    with normals, scale-invariance ranges and a similarity Scw = [s R | s t]; some key-frame features are matched on entry."""
    rng = np.random.default_rng(seed)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    fx = fy = np.float32(520.0); cx = np.float32(W / 2 - 0.5); cy = np.float32(H / 2 - 0.5)
    n, m = len(kps_kf), len(kps_src)
    ang = 0.02
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    t = np.array([motion, 0.004, -motion * 1.5])
    # camera-frame points seen near the source features' pixels (+3 px shift of the synthetic pair, jitter), then to world
    z = (1.5 + rng.random(m) * 2.0)
    u = kps_src["x"].astype(np.float64) - 3.0 + rng.normal(0, 2.0, m)
    v = kps_src["y"].astype(np.float64) - 3.0 + rng.normal(0, 2.0, m)
    pc = np.stack([(u - cx) * z / fx, (v - cy) * z / fy, z], 1)
    z[rng.random(m) < 0.03] *= -1.0                                    # a few behind the camera
    pc[:, 2] = z
    xyz = ((pc - t) @ R).astype(np.float32)                            # Pw = R^T (Pc - t)
    ow = -(R.T @ t)
    po = xyz.astype(np.float64) - ow
    d = np.linalg.norm(po, axis=1)
    dmax = (d * sf[np.clip(kps_src["octave"], 0, len(sf) - 1)] * rng.uniform(0.7, 1.4, m)).astype(np.float32)
    dmin = (dmax / sf[-1]).astype(np.float32)
    nrm = po / np.maximum(d[:, None], 1e-9) + rng.normal(0, 0.5, po.shape)
    nrm = (nrm / np.linalg.norm(nrm, axis=1, keepdims=True)).astype(np.float32)
    scw = (np.hstack([R, t[:, None]]) * scale).astype(np.float32)
    xy = np.stack([kps_kf["x"], kps_kf["y"]], 1).astype(np.float32)
    gs, gi, (mnx, mxx, mny, mxy, gwi, ghi) = frame_grid(xy, W, H)
    kf = dict(xy=np.ascontiguousarray(xy), octave=np.ascontiguousarray(kps_kf["octave"], np.int32), desc=np.ascontiguousarray(desc_kf),
              angle=np.ascontiguousarray(kps_kf["angle"], np.float32), grid_start=gs, grid_items=gi,
              cam4=np.array([fx, fy, cx, cy], np.float32), bounds4=np.array([0, 0, W, H], np.int32), gwi=gwi, ghi=ghi,
              scale_factors=sf, log_sf=np.float32(np.log(np.float32(1.2))))
    state = rng.choice(np.array([1, 2], np.uint8), m, p=[0.93, 0.07]).astype(np.uint8)
    mp = dict(state=state, xyz=np.ascontiguousarray(xyz), normal=np.ascontiguousarray(nrm),
              dist_range=np.ascontiguousarray(np.stack([dmin, dmax], 1)), desc=np.ascontiguousarray(desc_src))
    matched_in = np.full(n, -1, np.int32)
    pre = rng.choice(n, n // 10, replace=False)
    matched_in[pre] = rng.choice(m, len(pre), replace=False)          # already matched: those map points are "already found"
    return kf, mp, scw, matched_in


def fuse_case(kps_kf, desc_kf, kps_src, desc_src, scale_factors, W=640, H=480, seed=0, motion=0.05, scale=1.0, sim3=False):
    """LocalMapping::SearchInNeighbors-like inputs for ORBmatcher::Fuse(pKF, vpMapPoints, th): loop_projection_case with a rigid
    pose, plus what Fuse reads on top (mvuRight, mvInvLevelSigma2, mbf), NULL / bad / already-in-key-frame candidates with
    observation counts, and the map points the key frame already holds."""
    kf, mp, scw, _ = loop_projection_case(kps_kf, desc_kf, kps_src, desc_src, scale_factors, W, H, seed=seed, motion=motion, scale=scale)
    rng = np.random.default_rng(seed + 500)
    n, m = len(kps_kf), len(kps_src)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    bf = np.float32(40.0)
    kf["uright"] = np.where(rng.random(n) < 0.7, kf["xy"][:, 0] - bf / (1.5 + rng.random(n) * 2.0), -1).astype(np.float32)
    kf["inv_level_sigma2"] = (np.float32(1.0) / (sf * sf)).astype(np.float32)
    kf["mbf"] = bf
    kf["tcw"] = scw.astype(np.float32)
    R, t = scw[:, :3].astype(np.float64), scw[:, 3].astype(np.float64)
    kf["ow"] = (-(R.T @ t)).astype(np.float32)
    st = mp["state"].copy()
    r = rng.random(m)
    st[r < 0.05] = 0
    st[(r >= 0.05) & (r < 0.10)] = 3
    mp["state"] = st
    mp["nobs"] = rng.integers(1, 9, m).astype(np.int32)
    kf_points = dict(has=(rng.random(n) < 0.6).astype(np.uint8), nobs=rng.integers(1, 9, n).astype(np.int32),
                     bad=(rng.random(n) < 0.05).astype(np.uint8), uright=kf["uright"])
    if sim3:   # Fuse(pKF, Scw, ...): no NULL candidates, no IsInKeyFrame test; some candidates ARE points the key frame already holds
        kf["scw"] = scw.astype(np.float32)
        st = np.where(st == 0, 1, np.where(st == 3, 4, st)).astype(np.uint8)
        good = np.flatnonzero((kf_points["has"] == 1) & (kf_points["bad"] == 0))
        mp["alias"] = rng.choice(good, m).astype(np.int32)
        mp["state"] = st
    return kf, mp, kf_points


def sim3_case(kps1, desc1, kps2, desc2, scale_factors, W=640, H=480, seed=0, s12=1.2):
    """LoopClosing::ComputeSim3-like inputs for ORBmatcher::SearchBySim3: two key frames with their own map points (one per
    feature, some missing or bad) and a similarity (s12, R12, t12) from camera 2 to camera 1 under which the points of each key
    frame project near the corresponding features of the other (the synthetic pair differs by ~3 px)."""
    rng = np.random.default_rng(seed + 900)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    fx = fy = np.float32(520.0); cx = np.float32(W / 2 - 0.5); cy = np.float32(H / 2 - 0.5)
    def rot(ax, ang):
        c, s_ = np.cos(ang), np.sin(ang)
        return np.array([[c, 0, s_], [0, 1, 0], [-s_, 0, c]]) if ax == 1 else np.array([[1, 0, 0], [0, c, -s_], [0, s_, c]])
    R12 = rot(1, 0.03) @ rot(0, -0.02); t12 = np.array([0.05, -0.02, 0.04])
    R1w, t1w = rot(1, 0.2), np.array([0.3, 0.1, -0.2])
    R2w, t2w = rot(0, -0.1), np.array([-0.1, 0.2, 0.1])
    def kf_of(kps, desc, R, t):
        xy = np.stack([kps["x"], kps["y"]], 1).astype(np.float32)
        gs, gi, (mnx, mxx, mny, mxy, gwi, ghi) = frame_grid(xy, W, H)
        return dict(xy=np.ascontiguousarray(xy), octave=np.ascontiguousarray(kps["octave"], np.int32), desc=np.ascontiguousarray(desc),
                    grid_start=gs, grid_items=gi, cam4=np.array([fx, fy, cx, cy], np.float32), bounds4=np.array([0, 0, W, H], np.int32),
                    gwi=gwi, ghi=ghi, scale_factors=sf, log_sf=np.float32(np.log(np.float32(1.2))),
                    tcw=np.hstack([R, t[:, None]]).astype(np.float32))
    kf1, kf2 = kf_of(kps1, desc1, R1w, t1w), kf_of(kps2, desc2, R2w, t2w)
    def points(kps, desc, shift, to_other, R, t):
        n = len(kps)
        z = 1.5 + rng.random(n) * 2.0
        u = kps["x"].astype(np.float64) + shift + rng.normal(0, 2.0, n)
        v = kps["y"].astype(np.float64) + shift + rng.normal(0, 2.0, n)
        z[rng.random(n) < 0.03] *= -1.0
        p_other = np.stack([(u - cx) * np.abs(z) / fx, (v - cy) * np.abs(z) / fy, z], 1)   # in the OTHER camera
        p_own = to_other(p_other)                                                          # in this key frame's camera
        pw = (p_own - t) @ R                                                               # world
        d = np.linalg.norm(p_other, axis=1)
        dmax = (d * sf[np.clip(kps["octave"], 0, len(sf) - 1)] * rng.uniform(0.7, 1.4, n)).astype(np.float32)
        state = rng.choice(np.array([0, 1, 2], np.uint8), n, p=[0.15, 0.8, 0.05]).astype(np.uint8)
        return dict(state=state, xyz=np.ascontiguousarray(pw.astype(np.float32)), normal=np.zeros((n, 3), np.float32),
                    dist_range=np.ascontiguousarray(np.stack([dmax / sf[-1], dmax], 1).astype(np.float32)), desc=np.ascontiguousarray(desc))
    mp1 = points(kps1, desc1, +3.0, lambda p2: s12 * (p2 @ R12.T) + t12, R1w, t1w)        # p1 = s12 R12 p2 + t12
    mp2 = points(kps2, desc2, -3.0, lambda p1: ((p1 - t12) @ R12) / s12, R2w, t2w)        # p2 = R12^T (p1 - t12) / s12
    n1 = len(kps1)
    matched_in = np.full(n1, -1, np.int32)
    cand = np.flatnonzero(mp2["state"] == 1)
    pre = rng.choice(n1, n1 // 12, replace=False)
    matched_in[pre] = rng.choice(cand, len(pre), replace=False)
    return kf1, kf2, mp1, mp2, np.float32(s12), R12.astype(np.float32), t12.astype(np.float32), matched_in
