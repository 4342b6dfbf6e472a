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
