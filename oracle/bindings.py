"""ctypes bindings of oracle/_build/liboracle.so (the CPU restatement of the reference's
ORB / LSD / LBD / matcher hot path).  TEST INFRASTRUCTURE ONLY — see oracle/__init__.py."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28


def build(force=False):
    so = os.path.join(_HERE, "_build", "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cc", ".inc", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(so):
            so = build()
        _LIB = C.CDLL(so)
        L = _LIB
        L.oracle_fast_atan2.restype = C.c_float
        L.oracle_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.oracle_sincos.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.oracle_orb_create.restype = C.c_void_p
        L.oracle_orb_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.oracle_orb_destroy.argtypes = [C.c_void_p]
        L.oracle_orb_pattern.restype = C.POINTER(C.c_int)
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    lib().oracle_resize_linear_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dw)
    return dst


def blur7(src, k=(18, 34, 48, 56, 48, 34, 18)):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty_like(src)
    kk = np.asarray(k, np.int32)
    lib().oracle_blur7_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dst.strides[0], _p(kk))
    return dst


def fast9(img, th):
    img = np.ascontiguousarray(img, np.uint8)
    cap = img.size
    out = np.empty((cap, 3), np.int32)
    n = lib().oracle_fast9(_p(img), img.shape[1], img.shape[0], img.strides[0], int(th), _p(out), cap)
    return out[:n].copy()


def fast_atan2(y, x):
    return float(lib().oracle_fast_atan2(float(y), float(x)))


def sincos(x):
    s, c = C.c_float(), C.c_float()
    lib().oracle_sincos(float(x), C.byref(s), C.byref(c))
    return s.value, c.value


def orb_pattern():
    return np.ctypeslib.as_array(lib().oracle_orb_pattern(), shape=(1024,)).copy()


class OrbOracle:
    """Restatement of ORB_SLAM2::ORBextractor (include/ORBextractor.h:45-111)."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.nlevels = nlevels
        self.nfeatures = nfeatures
        self.h = C.c_void_p(lib().oracle_orb_create(nfeatures, scale_factor, nlevels, ini_th, min_th))

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_orb_destroy(self.h)
            self.h = None

    def set_blur_kernel(self, k):
        kk = np.asarray(k, np.int32)
        lib().oracle_orb_set_blur_kernel(self.h, _p(kk))

    def tables(self):
        n = self.nlevels
        f = [np.empty(n, np.float32) for _ in range(4)]
        quota = np.empty(n, np.int32)
        umax = np.empty(16, np.int32)
        lib().oracle_orb_tables(self.h, _p(f[0]), _p(f[1]), _p(f[2]), _p(f[3]), _p(quota), _p(umax))
        return dict(scale=f[0], inv_scale=f[1], sigma2=f[2], inv_sigma2=f[3], quota=quota, umax=umax)

    def extract(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        cap = self.nfeatures * 2 + 64
        kps = np.empty(cap, KP_DTYPE)
        desc = np.empty((cap, 32), np.uint8)
        n = lib().oracle_orb_extract(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0], _p(kps), _p(desc), cap)
        assert n <= cap
        return kps[:n].copy(), desc[:n].copy()

    def level(self, l, blurred=False):
        w, h = C.c_int(), C.c_int()
        assert lib().oracle_orb_level_size(self.h, l, C.byref(w), C.byref(h)) == 0
        out = np.empty((h.value, w.value), np.uint8)
        n = lib().oracle_orb_level_copy(self.h, l, int(blurred), _p(out))
        return out if n else None

    def candidates(self, l):
        cap = 1 << 18
        out = np.empty((cap, 3), np.int32)
        n = lib().oracle_orb_candidates(self.h, l, _p(out), cap)
        assert n <= cap
        return out[:n].copy()

    def distribute(self, xyr, minX, maxX, minY, maxY, N):
        xyr = np.ascontiguousarray(xyr, np.int32)
        out = np.empty(len(xyr) + 8, np.int32)
        n = lib().oracle_orb_distribute(self.h, _p(xyr), len(xyr), minX, maxX, minY, maxY, N, _p(out), len(out))
        return out[:n].copy()


# ---------------------------------------------------------------------------
# line side (oracle/lsd_oracle.cc)
# ---------------------------------------------------------------------------
KEYLINE_DTYPE = np.dtype([("angle", "<f4"), ("class_id", "<i4"), ("octave", "<i4"), ("pt_x", "<f4"),
                          ("pt_y", "<f4"), ("response", "<f4"), ("size", "<f4"),
                          ("startPointX", "<f4"), ("startPointY", "<f4"), ("endPointX", "<f4"),
                          ("endPointY", "<f4"), ("sPointInOctaveX", "<f4"), ("sPointInOctaveY", "<f4"),
                          ("ePointInOctaveX", "<f4"), ("ePointInOctaveY", "<f4"), ("lineLength", "<f4"),
                          ("numOfPixels", "<i4")])
assert KEYLINE_DTYPE.itemsize == 68


def gauss_table_u8(sigma, ksize):
    k = np.zeros(ksize, np.int32)
    lib().oracle_gauss_table_u8(C.c_double(sigma), ksize, _p(k))
    return k


def gauss_blur_u8(img, k):
    img = np.ascontiguousarray(img, np.uint8)
    k = np.asarray(k, np.int32)
    out = np.empty_like(img)
    lib().oracle_gauss_blur_u8(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(out), out.strides[0], _p(k), len(k))
    return out


def resize_linear_exact(img, dw, dh, inv_scale=0.0):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty((dh, dw), np.uint8)
    lib().oracle_resize_linear_exact_u8(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(out), dw, dh, dw,
                                        C.c_double(inv_scale))
    return out


def lsd_detect(img, compat=0):
    """LSD_REFINE_ADV segments: (n, 7) doubles x1, y1, x2, y2 (float32 values), width, prec, log-NFA; plus stats."""
    img = np.ascontiguousarray(img, np.uint8)
    cap = 1 << 15
    out = np.empty((cap, 7))
    st = (C.c_long * 4)()
    n = lib().oracle_lsd_detect(_p(img), img.shape[1], img.shape[0], img.strides[0], int(compat), _p(out), cap, st)
    assert n <= cap
    return out[:n].copy(), dict(regions=st[0], region_points=st[1], rects=st[2], defined=st[3])


def lsd_stage(img):
    """scaled image, level-line angles (rad, -1024 = NOTDEF) and gradient norms of the LSD front half."""
    img = np.ascontiguousarray(img, np.uint8)
    H, W = img.shape
    dw, dh = int(np.rint(W * 0.8)), int(np.rint(H * 0.8))
    sc = np.empty((dh, dw), np.uint8)
    ang = np.empty((dh, dw))
    mg = np.empty((dh, dw))
    wh = (C.c_int * 2)()
    lib().oracle_lsd_stage(_p(img), W, H, img.strides[0], _p(sc), _p(ang), _p(mg), wh)
    assert (wh[0], wh[1]) == (dw, dh)
    return sc, ang, mg


def extract_lines(img, max_lines=40, compat=0):
    """LineSegment::ExtractLineSegment restatement -> (keylines, desc[n,32], funcs[n,3], n_detected)."""
    img = np.ascontiguousarray(img, np.uint8)
    cap = 1 << 14
    kl = np.empty(cap, KEYLINE_DTYPE)
    desc = np.empty((cap, 32), np.uint8)
    funcs = np.empty((cap, 3))
    nd = C.c_int()
    n = lib().oracle_extract_lines(_p(img), img.shape[1], img.shape[0], img.strides[0], int(compat), int(max_lines),
                                   _p(kl), _p(desc), _p(funcs), cap, C.byref(nd))
    assert n >= 0
    return kl[:n].copy(), desc[:n].copy(), funcs[:n].copy(), nd.value


def lbd(img, keylines, compat=0):
    img = np.ascontiguousarray(img, np.uint8)
    kl = np.ascontiguousarray(keylines)
    n = len(kl)
    d32 = np.empty((n, 32), np.uint8)
    d72 = np.empty((n, 72), np.float32)
    lib().oracle_lbd(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(kl), n, int(compat), _p(d32), _p(d72))
    return d32, d72


def pl_sincos(x):
    s, c = C.c_double(), C.c_double()
    lib().oracle_pl_sincos(C.c_double(x), C.byref(s), C.byref(c))
    return s.value, c.value


def pl_atan2f(y, x):
    f = lib().oracle_pl_atan2f
    f.restype = C.c_float
    f.argtypes = [C.c_float, C.c_float]
    return float(f(float(y), float(x)))


# ---------------------------------------------------------------------------
# matchers (oracle/match_oracle.cc)
# ---------------------------------------------------------------------------
def descriptor_distance(a, b):
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return lib().oracle_descriptor_distance(_p(a), _p(b))


def knn2(query, train):
    query = np.ascontiguousarray(query, np.uint8).reshape(-1, 32)
    train = np.ascontiguousarray(train, np.uint8).reshape(-1, 32)
    out = np.empty((len(query), 4), np.int32)
    lib().oracle_knn2(_p(query), len(query), _p(train), len(train), _p(out))
    return out


def search_by_bow(kf, f, nnratio=0.7, check_ori=True):
    """kf / f: dicts with desc, angle, (kf: valid), nodes, start, idx.  -> (matchF, nmatches)"""
    n2 = len(f["desc"])
    match = np.empty(n2, np.int32)
    L = lib()
    L.oracle_search_by_bow.argtypes = [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_int] + \
        [C.c_void_p] * 3 + [C.c_float, C.c_int, C.c_void_p]
    n = L.oracle_search_by_bow(_p(kf["desc"]), _p(kf["angle"]), _p(kf["valid"]), len(kf["nodes"]), _p(kf["nodes"]),
                               _p(kf["start"]), _p(kf["idx"]), _p(f["desc"]), _p(f["angle"]), n2, len(f["nodes"]),
                               _p(f["nodes"]), _p(f["start"]), _p(f["idx"]), nnratio, int(check_ori), _p(match))
    return match, n


def search_by_bow_kfkf(kf1, kf2, nnratio=0.75, check_ori=True):
    """ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) -> (match12 int32 [N1], nmatches)."""
    n1, n2 = len(kf1["desc"]), len(kf2["desc"])
    m = np.empty(max(n1, 1), np.int32)
    a = [np.ascontiguousarray(x) for x in (kf1["desc"], np.asarray(kf1["angle"], np.float32), np.asarray(kf1["valid"], np.uint8),
                                           np.asarray(kf1["nodes"], np.int32), np.asarray(kf1["start"], np.int32), np.asarray(kf1["idx"], np.int32),
                                           kf2["desc"], np.asarray(kf2["angle"], np.float32), np.asarray(kf2["valid"], np.uint8),
                                           np.asarray(kf2["nodes"], np.int32), np.asarray(kf2["start"], np.int32), np.asarray(kf2["idx"], np.int32))]
    L = lib()
    L.oracle_search_by_bow_kfkf.argtypes = [C.c_void_p] * 3 + [C.c_int, C.c_int] + [C.c_void_p] * 3 + [C.c_void_p] * 3 + \
        [C.c_int, C.c_int] + [C.c_void_p] * 3 + [C.c_float, C.c_int, C.c_void_p]
    n = L.oracle_search_by_bow_kfkf(_p(a[0]), _p(a[1]), _p(a[2]), n1, len(a[3]), _p(a[3]), _p(a[4]), _p(a[5]), _p(a[6]), _p(a[7]),
                                    _p(a[8]), n2, len(a[9]), _p(a[9]), _p(a[10]), _p(a[11]), nnratio, int(check_ori), _p(m))
    return m[:n1], n


def logf(x):
    L = lib()
    L.oracle_logf.argtypes, L.oracle_logf.restype = [C.c_float], C.c_float
    return L.oracle_logf(x)


def predict_scale(max_distance, dist, log_scale_factor, n_levels):
    L = lib()
    L.oracle_predict_scale.argtypes, L.oracle_predict_scale.restype = [C.c_float, C.c_float, C.c_float, C.c_int], C.c_int
    return L.oracle_predict_scale(max_distance, dist, log_scale_factor, n_levels)


def search_by_projection_sim3(kf, mp, scw, matched_in, th):
    """ORBmatcher::SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) (layout: tests/matchdata.py loop_projection_case)
    -> (match_kf int32 [N]: map-point index newly assigned to each key-frame feature or -1, nmatches)."""
    m, n = len(mp["desc"]), len(kf["desc"])
    matched_in = np.asarray(matched_in, np.int32)
    found = np.zeros(m, bool)
    found[matched_in[matched_in >= 0]] = True
    valid = np.ascontiguousarray((np.asarray(mp["state"]) == 1) & ~found, np.uint8)
    a = [valid, np.ascontiguousarray(mp["xyz"], np.float32), np.ascontiguousarray(mp["normal"], np.float32),
         np.ascontiguousarray(mp["dist_range"], np.float32), np.ascontiguousarray(mp["desc"], np.uint8)]
    b = [np.ascontiguousarray(kf["xy"], np.float32), np.ascontiguousarray(kf["octave"], np.int32), np.ascontiguousarray(kf["desc"], np.uint8),
         np.ascontiguousarray(matched_in >= 0, np.uint8), np.ascontiguousarray(kf["grid_start"], np.int32), np.ascontiguousarray(kf["grid_items"], np.int32)]
    c = [np.ascontiguousarray(scw, np.float32).reshape(12), np.ascontiguousarray(kf["cam4"], np.float32), np.ascontiguousarray(kf["bounds4"], np.int32)]
    sf = np.ascontiguousarray(kf["scale_factors"], np.float32)
    out = np.empty(max(n, 1), np.int32)
    L = lib()
    L.oracle_search_by_projection_sim3.argtypes = [C.c_int] + [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 6 + [C.c_int, C.c_int] + \
        [C.c_void_p] * 3 + [C.c_float, C.c_float, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p]
    L.oracle_search_by_projection_sim3.restype = C.c_int
    nm = L.oracle_search_by_projection_sim3(m, *[_p(x) for x in a], n, *[_p(x) for x in b], 64, 48, *[_p(x) for x in c], float(kf["gwi"]),
                                            float(kf["ghi"]), _p(sf), len(sf), float(kf["log_sf"]), int(th), _p(out))
    return out[:n], nm


def fuse_search(kf, mp, th):
    """Matching core of ORBmatcher::Fuse(pKF, vpMapPoints, th) (layout: tests/matchdata.py fuse_case) -> best_idx int32 [M]
    (key-frame feature each map point would be fused into, -1 = none)."""
    m, n = len(mp["desc"]), len(kf["desc"])
    valid = np.ascontiguousarray(np.asarray(mp["state"]) == 1, np.uint8)
    a = [valid, np.ascontiguousarray(mp["xyz"], np.float32), np.ascontiguousarray(mp["normal"], np.float32),
         np.ascontiguousarray(mp["dist_range"], np.float32), np.ascontiguousarray(mp["desc"], np.uint8)]
    b = [np.ascontiguousarray(kf["xy"], np.float32), np.ascontiguousarray(kf["octave"], np.int32), np.ascontiguousarray(kf["uright"], np.float32),
         np.ascontiguousarray(kf["desc"], np.uint8), np.ascontiguousarray(kf["grid_start"], np.int32), np.ascontiguousarray(kf["grid_items"], np.int32)]
    c = [np.ascontiguousarray(kf["tcw"], np.float32).reshape(12), np.ascontiguousarray(kf["ow"], np.float32).reshape(3),
         np.ascontiguousarray(list(kf["cam4"]) + [kf["mbf"]], np.float32), np.ascontiguousarray(kf["bounds4"], np.int32)]
    sf = np.ascontiguousarray(kf["scale_factors"], np.float32); inv = np.ascontiguousarray(kf["inv_level_sigma2"], np.float32)
    out = np.empty(max(m, 1), np.int32)
    L = lib()
    L.oracle_fuse_search.argtypes = [C.c_int] + [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 6 + [C.c_int, C.c_int] + [C.c_void_p] * 4 + \
        [C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p]
    L.oracle_fuse_search.restype = None
    L.oracle_fuse_search(m, *[_p(x) for x in a], n, *[_p(x) for x in b], 64, 48, *[_p(x) for x in c], float(kf["gwi"]), float(kf["ghi"]),
                         _p(sf), _p(inv), len(sf), float(kf["log_sf"]), float(th), _p(out))
    return out[:m]


def sim3_inputs(kf1, kf2, mp1, mp2, matched_in):
    """The flattened arrays both the oracle and the C-ABI take for SearchBySim3 (layout: tests/matchdata.py sim3_case)."""
    n1, n2 = len(kf1["desc"]), len(kf2["desc"])
    matched_in = np.asarray(matched_in, np.int32)
    m1 = np.ascontiguousarray(matched_in >= 0, np.uint8)
    m2 = np.zeros(n2, np.uint8)
    m2[matched_in[matched_in >= 0]] = 1          # (GetIndexInKeyFrame(pKF2) of a KF2 point is its feature index)
    def side(kf, mp, m):
        return [np.ascontiguousarray(np.asarray(mp["state"]) == 1, np.uint8), m, np.ascontiguousarray(mp["xyz"], np.float32),
                np.ascontiguousarray(mp["dist_range"], np.float32), np.ascontiguousarray(mp["desc"], np.uint8),
                np.ascontiguousarray(kf["xy"], np.float32), np.ascontiguousarray(kf["octave"], np.int32), np.ascontiguousarray(kf["desc"], np.uint8),
                np.ascontiguousarray(kf["grid_start"], np.int32), np.ascontiguousarray(kf["grid_items"], np.int32),
                np.ascontiguousarray(kf["tcw"], np.float32).reshape(12), np.ascontiguousarray(kf["cam4"], np.float32),
                np.ascontiguousarray(kf["bounds4"], np.int32)]
    return n1, n2, side(kf1, mp1, m1), side(kf2, mp2, m2)


def search_by_sim3(kf1, kf2, mp1, mp2, s12, R12, t12, th, matched_in):
    """ORBmatcher::SearchBySim3 -> (match12 int32 [N1]: KF2 feature newly matched to each KF1 feature or -1, nFound)."""
    n1, n2, a, b = sim3_inputs(kf1, kf2, mp1, mp2, matched_in)
    sf = np.ascontiguousarray(kf1["scale_factors"], np.float32)
    R = np.ascontiguousarray(R12, np.float32).reshape(9); t = np.ascontiguousarray(t12, np.float32).reshape(3)
    out = np.empty(max(n1, 1), np.int32)
    L = lib()
    L.oracle_search_by_sim3.argtypes = [C.c_int] + [C.c_void_p] * 13 + [C.c_int] + [C.c_void_p] * 13 + [C.c_float, C.c_float, C.c_int, C.c_int,
                                        C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
    L.oracle_search_by_sim3.restype = C.c_int
    nf = L.oracle_search_by_sim3(n1, *[_p(x) for x in a], n2, *[_p(x) for x in b], float(kf1["gwi"]), float(kf1["ghi"]), 64, 48, _p(sf), len(sf),
                                 float(kf1["log_sf"]), float(s12), _p(R), _p(t), float(th), _p(out))
    return out[:n1], nf


def fuse_search_sim3(kf, mp, scw, th):
    """Matching core of ORBmatcher::Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) -> best_idx int32 [M]."""
    m, n = len(mp["desc"]), len(kf["desc"])
    valid = np.ascontiguousarray(np.asarray(mp["state"]) == 1, np.uint8)
    a = [valid, np.ascontiguousarray(mp["xyz"], np.float32), np.ascontiguousarray(mp["normal"], np.float32),
         np.ascontiguousarray(mp["dist_range"], np.float32), np.ascontiguousarray(mp["desc"], np.uint8)]
    b = [np.ascontiguousarray(kf["xy"], np.float32), np.ascontiguousarray(kf["octave"], np.int32), np.ascontiguousarray(kf["desc"], np.uint8),
         np.ascontiguousarray(kf["grid_start"], np.int32), np.ascontiguousarray(kf["grid_items"], np.int32)]
    c = [np.ascontiguousarray(scw, np.float32).reshape(12), np.ascontiguousarray(kf["cam4"], np.float32), np.ascontiguousarray(kf["bounds4"], np.int32)]
    sf = np.ascontiguousarray(kf["scale_factors"], np.float32)
    out = np.empty(max(m, 1), np.int32)
    L = lib()
    L.oracle_fuse_search_sim3.argtypes = [C.c_int] + [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_int] + [C.c_void_p] * 3 + \
        [C.c_float, C.c_float, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p]
    L.oracle_fuse_search_sim3.restype = None
    L.oracle_fuse_search_sim3(m, *[_p(x) for x in a], n, *[_p(x) for x in b], 64, 48, *[_p(x) for x in c], float(kf["gwi"]), float(kf["ghi"]),
                              _p(sf), len(sf), float(kf["log_sf"]), float(th), _p(out))
    return out[:m]


def fuse_replay_sim3(best_idx, mp, kf_points):
    """Bookkeeping of Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) over precomputed matches -> (nFused, log, replace int32 [M]:
    pool index M + feature of the key frame's point that should replace map point i, -1 = none)."""
    m, n = len(best_idx), len(kf_points["has"])
    slot = [m + i if kf_points["has"][i] else -1 for i in range(n)]
    kbad = np.asarray(kf_points["bad"], bool)
    ur = kf_points["uright"]
    replace = np.full(m, -1, np.int32)
    log, nf = [], 0
    for i in range(m):
        if mp["state"][i] != 1 or best_idx[i] < 0:
            continue
        idx = int(best_idx[i])
        p = slot[idx]
        if p >= 0:
            if p >= m and not kbad[p - m]:
                replace[i] = p
            elif p < m:                       # a point added earlier in this call (never bad here)
                replace[i] = p
        else:
            log.append((1, i, idx)); log.append((2, i, idx)); slot[idx] = i
        nf += 1
    return nf, log, replace


def fuse_replay(best_idx, mp, kf_points):
    """The bookkeeping of ORBmatcher::Fuse replayed over precomputed matches, against the harness stubs of
    tests/golden/reference_code.py (AddObservation adds 2 observations for a stereo feature, Replace marks the replaced point bad
    and redirects the key frame's slots): -> (nFused, log) in the format of RefLibrary.fuse."""
    m = len(best_idx)
    n = len(kf_points["has"])
    slot = [m + i if kf_points["has"][i] else -1 for i in range(n)]        # point index held by each key-frame feature
    bad = list(np.asarray(mp["state"]) == 2) + list(np.asarray(kf_points["bad"], bool))
    nobs = list(np.asarray(mp["nobs"])) + list(np.asarray(kf_points["nobs"]))
    inkf = list(np.asarray(mp["state"]) == 3)
    ur = kf_points["uright"]
    log, nf = [], 0
    for i in range(m):
        if mp["state"][i] == 0 or bad[i] or inkf[i] or best_idx[i] < 0:
            continue
        idx = int(best_idx[i])
        p = slot[idx]
        if p >= 0:
            if not bad[p]:
                if nobs[p] > nobs[i]:
                    log.append((3, i, p)); bad[i] = True
                else:
                    log.append((3, p, i)); bad[p] = True
                    slot = [i if q == p else q for q in slot]
        else:
            log.append((1, i, idx)); nobs[i] += 2 if ur[idx] >= 0 else 1; inkf[i] = True
            log.append((2, i, idx)); slot[idx] = i
        nf += 1
    return nf, log


def undistort_keylines(calib10, xy4):
    """Frame::UndistortKeyLines (UNPINNED definition, see frame_oracle.cc): n x 4 end points -> n x 4."""
    xy4 = np.ascontiguousarray(xy4, np.float32)
    out = np.empty_like(xy4)
    c = np.ascontiguousarray(calib10, np.float32)
    L = lib()
    L.oracle_undistort_keylines.argtypes, L.oracle_undistort_keylines.restype = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p], None
    L.oracle_undistort_keylines(_p(c), _p(xy4), len(xy4), _p(out))
    return out


def get_lines_in_area(query7, lines4):
    """Frame::GetLinesInArea (UNPINNED definition): query (x1, y1, x2, y2, r, minLevel, maxLevel) -> candidate indices."""
    lines4 = np.ascontiguousarray(lines4, np.float32)
    out = np.empty(max(len(lines4), 1), np.int32)
    L = lib()
    L.oracle_get_lines_in_area.argtypes = [C.c_float] * 5 + [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.oracle_get_lines_in_area.restype = C.c_int
    q = [float(v) for v in query7]
    n = L.oracle_get_lines_in_area(q[0], q[1], q[2], q[3], q[4], int(q[5]), int(q[6]), _p(lines4), len(lines4), _p(out), len(out))
    return out[:n].copy()


def line_in_frustum(sp_ep, normal, dist_range, cam8, tcw, ow, mbf, log_scale_factor, n_levels, cos_limit):
    """Frame::isInFrustum(MapLine*, float) (UNPINNED definition) -> dict(in_view, proj [M,6], level, viewcos)."""
    m = len(sp_ep)
    a = [np.ascontiguousarray(sp_ep, np.float32), np.ascontiguousarray(normal, np.float32), np.ascontiguousarray(dist_range, np.float32),
         np.ascontiguousarray(cam8, np.float32), np.ascontiguousarray(tcw, np.float32).reshape(12), np.ascontiguousarray(ow, np.float32).reshape(3)]
    out = dict(in_view=np.zeros(max(m, 1), np.uint8), proj=np.zeros((max(m, 1), 6), np.float32), level=np.zeros(max(m, 1), np.int32),
               viewcos=np.zeros(max(m, 1), np.float32))
    L = lib()
    L.oracle_line_in_frustum.argtypes = [C.c_int] + [C.c_void_p] * 6 + [C.c_float, C.c_float, C.c_int, C.c_float] + [C.c_void_p] * 4
    L.oracle_line_in_frustum.restype = None
    L.oracle_line_in_frustum(m, *[_p(x) for x in a], float(mbf), float(log_scale_factor), int(n_levels), float(cos_limit),
                             _p(out["in_view"]), _p(out["proj"]), _p(out["level"]), _p(out["viewcos"]))
    return {k: v[:m] for k, v in out.items()}


def is_in_frustum(xyz, normal, dist_range, cam8, tcw, ow, mbf, log_scale_factor, n_levels, cos_limit):
    """Frame::isInFrustum over M map points -> dict(in_view, proj [M,3], level, viewcos)."""
    m = len(xyz)
    a = [np.ascontiguousarray(xyz, np.float32), np.ascontiguousarray(normal, np.float32), np.ascontiguousarray(dist_range, np.float32),
         np.ascontiguousarray(cam8, np.float32), np.ascontiguousarray(tcw, np.float32).reshape(12), np.ascontiguousarray(ow, np.float32).reshape(3)]
    out = dict(in_view=np.zeros(max(m, 1), np.uint8), proj=np.zeros((max(m, 1), 3), np.float32), level=np.zeros(max(m, 1), np.int32),
               viewcos=np.zeros(max(m, 1), np.float32))
    L = lib()
    L.oracle_is_in_frustum.argtypes = [C.c_int] + [C.c_void_p] * 6 + [C.c_float, C.c_float, C.c_int, C.c_float] + [C.c_void_p] * 4
    L.oracle_is_in_frustum.restype = None
    L.oracle_is_in_frustum(m, *[_p(x) for x in a], float(mbf), float(log_scale_factor), int(n_levels), float(cos_limit),
                           _p(out["in_view"]), _p(out["proj"]), _p(out["level"]), _p(out["viewcos"]))
    return {k: v[:m] for k, v in out.items()}


def search_by_projection_kf(kf, cur, cam, scale_factors, log_scale_factor, tcw_cur, th, orb_dist, check_ori=True):
    """ORBmatcher::SearchByProjection(Frame&, KeyFrame*, set<MapPoint*>&, th, ORBdist) (layout: tests/matchdata.py
    relocalisation_case) -> (match_cur int32 [N2], nmatches)."""
    m, n2 = len(kf["desc"]), len(cur["desc"])
    match = np.empty(max(n2, 1), np.int32)
    a = [np.ascontiguousarray(kf["valid"], np.uint8), np.ascontiguousarray(kf["xyz"], np.float32), np.ascontiguousarray(kf["desc"], np.uint8),
         np.ascontiguousarray(kf["dist_range"], np.float32), np.ascontiguousarray(kf["angle"], np.float32)]
    b = [np.ascontiguousarray(cur["xy"], np.float32), np.ascontiguousarray(cur["octave"], np.int32), np.ascontiguousarray(cur["angle"], np.float32),
         np.ascontiguousarray(cur["desc"], np.uint8), np.ascontiguousarray(cur["taken"], np.uint8),
         np.ascontiguousarray(cur["grid_start"], np.int32), np.ascontiguousarray(cur["grid_items"], np.int32)]
    c = [np.ascontiguousarray(cam, np.float32), np.ascontiguousarray(scale_factors, np.float32), np.ascontiguousarray(tcw_cur, np.float32).reshape(12)]
    L = lib()
    L.oracle_search_by_projection_kf.argtypes = [C.c_int] + [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 9 + [C.c_int, C.c_float, C.c_void_p,
                                                                                                      C.c_float, C.c_int, C.c_int, C.c_void_p]
    L.oracle_search_by_projection_kf.restype = C.c_int
    n = L.oracle_search_by_projection_kf(m, *[_p(x) for x in a], n2, *[_p(x) for x in b], _p(c[0]), _p(c[1]), len(c[1]),
                                         float(log_scale_factor), _p(c[2]), float(th), int(orb_dist), int(check_ori), _p(match))
    return match[:n2], n


def search_by_projection(last, cur, cam, scale_factors, tcw_cur, tcw_last, th, mono=False, check_ori=True):
    n1, n2 = len(last["desc"]), len(cur["desc"])
    match = np.empty(n2, np.int32)
    L = lib()
    L.oracle_search_by_projection.argtypes = [C.c_int] + [C.c_void_p] * 6 + [C.c_int] + [C.c_void_p] * 12 + \
        [C.c_float, C.c_int, C.c_int, C.c_void_p]
    cam = np.ascontiguousarray(cam, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
    tc = np.ascontiguousarray(tcw_cur, np.float32); tl = np.ascontiguousarray(tcw_last, np.float32)
    n = L.oracle_search_by_projection(n1, _p(last["valid"]), _p(last["xyz"]), _p(last["desc"]), _p(last["octave"]),
                                      _p(last["angle"]), _p(last["obs"]), n2, _p(cur["xy"]), _p(cur["octave"]),
                                      _p(cur["angle"]), _p(cur["desc"]), _p(cur["uright"]), _p(cur["taken"]),
                                      _p(cur["grid_start"]), _p(cur["grid_items"]), _p(cam), _p(sf), _p(tc), _p(tl),
                                      th, int(mono), int(check_ori), _p(match))
    return match, n


# ---------------------------------------------------------------------------
# bag of words (oracle/bow_oracle.cc) and the reference's own DBoW2 build (oracle/_ref/libdbow2_ref.so)
# ---------------------------------------------------------------------------
def write_vocabulary_text(path, k, L, seed=0, leaf_fraction_early=0.0):
    """Synthetic vocabulary in the ORBvoc.txt format (TemplatedVocabulary.h:1362-1448): first line 'k L 0 0',
    then one line per node 'parent isLeaf d0..d31 weight' in breadth-first order; no trailing newline."""
    rng = np.random.default_rng(seed)
    lines = ["%d %d 0 0" % (k, L)]
    frontier = [0]
    next_id = 1
    for level in range(1, L + 1):
        new_frontier = []
        for parent in frontier:
            base = rng.integers(0, 256, 32)
            for _ in range(k):
                d = base ^ (rng.integers(0, 256, 32) & rng.integers(0, 256, 32) & rng.integers(0, 256, 32))
                leaf = level == L or (leaf_fraction_early > 0 and level > 1 and rng.random() < leaf_fraction_early)
                w = float(np.round(rng.random() * 10, 5)) if leaf else 0.0
                lines.append("%d %d %s %g" % (parent, int(leaf), " ".join(str(int(x)) for x in d), w))
                if not leaf:
                    new_frontier.append(next_id)
                next_id += 1
        frontier = new_frontier
    with open(path, "w") as f:
        f.write("\n".join(lines))
    return next_id


class VocOracle:
    def __init__(self, path):
        L = lib()
        L.oracle_voc_load_text.restype = C.c_void_p
        L.oracle_voc_load_text.argtypes = [C.c_char_p]
        for fn in (L.oracle_voc_free, L.oracle_voc_nodes, L.oracle_voc_words):
            fn.argtypes = [C.c_void_p]
        L.oracle_voc_export.argtypes = [C.c_void_p] * 8
        L.oracle_voc_transform_features.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_voc_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 7
        self.h = L.oracle_voc_load_text(path.encode())
        assert self.h, "cannot load vocabulary " + path
        self.n_nodes = L.oracle_voc_nodes(self.h)
        self.n_words = L.oracle_voc_words(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_voc_free(self.h)
            self.h = None

    def export(self):
        n = self.n_nodes
        out = dict(parent=np.empty(n, np.int32), first_child=np.empty(n, np.int32), n_children=np.empty(n, np.int32),
                   desc=np.empty((n, 32), np.uint8), weight=np.empty(n, np.float64), word_id=np.empty(n, np.int32))
        kL = np.empty(2, np.int32)
        lib().oracle_voc_export(self.h, _p(out["parent"]), _p(out["first_child"]), _p(out["n_children"]), _p(out["desc"]),
                                _p(out["weight"]), _p(out["word_id"]), _p(kL))
        out["k"], out["L"] = int(kL[0]), int(kL[1])
        return out

    def transform_features(self, desc, levelsup=4):
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        word = np.empty(n, np.int32); weight = np.empty(n, np.float64); node = np.empty(n, np.int32)
        lib().oracle_voc_transform_features(self.h, _p(desc), n, levelsup, _p(word), _p(weight), _p(node))
        return word, weight, node

    def transform(self, desc, levelsup=4):
        return _voc_transform(lib().oracle_voc_transform, self.h, desc, levelsup)


def _voc_transform(fn, h, desc, levelsup):
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    n = len(desc)
    ids = np.empty(n, np.uint32); vals = np.empty(n, np.float64); nb = C.c_int()
    nodes = np.empty(n, np.uint32); start = np.empty(n + 1, np.int32); idx = np.empty(n, np.uint32); nf = C.c_int()
    fn(C.c_void_p(h), _p(desc), n, levelsup, _p(ids), _p(vals), C.byref(nb), _p(nodes), _p(start), _p(idx), C.byref(nf))
    return dict(bow_ids=ids[:nb.value].copy(), bow_vals=vals[:nb.value].copy(), fv_nodes=nodes[:nf.value].copy(),
                fv_start=start[:nf.value + 1].copy(), fv_idx=idx[:start[nf.value]].copy())


_REF = None


def dbow2_ref():
    """The reference's own DBoW2 sources compiled into oracle/_ref/libdbow2_ref.so (None if never built)."""
    global _REF
    if _REF is None:
        so = os.path.join(_HERE, "_ref", "libdbow2_ref.so")
        if not os.path.exists(so):
            return None
        _REF = C.CDLL(so)
        _REF.dbow2_ref_load_text.restype = C.c_void_p
        _REF.dbow2_ref_load_text.argtypes = [C.c_char_p]
        _REF.dbow2_ref_free.argtypes = [C.c_void_p]
        _REF.dbow2_ref_size.argtypes = [C.c_void_p]
        _REF.dbow2_ref_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 7
    return _REF


class VocReference:
    def __init__(self, path):
        R = dbow2_ref()
        self.h = R.dbow2_ref_load_text(path.encode())
        assert self.h
        self.n_words = R.dbow2_ref_size(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            dbow2_ref().dbow2_ref_free(self.h)
            self.h = None

    def transform(self, desc, levelsup=4):
        return _voc_transform(dbow2_ref().dbow2_ref_transform, self.h, desc, levelsup)

    @staticmethod
    def forb_distance(a, b):
        a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
        return dbow2_ref().dbow2_ref_forb_distance(_p(a), _p(b))


# ---- Frame post-extraction steps (frame_oracle.cc) ----
def _calib10(calib):
    """calib: dict fx fy cx cy k1 k2 p1 p2 k3 bf -> 10 float32"""
    return np.array([calib[k] for k in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2", "k3", "bf")], np.float32)


def undistort_points(calib, xy):
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    out = np.empty_like(xy)
    c = _calib10(calib)
    lib().oracle_undistort_points(_p(c), _p(xy), len(xy), _p(out))
    return out


def image_bounds(calib, cols, rows):
    b = np.empty(4, np.float32)
    c = _calib10(calib)
    lib().oracle_image_bounds(_p(c), cols, rows, _p(b))
    return b


def frame_post(calib, bounds, xy, depth):
    """-> dict un_xy (n,2) f32, uright (n,), depth (n,), grid_start (64*48+1,), grid_items (inside-grid count,)"""
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    depth = np.ascontiguousarray(depth, np.float32)
    n = len(xy)
    un, ur, z = np.empty((n, 2), np.float32), np.empty(n, np.float32), np.empty(n, np.float32)
    gs, gi = np.empty(64 * 48 + 1, np.int32), np.empty(max(n, 1), np.int32)
    c, b = _calib10(calib), np.ascontiguousarray(bounds, np.float32)
    lib().oracle_frame_post(_p(c), _p(b), _p(xy), n, _p(depth), depth.shape[1], depth.shape[0], depth.strides[0] // 4,
                            _p(un), _p(ur), _p(z), _p(gs), _p(gi))
    return {"un_xy": un, "uright": ur, "depth": z, "grid_start": gs, "grid_items": gi[:gs[-1]]}


def search_local_points(mp, fr, cam4, scale_factors, th, nnratio=0.8):
    """ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th): mp / fr dicts as tests/matchdata.local_points_case."""
    m, n = len(mp["desc"]), len(fr["desc"])
    match = np.empty(max(n, 1), np.int32)
    L = lib()
    L.oracle_search_local_points.argtypes = [C.c_int] + [C.c_void_p] * 6 + [C.c_int] + [C.c_void_p] * 9 + [C.c_float, C.c_float, C.c_void_p]
    k = {a: np.ascontiguousarray(b) for a, b in mp.items()}
    f = {a: np.ascontiguousarray(b) for a, b in fr.items()}
    cam = np.ascontiguousarray(cam4, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
    cnt = L.oracle_search_local_points(m, _p(k["valid"]), _p(k["proj"]), _p(k["level"]), _p(k["viewcos"]), _p(k["desc"]), _p(k["obs"]),
                                       n, _p(f["xy"]), _p(f["octave"]), _p(f["desc"]), _p(f["uright"]), _p(f["taken"]),
                                       _p(f["grid_start"]), _p(f["grid_items"]), _p(cam), _p(sf), th, nnratio, _p(match))
    return match[:n], cnt


# ---- on-disk formats (tum_oracle.cc) ----
NAME_STRIDE = 256


def _names(buf, n):
    return [buf[i * NAME_STRIDE:(i + 1) * NAME_STRIDE].split(b"\0", 1)[0].decode("latin-1") for i in range(n)]


def tum_load_associations(path):
    """LoadImages (rgbd_tum.cc:151-176) -> (timestamps float64 [n], rgb names, depth names)"""
    L = lib()
    n = L.oracle_tum_load(os.fsencode(path), None, None, None, NAME_STRIDE, 0)
    if n < 0:
        raise FileNotFoundError(path)
    ts = np.empty(n, np.float64)
    rgb, dep = C.create_string_buffer(max(n, 1) * NAME_STRIDE), C.create_string_buffer(max(n, 1) * NAME_STRIDE)
    assert L.oracle_tum_load(os.fsencode(path), _p(ts), rgb, dep, NAME_STRIDE, n) == n
    return ts, _names(rgb.raw, n), _names(dep.raw, n)


def tum_pose(Tcw):
    """-> float32 [7]: twc, quaternion x y z w of Rwc, as System::SaveTrajectoryTUM prints them"""
    T = np.ascontiguousarray(np.asarray(Tcw, np.float32)[:3, :4])
    out = np.empty(7, np.float32)
    lib().oracle_tum_pose(_p(T), _p(out))
    return out


def tum_pose_line(timestamp, Tcw):
    T = np.ascontiguousarray(np.asarray(Tcw, np.float32)[:3, :4])
    buf = C.create_string_buffer(256)
    n = lib().oracle_tum_pose_line(C.c_double(timestamp), _p(T), buf, 256)
    assert n > 0
    return buf.raw[:n].decode("ascii")


# ---- leaf functions pinned against the reference's machine code (tests/golden/reference_code.py) ----
def radius_by_viewing_cos(v):
    L = lib()
    L.oracle_radius_by_viewing_cos.restype = C.c_float
    L.oracle_radius_by_viewing_cos.argtypes = [C.c_float]
    return float(L.oracle_radius_by_viewing_cos(float(v)))


def three_maxima(sizes):
    s = np.ascontiguousarray(sizes, np.int32)
    out = np.full(3, -1, np.int32)
    lib().oracle_three_maxima(_p(s), len(s), _p(out))
    return tuple(int(x) for x in out)


def check_dist_epipolar_line(kp1_xy, kp2_xy, F12, sigma2_kp2):
    L = lib()
    L.oracle_check_dist_epipolar_line.argtypes = [C.c_float] * 4 + [C.c_void_p, C.c_float]
    F = np.ascontiguousarray(F12, np.float32).reshape(9)
    return bool(L.oracle_check_dist_epipolar_line(float(kp1_xy[0]), float(kp1_xy[1]), float(kp2_xy[0]), float(kp2_xy[1]), _p(F),
                                                  float(sigma2_kp2)))


def epipole(R2w, t2w, Cw, fx, fy, cx, cy):
    L = lib()
    L.oracle_epipole.argtypes = [C.c_void_p] * 3 + [C.c_float] * 4 + [C.POINTER(C.c_float)] * 2
    R, t, c = (np.ascontiguousarray(x, np.float32).ravel() for x in (R2w, t2w, Cw))
    ex, ey = C.c_float(), C.c_float()
    L.oracle_epipole(_p(R), _p(t), _p(c), fx, fy, cx, cy, C.byref(ex), C.byref(ey))
    return ex.value, ey.value


def search_for_triangulation(kf1, kf2, F12, ex, ey, scale_factors, level_sigma2, only_stereo=False, check_ori=True):
    """kf*: dict desc (N,32) u8, xy (N,2) f32, angle, uright f32, has_mp u8, octave i32 (kf2), nodes/start/idx CSR
    -> (match12 int32 [N1], nmatches)"""
    g = lambda d, k, t: np.ascontiguousarray(d[k], t)
    a = dict(desc=g(kf1, "desc", np.uint8), xy=g(kf1, "xy", np.float32), angle=g(kf1, "angle", np.float32),
             uright=g(kf1, "uright", np.float32), has_mp=g(kf1, "has_mp", np.uint8), nodes=g(kf1, "nodes", np.int32),
             start=g(kf1, "start", np.int32), idx=g(kf1, "idx", np.int32))
    b = dict(desc=g(kf2, "desc", np.uint8), xy=g(kf2, "xy", np.float32), angle=g(kf2, "angle", np.float32),
             octave=g(kf2, "octave", np.int32), uright=g(kf2, "uright", np.float32), has_mp=g(kf2, "has_mp", np.uint8),
             nodes=g(kf2, "nodes", np.int32), start=g(kf2, "start", np.int32), idx=g(kf2, "idx", np.int32))
    sf, sg = np.ascontiguousarray(scale_factors, np.float32), np.ascontiguousarray(level_sigma2, np.float32)
    F = np.ascontiguousarray(F12, np.float32).reshape(9)
    n1, n2 = len(a["desc"]), len(b["desc"])
    m = np.empty(max(n1, 1), np.int32)
    L = lib()
    L.oracle_search_for_triangulation.argtypes = (
        [C.c_int] + [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 6 + [C.c_int] + [C.c_void_p] * 3 +
        [C.c_void_p] * 3 + [C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p])
    n = L.oracle_search_for_triangulation(
        n1, _p(a["desc"]), _p(a["xy"]), _p(a["angle"]), _p(a["uright"]), _p(a["has_mp"]), len(a["nodes"]), _p(a["nodes"]),
        _p(a["start"]), _p(a["idx"]), n2, _p(b["desc"]), _p(b["xy"]), _p(b["angle"]), _p(b["octave"]), _p(b["uright"]),
        _p(b["has_mp"]), len(b["nodes"]), _p(b["nodes"]), _p(b["start"]), _p(b["idx"]), _p(sf), _p(sg), _p(F), ex, ey,
        int(only_stereo), int(check_ori), _p(m))
    return m[:n1], n


def get_features_in_area(x, y, r, min_level, max_level, cam4, grid_start, grid_items, xy, octave):
    """Frame::GetFeaturesInArea on a CSR grid -> candidate indices in the reference's order"""
    L = lib()
    L.oracle_get_features_in_area.argtypes = [C.c_float] * 3 + [C.c_int] * 2 + [C.c_void_p] * 5 + [C.c_void_p, C.c_int]
    cam = np.ascontiguousarray(cam4, np.float32)
    gs, gi = np.ascontiguousarray(grid_start, np.int32), np.ascontiguousarray(grid_items, np.int32)
    xy, octv = np.ascontiguousarray(xy, np.float32), np.ascontiguousarray(octave, np.int32)
    out = np.empty(max(len(octv), 1), np.int32)
    n = L.oracle_get_features_in_area(float(x), float(y), float(r), int(min_level), int(max_level), _p(cam), _p(gs), _p(gi), _p(xy),
                                      _p(octv), _p(out), len(out))
    return out[:n].copy()


def search_for_initialization(f1, f2, cam4, prev_matched, window_size=100, nnratio=0.9, check_ori=True):
    """f1: dict xy, octave, angle, desc; f2: the same plus grid_start, grid_items -> (matches12 int32 [N1], nmatches, updated prev)"""
    g = lambda d, k, t: np.ascontiguousarray(d[k], t)
    a = [g(f1, "xy", np.float32), g(f1, "octave", np.int32), g(f1, "angle", np.float32), g(f1, "desc", np.uint8)]
    b = [g(f2, "xy", np.float32), g(f2, "octave", np.int32), g(f2, "angle", np.float32), g(f2, "desc", np.uint8),
         g(f2, "grid_start", np.int32), g(f2, "grid_items", np.int32)]
    cam = np.ascontiguousarray(cam4, np.float32)
    prev = np.array(prev_matched, np.float32).reshape(-1, 2).copy()
    n1, n2 = len(a[3]), len(b[3])
    m = np.empty(max(n1, 1), np.int32)
    L = lib()
    L.oracle_search_for_initialization.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 6 + [C.c_void_p, C.c_void_p,
                                                                                                                 C.c_int, C.c_float, C.c_int, C.c_void_p]
    n = L.oracle_search_for_initialization(n1, _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), n2, _p(b[0]), _p(b[1]), _p(b[2]), _p(b[3]),
                                           _p(b[4]), _p(b[5]), _p(cam), _p(prev), int(window_size), nnratio, int(check_ori), _p(m))
    return m[:n1].copy(), n, prev
