"""Python harness over libplslam_b200.so (the C-ABI in include/plslam_b200.h).

The product is the CUDA library + the C++ classes in rgbd-pl-slam_b200/host/; this module only
binds the C-ABI with ctypes so tests and bench.py can drive it, and uses torch for device memory,
streams and torch.distributed.  There is no CPU fallback: importing works anywhere, but every
compute call raises PlslamError when the library or a CUDA device is missing.
"""
import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.environ.get("PLSLAM_LIB") or os.path.join(_ROOT, "libplslam_b200.so")  # PLSLAM_LIB: experiment builds (tools/)

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
KEYLINE_DTYPE = np.dtype([("angle", "<f4"), ("class_id", "<i4"), ("octave", "<i4"), ("pt_x", "<f4"),
                          ("pt_y", "<f4"), ("response", "<f4"), ("size", "<f4"),
                          ("startPointX", "<f4"), ("startPointY", "<f4"), ("endPointX", "<f4"),
                          ("endPointY", "<f4"), ("sPointInOctaveX", "<f4"), ("sPointInOctaveY", "<f4"),
                          ("ePointInOctaveX", "<f4"), ("ePointInOctaveY", "<f4"), ("lineLength", "<f4"),
                          ("numOfPixels", "<i4")])

OK, ERR_INVALID, ERR_CUDA, ERR_CAPACITY, ERR_OVERFLOW = 0, 1, 2, 3, 4


class PlslamError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("plslam_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    """Load libplslam_b200.so (built in-tree by `make -C rgbd-pl-slam_b200` / __graft_entry__.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PlslamError(ERR_CUDA, "CUDA extension %s is missing — run __graft_entry__.build()" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.plslam_last_error.restype = C.c_char_p
        L.plslam_version.restype = C.c_char_p
        L.plslam_orb_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.plslam_orb_destroy.argtypes = [C.c_void_p]
        L.plslam_orb_destroy.restype = None
        _lib = L
    return _lib


def _check(rc):
    if rc != OK:
        raise PlslamError(rc, lib().plslam_last_error().decode())


def _vp(x):
    """void* of a numpy array, a torch tensor or an int."""
    if x is None:
        return C.c_void_p(0)
    if isinstance(x, np.ndarray):
        return C.c_void_p(x.ctypes.data)
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    return C.c_void_p(int(x))


def _stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


class ORBextractor:
    """Mirror of ORB_SLAM2::ORBextractor (reference include/ORBextractor.h:45-111) over the C-ABI."""

    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7):
        self._h = C.c_void_p()
        _check(lib().plslam_orb_create(C.byref(self._h), nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST))
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self.max_keypoints = lib().plslam_orb_max_keypoints(self._h)

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                lib().plslam_orb_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:  # interpreter shutdown
            pass

    # getters (ORBextractor.h:63-83)
    def GetLevels(self):
        return self.nlevels

    def tables(self):
        n = self.nlevels
        f = [np.empty(n, np.float32) for _ in range(4)]
        quota = np.empty(n, np.int32)
        umax = np.empty(16, np.int32)
        _check(lib().plslam_orb_tables(self._h, _vp(f[0]), _vp(f[1]), _vp(f[2]), _vp(f[3]), _vp(quota), _vp(umax)))
        return dict(scale=f[0], inv_scale=f[1], sigma2=f[2], inv_sigma2=f[3], quota=quota, umax=umax)

    def GetScaleFactors(self):
        return self.tables()["scale"]

    def set_blur_kernel(self, k):
        kk = np.asarray(k, np.int32)
        _check(lib().plslam_orb_set_blur_kernel(self._h, _vp(kk)))

    def __call__(self, image, mask=None):
        """operator()(image, mask, keypoints, descriptors): one host image -> (keypoints, descriptors)."""
        if image is None or image.size == 0:
            return np.empty(0, KP_DTYPE), np.empty((0, 32), np.uint8)
        image = np.ascontiguousarray(image, np.uint8)
        cap = self.max_keypoints
        kps = np.empty(cap, KP_DTYPE)
        desc = np.empty((cap, 32), np.uint8)
        n = C.c_int(0)
        _check(lib().plslam_orb_extract(self._h, _vp(image), image.shape[1], image.shape[0], image.strides[0],
                                        _vp(kps), _vp(desc), cap, C.byref(n)))
        return kps[:n.value].copy(), desc[:n.value].copy()

    def extract_batch_host(self, images):
        """images: (B, H, W) uint8 numpy (pinned or pageable).  Returns (kps[B, cap], desc[B, cap, 32], counts[B])."""
        images = np.ascontiguousarray(images, np.uint8)
        B, H, W = images.shape
        cap = self.max_keypoints
        kps = np.empty((B, cap), KP_DTYPE)
        desc = np.empty((B, cap, 32), np.uint8)
        counts = np.empty(B, np.int32)
        _check(lib().plslam_orb_extract_batch_host(self._h, _vp(images), B, W, H, images.strides[1],
                                                   C.c_size_t(images.strides[0]), _vp(kps), _vp(desc), cap, _vp(counts)))
        return kps, desc, counts

    def extract_batch_device(self, d_images, out=None, stream=None):
        """d_images: (B, H, W) uint8 CUDA tensor.  Asynchronous on the current torch stream.
        Returns CUDA tensors (kps[B, cap, 7] int32 view, desc[B, cap, 32] uint8, counts[B] int32)."""
        import torch
        assert d_images.is_cuda and d_images.dtype == torch.uint8 and d_images.dim() == 3
        B, H, W = d_images.shape
        assert d_images.stride(2) == 1
        cap = self.max_keypoints
        if out is None:
            out = (torch.empty((B, cap, 7), dtype=torch.int32, device=d_images.device),
                   torch.empty((B, cap, 32), dtype=torch.uint8, device=d_images.device),
                   torch.empty((B,), dtype=torch.int32, device=d_images.device))
        kps, desc, counts = out
        _check(lib().plslam_orb_extract_batch_device(self._h, _vp(d_images), B, W, H, d_images.stride(1),
                                                     C.c_size_t(d_images.stride(0)), _vp(kps), _vp(desc), cap,
                                                     _vp(counts), _stream_ptr(stream)))
        return out

    def check_status(self, stream=None):
        _check(lib().plslam_orb_check_status(self._h, _stream_ptr(stream)))

    # parity / mvImagePyramid accessors
    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        _check(lib().plslam_orb_level_size(self._h, level, C.byref(w), C.byref(h)))
        return w.value, h.value

    def level(self, frame, level, blurred=False):
        w, h = self.level_size(level)
        out = np.empty((h, w), np.uint8)
        _check(lib().plslam_orb_copy_level(self._h, frame, level, int(blurred), _vp(out), C.c_size_t(out.size)))
        return out

    def candidates(self, frame, level):
        w, h = self.level_size(level)
        cap = ((w + 1) // 2) * ((h + 1) // 2)
        out = np.empty((cap, 3), np.int32)
        n = C.c_int()
        _check(lib().plslam_orb_copy_candidates(self._h, frame, level, _vp(out), cap, C.byref(n)))
        return out[:n.value].copy()


def kps_from_tensor(t):
    """(.., 7) int32 torch tensor (device layout of plslam_keypoint_t) -> numpy structured array."""
    a = t.detach().cpu().numpy()
    return a.view(KP_DTYPE).reshape(a.shape[:-1])


class LineSegment:
    """Mirror of ORB_SLAM2::LineSegment (reference include/ExtractLineSegment.h:30-55) over the C-ABI."""

    def __init__(self, max_lines=40):
        self._h = C.c_void_p()
        L = lib()
        L.plslam_lines_create.argtypes = [C.POINTER(C.c_void_p)]
        L.plslam_lines_destroy.argtypes = [C.c_void_p]
        L.plslam_lines_destroy.restype = None
        _check(L.plslam_lines_create(C.byref(self._h)))
        _check(L.plslam_lines_set_max_lines(self._h, int(max_lines)))
        self.capacity = L.plslam_lines_capacity(self._h)

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                lib().plslam_lines_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def ExtractLineSegment(self, img):
        """-> (keylines[n], ldesc[n, 32], keylineFunctions[n, 3])"""
        if img is None or img.size == 0:
            return np.empty(0, KEYLINE_DTYPE), np.empty((0, 32), np.uint8), np.empty((0, 3))
        img = np.ascontiguousarray(img, np.uint8)
        cap = self.capacity
        kl = np.empty(cap, KEYLINE_DTYPE)
        desc = np.empty((cap, 32), np.uint8)
        funcs = np.empty((cap, 3), np.float64)
        n = C.c_int(0)
        _check(lib().plslam_lines_extract(self._h, _vp(img), img.shape[1], img.shape[0], img.strides[0], _vp(kl),
                                          _vp(desc), _vp(funcs), cap, C.byref(n)))
        n = n.value
        return kl[:n].copy(), desc[:n].copy(), funcs[:n].copy()

    def extract_batch_host(self, images):
        images = np.ascontiguousarray(images, np.uint8)
        B, H, W = images.shape
        cap = self.capacity
        kl = np.empty((B, cap), KEYLINE_DTYPE)
        desc = np.empty((B, cap, 32), np.uint8)
        funcs = np.empty((B, cap, 3), np.float64)
        counts = np.empty(B, np.int32)
        _check(lib().plslam_lines_extract_batch_host(self._h, _vp(images), B, W, H, images.strides[1],
                                                     C.c_size_t(images.strides[0]), _vp(kl), _vp(desc), _vp(funcs), cap,
                                                     _vp(counts)))
        return kl, desc, funcs, counts

    def extract_batch_device(self, d_images, out=None, stream=None):
        import torch
        assert d_images.is_cuda and d_images.dtype == torch.uint8 and d_images.dim() == 3 and d_images.stride(2) == 1
        B, H, W = d_images.shape
        cap = self.capacity
        if out is None:
            dev = d_images.device
            out = (torch.empty((B, cap, 17), dtype=torch.int32, device=dev),
                   torch.empty((B, cap, 32), dtype=torch.uint8, device=dev),
                   torch.empty((B, cap, 3), dtype=torch.float64, device=dev),
                   torch.empty((B,), dtype=torch.int32, device=dev))
        kl, desc, funcs, counts = out
        _check(lib().plslam_lines_extract_batch_device(self._h, _vp(d_images), B, W, H, d_images.stride(1),
                                                       C.c_size_t(d_images.stride(0)), _vp(kl), _vp(desc), _vp(funcs),
                                                       cap, _vp(counts), _stream_ptr(stream)))
        return out

    def compute_lbd(self, img, keylines):
        """BinaryDescriptor::compute on given key lines (KEYLINE_DTYPE array) -> [n, 32] LBD bytes."""
        img = np.ascontiguousarray(img, np.uint8)
        kl = np.ascontiguousarray(keylines, KEYLINE_DTYPE)
        desc = np.empty((len(kl), 32), np.uint8)
        _check(lib().plslam_lines_compute_lbd(self._h, _vp(img), img.shape[1], img.shape[0], img.strides[0], _vp(kl), len(kl),
                                              _vp(desc)))
        return desc

    def check_status(self, stream=None):
        _check(lib().plslam_lines_check_status(self._h, _stream_ptr(stream)))

    # parity accessors
    def scaled(self, frame):
        w, h = C.c_int(), C.c_int()
        _check(lib().plslam_lines_scaled_size(self._h, C.byref(w), C.byref(h)))
        out = np.empty((h.value, w.value), np.uint8)
        _check(lib().plslam_lines_copy_scaled(self._h, frame, _vp(out), C.c_size_t(out.size)))
        return out

    def level_lines(self, frame):
        w, h = C.c_int(), C.c_int()
        _check(lib().plslam_lines_scaled_size(self._h, C.byref(w), C.byref(h)))
        deg = np.empty((h.value, w.value), np.float32)
        g2 = np.empty((h.value, w.value), np.int32)
        _check(lib().plslam_lines_copy_level_lines(self._h, frame, _vp(deg), _vp(g2), C.c_size_t(deg.size)))
        return deg, g2

    def segments(self, frame):
        cap = 1 << 17
        out = np.empty((cap, 7), np.float64)
        n = C.c_int()
        _check(lib().plslam_lines_copy_segments(self._h, frame, _vp(out), cap, C.byref(n)))
        return out[:n.value].copy()


def keylines_from_tensor(t):
    a = t.detach().cpu().numpy()
    return a.view(KEYLINE_DTYPE).reshape(a.shape[:-1])


# ---------------------------------------------------------------------------
# matchers (ORBmatcher / LSDmatcher Hamming cores)
# ---------------------------------------------------------------------------
TH_LOW, TH_HIGH, HISTO_LENGTH = 50, 100, 30
FRAME_GRID_COLS, FRAME_GRID_ROWS = 64, 48


class KnnJob(C.Structure):
    _fields_ = [("query", C.c_void_p), ("train", C.c_void_p), ("out", C.c_void_p), ("nq", C.c_int32), ("nt", C.c_int32)]


class BowJob(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("kf_desc", "kf_angle", "kf_valid", "kf_nodes", "kf_start", "kf_idx", "f_desc",
                                          "f_angle", "f_nodes", "f_start", "f_idx", "match_f", "nmatches")] + \
               [("n1", C.c_int32), ("n2", C.c_int32), ("n_kf_nodes", C.c_int32), ("n_f_nodes", C.c_int32),
                ("nnratio", C.c_float), ("check_orientation", C.c_int32), ("f_valid", C.c_void_p), ("strict_low", C.c_int32)]


class ProjJob(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("last_valid", "last_xyz", "last_desc", "last_octave", "last_angle", "last_obs",
                                          "cur_xy", "cur_octave", "cur_angle", "cur_desc", "cur_uright", "cur_taken",
                                          "grid_start", "grid_items", "scale_factors", "match_cur", "nmatches")] + \
               [("cam", C.c_float * 12), ("tcw_cur", C.c_float * 12), ("tcw_last", C.c_float * 12), ("th", C.c_float),
                ("n1", C.c_int32), ("n2", C.c_int32), ("mono", C.c_int32), ("check_orientation", C.c_int32),
                ("report_removed", C.c_int32), ("last_dist_range", C.c_void_p), ("mode", C.c_int32), ("orb_dist", C.c_int32),
                ("n_levels", C.c_int32), ("log_scale_factor", C.c_float)]


class TriJob(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("kf1_desc", "kf1_xy", "kf1_angle", "kf1_uright", "kf1_has_mp", "kf1_nodes", "kf1_start",
                                          "kf1_idx", "kf2_desc", "kf2_xy", "kf2_angle", "kf2_octave", "kf2_uright", "kf2_has_mp",
                                          "kf2_nodes", "kf2_start", "kf2_idx", "scale_factors", "level_sigma2", "match12",
                                          "nmatches")] + \
               [("F12", C.c_float * 9), ("ex", C.c_float), ("ey", C.c_float), ("n1", C.c_int32), ("n2", C.c_int32),
                ("n1_nodes", C.c_int32), ("n2_nodes", C.c_int32), ("only_stereo", C.c_int32), ("check_orientation", C.c_int32)]


assert C.sizeof(KnnJob) == 32 and C.sizeof(BowJob) == 144 and C.sizeof(ProjJob) == 328 and C.sizeof(TriJob) == 240


def _jobs_to_device(jobs, device):
    import torch
    arr = (type(jobs[0]) * len(jobs))(*jobs)
    host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
    return host.to(device)


def DescriptorDistance(a, b):
    """ORBmatcher::DescriptorDistance (static, host scalar)."""
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return lib().plslam_descriptor_distance(_vp(a), _vp(b))


def knn2_host(query, train):
    """BFMatcher(NORM_HAMMING).knnMatch(k=2) on host arrays -> (nq, 4) int32 idx1, dist1, idx2, dist2."""
    query = np.ascontiguousarray(query, np.uint8).reshape(-1, 32)
    train = np.ascontiguousarray(train, np.uint8).reshape(-1, 32)
    out = np.empty((len(query), 4), np.int32)
    _check(lib().plslam_match_knn2_host(_vp(query), len(query), _vp(train), len(train), _vp(out)))
    return out


def knn2_batch_device(pairs, stream=None, jobs_dev=None):
    """pairs: list of (query CUDA uint8 [nq,32], train CUDA uint8 [nt,32], out CUDA int32 [nq,4]).  Async."""
    if jobs_dev is None:
        jobs = [KnnJob(q.data_ptr(), t.data_ptr(), o.data_ptr(), q.shape[0], t.shape[0]) for q, t, o in pairs]
        jobs_dev = _jobs_to_device(jobs, pairs[0][0].device)
    max_nq = max(int(q.shape[0]) for q, _, _ in pairs)
    _check(lib().plslam_match_knn2_batch_device(_vp(jobs_dev), len(pairs), max_nq, _stream_ptr(stream)))
    return jobs_dev


def knn2_jobs_device(jobs_dev, njobs, max_nq, stream=None):
    _check(lib().plslam_match_knn2_batch_device(_vp(jobs_dev), njobs, max_nq, _stream_ptr(stream)))


def bow_batch_device(jobs, max_n, device, stream=None):
    jd = _jobs_to_device(jobs, device)
    _check(lib().plslam_match_bow_batch_device(_vp(jd), len(jobs), int(max_n), _stream_ptr(stream)))
    return jd


def projection_batch_device(jobs, max_n1, max_n2, device, stream=None):
    jd = _jobs_to_device(jobs, device)
    _check(lib().plslam_match_projection_batch_device(_vp(jd), len(jobs), int(max_n1), int(max_n2), _stream_ptr(stream)))
    return jd


# ---------------------------------------------------------------------------
# batched front-end (Frame::ExtractORB + Frame::ExtractLSD + pair matching)
# ---------------------------------------------------------------------------
class FrontendIO(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("keypoints", "descriptors", "kp_counts", "keylines", "line_descriptors",
                                          "line_functions", "line_counts", "orb_matches", "line_matches")]


class Frontend:
    """plslam_frontend_*: ORB + LSD/LBD (+ frame-pair kNN) for a batch of frames in one call."""

    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7, max_lines=40, depth=1):
        L = lib()
        self._h = C.c_void_p()
        L.plslam_frontend_create_pipelined.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_float, C.c_int, C.c_int,
                                                       C.c_int, C.c_int, C.c_int]
        L.plslam_frontend_destroy.argtypes = [C.c_void_p]
        L.plslam_frontend_destroy.restype = None
        _check(L.plslam_frontend_create_pipelined(C.byref(self._h), nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST,
                                                  max_lines, depth))
        self.depth = depth
        a, b = C.c_int(), C.c_int()
        _check(L.plslam_frontend_capacities(self._h, C.byref(a), C.byref(b)))
        self.kp_capacity, self.line_capacity = a.value, b.value

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                lib().plslam_frontend_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def alloc(self, batch, device=None, pinned=False):
        """Output block for `batch` frames: torch tensors on `device`, or pinned/pageable host tensors."""
        import torch
        kw = dict(device=device) if device is not None else dict(pin_memory=pinned)
        kc, lc, npairs = self.kp_capacity, self.line_capacity, max(batch // 2, 1)
        return dict(keypoints=torch.empty((batch, kc, 7), dtype=torch.int32, **kw),
                    descriptors=torch.empty((batch, kc, 32), dtype=torch.uint8, **kw),
                    kp_counts=torch.empty((batch,), dtype=torch.int32, **kw),
                    keylines=torch.empty((batch, lc, 17), dtype=torch.int32, **kw),
                    line_descriptors=torch.empty((batch, lc, 32), dtype=torch.uint8, **kw),
                    line_functions=torch.empty((batch, lc, 3), dtype=torch.float64, **kw),
                    line_counts=torch.empty((batch,), dtype=torch.int32, **kw),
                    orb_matches=torch.empty((npairs, kc, 4), dtype=torch.int32, **kw),
                    line_matches=torch.empty((npairs, lc, 4), dtype=torch.int32, **kw))

    @staticmethod
    def _io(out):
        io = FrontendIO()
        for k, _ in FrontendIO._fields_:
            setattr(io, k, out[k].data_ptr())
        return io

    def process_device(self, d_images, out, match_pairs=True, stream=None):
        B, H, W = d_images.shape
        io = self._io(out)
        _check(lib().plslam_frontend_process_device(self._h, _vp(d_images), B, W, H, d_images.stride(1),
                                                    C.c_size_t(d_images.stride(0)), C.byref(io), int(match_pairs),
                                                    _stream_ptr(stream)))
        return out

    def _host_call(self, fn, images, out, match_pairs):
        B, H, W = images.shape
        io = self._io(out)
        ptr = images.data_ptr() if hasattr(images, "data_ptr") else images.ctypes.data
        st0 = images.stride(0) if hasattr(images, "stride") else images.strides[0]
        st1 = images.stride(1) if hasattr(images, "stride") else images.strides[1]
        _check(fn(self._h, C.c_void_p(ptr), B, W, H, st1, C.c_size_t(st0), C.byref(io), int(match_pairs)))
        return out

    def process_host(self, images, out, match_pairs=True):
        """images: (B, H, W) uint8 host tensor/array (pinned => asynchronous copies); out: host block from alloc()."""
        return self._host_call(lib().plslam_frontend_process_host, images, out, match_pairs)

    def submit_host(self, images, out, match_pairs=True):
        """Asynchronous process_host on the next pipeline slot; `out` is valid after wait_host()."""
        return self._host_call(lib().plslam_frontend_submit_host, images, out, match_pairs)

    def submit_host_wave(self, images_list, outs, match_pairs=True):
        """plslam_frontend_submit_host_wave: len(images_list) <= depth batches (pinned uint8 tensors [B, H, W]) enter the
        pipeline together; outs[i] receives batch i after wait_host()."""
        n = len(images_list)
        assert n == len(outs) and n >= 1
        B, H, W = images_list[0].shape
        st0, st1 = images_list[0].stride(0), images_list[0].stride(1)
        ptrs = (C.c_void_p * n)()
        ios = (FrontendIO * n)()
        for i, (im, out) in enumerate(zip(images_list, outs)):
            assert tuple(im.shape) == (B, H, W) and im.stride(0) == st0 and im.stride(1) == st1 and im.stride(2) == 1
            ptrs[i] = im.data_ptr()
            ios[i] = self._io(out)
        _check(lib().plslam_frontend_submit_host_wave(self._h, ptrs, n, B, W, H, st1, C.c_size_t(st0), ios, int(bool(match_pairs))))

    def acquire_slot(self):
        """Blocks until a pipeline slot is idle and returns its index (keep one set of host output buffers per slot)."""
        return lib().plslam_frontend_acquire_slot(self._h)

    def submit_host_slot(self, slot, images, out, match_pairs=True):
        """submit_host on the slot returned by acquire_slot()."""
        B, H, W = images.shape
        io = self._io(out)
        ptr = images.data_ptr() if hasattr(images, "data_ptr") else images.ctypes.data
        st0 = images.stride(0) if hasattr(images, "stride") else images.strides[0]
        st1 = images.stride(1) if hasattr(images, "stride") else images.strides[1]
        _check(lib().plslam_frontend_submit_host_slot(self._h, int(slot), C.c_void_p(ptr), B, W, H, st1, C.c_size_t(st0),
                                                      C.byref(io), int(match_pairs)))
        return out

    def wait_host(self):
        _check(lib().plslam_frontend_wait_host(self._h))

    def check_status(self, stream=None):
        _check(lib().plslam_frontend_check_status(self._h, _stream_ptr(stream)))

    def enable_timing(self, on=True):
        _check(lib().plslam_frontend_enable_timing(self._h, int(on)))

    def stage_times(self):
        names = (C.c_char_p * 64)()
        ms = (C.c_float * 64)()
        n = lib().plslam_frontend_stage_times(self._h, names, ms, 64)
        return [(names[i].decode(), float(ms[i])) for i in range(n)]

    def launches_per_call(self, match_pairs=True):
        return lib().plslam_frontend_launches_per_call(self._h, int(match_pairs))


# ---------------------------------------------------------------------------
# ORB vocabulary (DBoW2 transform as used by Frame::ComputeBoW)
# ---------------------------------------------------------------------------
class ORBVocabulary:
    """plslam_voc_*: loadFromTextFile + per-feature tree descent on the GPU; BowVector / FeatureVector assembly
    (std::map accumulation in feature order, L1 normalisation) on the host, as DBoW2's transform()."""

    def __init__(self, path=None, handle=None):
        L = lib()
        L.plslam_voc_load_text.argtypes = [C.POINTER(C.c_void_p), C.c_char_p]
        L.plslam_voc_destroy.argtypes = [C.c_void_p]
        L.plslam_voc_destroy.restype = None
        L.plslam_voc_blob_bytes.restype = C.c_size_t
        L.plslam_voc_blob_bytes.argtypes = [C.c_void_p]
        self._h = C.c_void_p()
        if handle is not None:
            self._h = handle
        else:
            _check(L.plslam_voc_load_text(C.byref(self._h), path.encode()))
        k, Lv, nn, nw = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _check(L.plslam_voc_info(self._h, C.byref(k), C.byref(Lv), C.byref(nn), C.byref(nw)))
        self.k, self.L, self.n_nodes, self.n_words = k.value, Lv.value, nn.value, nw.value

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                lib().plslam_voc_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    @classmethod
    def from_arrays(cls, k, L, parent, is_leaf, descriptors, weights):
        """plslam_voc_create: nodes in loadFromTextFile order (entry 0 = root)."""
        parent = np.ascontiguousarray(parent, np.int32); is_leaf = np.ascontiguousarray(is_leaf, np.uint8)
        descriptors = np.ascontiguousarray(descriptors, np.uint8); weights = np.ascontiguousarray(weights, np.float64)
        h = C.c_void_p()
        _check(lib().plslam_voc_create(C.byref(h), int(k), int(L), len(parent), _vp(parent), _vp(is_leaf), _vp(descriptors),
                                       _vp(weights)))
        return cls(handle=h)

    def export_blob(self):
        import torch
        n = lib().plslam_voc_blob_bytes(self._h)
        t = torch.empty(n, dtype=torch.uint8, device="cuda")
        _check(lib().plslam_voc_export_blob(self._h, _vp(t), _stream_ptr()))
        torch.cuda.synchronize()
        return t

    @classmethod
    def from_blob(cls, blob):
        h = C.c_void_p()
        _check(lib().plslam_voc_import_blob(C.byref(h), _vp(blob), C.c_size_t(blob.numel())))
        return cls(handle=h)

    def transform_features_device(self, d_desc, levelsup=4, stream=None):
        import torch
        n = d_desc.shape[0]
        word = torch.empty(n, dtype=torch.int32, device=d_desc.device)
        weight = torch.empty(n, dtype=torch.float64, device=d_desc.device)
        node = torch.empty(n, dtype=torch.int32, device=d_desc.device)
        _check(lib().plslam_voc_transform_device(self._h, _vp(d_desc), n, levelsup, _vp(word), _vp(weight), _vp(node),
                                                 _stream_ptr(stream)))
        return word, weight, node

    def featvec_batch_device(self, d_desc, d_counts, levelsup=4, out=None, stream=None):
        """Frame::ComputeBoW for a batch: d_desc [B][cap][32] uint8, d_counts [B] -> dict of device tensors word, weight,
        node [B][cap], fv_nodes [B][cap], fv_start [B][cap+1], fv_idx [B][cap], fv_count [B]."""
        import torch
        B, cap = d_desc.shape[0], d_desc.shape[1]
        dev = d_desc.device
        if out is None:
            i32 = lambda *s: torch.empty(s, dtype=torch.int32, device=dev)
            out = dict(word=i32(B, cap), weight=torch.empty((B, cap), dtype=torch.float64, device=dev), node=i32(B, cap),
                       fv_nodes=i32(B, cap), fv_start=i32(B, cap + 1), fv_idx=i32(B, cap), fv_count=i32(B))
        _check(lib().plslam_voc_featvec_batch_device(self._h, _vp(d_desc), _vp(d_counts), B, cap, int(levelsup), _vp(out["word"]),
                                                     _vp(out["weight"]), _vp(out["node"]), _vp(out["fv_nodes"]),
                                                     _vp(out["fv_start"]), _vp(out["fv_idx"]), _vp(out["fv_count"]),
                                                     _stream_ptr(stream)))
        return out

    def transform_features(self, desc, levelsup=4):
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        word = np.empty(n, np.int32); weight = np.empty(n, np.float64); node = np.empty(n, np.int32)
        _check(lib().plslam_voc_transform_host(self._h, _vp(desc), n, levelsup, _vp(word), _vp(weight), _vp(node)))
        return word, weight, node

    def transform(self, desc, levelsup=4):
        """Frame::ComputeBoW of one frame -> dict(bow_ids, bow_vals, fv_nodes, fv_start, fv_idx): BowVector and FeatureVector
        in std::map order, assembled on the device (plslam_voc_compute_bow_host)."""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        ids = np.empty(max(n, 1), np.int32); vals = np.empty(max(n, 1), np.float64)
        fn = np.empty(max(n, 1), np.int32); fs = np.zeros(n + 1, np.int32); fi = np.empty(max(n, 1), np.int32)
        nb, nf = C.c_int(0), C.c_int(0)
        _check(lib().plslam_voc_compute_bow_host(self._h, _vp(desc), n, int(levelsup), _vp(ids), _vp(vals), C.byref(nb), _vp(fn),
                                                 _vp(fs), _vp(fi), C.byref(nf)))
        nb, nf = nb.value, nf.value
        return dict(bow_ids=ids[:nb].astype(np.uint32), bow_vals=vals[:nb].copy(), fv_nodes=fn[:nf].astype(np.uint32),
                    fv_start=fs[:nf + 1].copy(), fv_idx=fi[:fs[nf]].astype(np.uint32))

    def bowvec_batch_device(self, fv, d_counts, out=None, stream=None):
        """mBowVec of every frame of a batch from the outputs of featvec_batch_device -> dict(bow_ids, bow_vals [B][cap],
        bow_count [B]) on the device."""
        import torch
        B, cap = fv["word"].shape
        dev = fv["word"].device
        if out is None:
            out = dict(bow_ids=torch.empty((B, cap), dtype=torch.int32, device=dev),
                       bow_vals=torch.empty((B, cap), dtype=torch.float64, device=dev),
                       bow_count=torch.empty((B,), dtype=torch.int32, device=dev))
        _check(lib().plslam_voc_bowvec_batch_device(_vp(fv["word"]), _vp(fv["weight"]), _vp(d_counts), B, cap, _vp(out["bow_ids"]),
                                                    _vp(out["bow_vals"]), _vp(out["bow_count"]), _stream_ptr(stream)))
        return out


def assemble_bow(word, weight, node):
    """Host restatement of the assembly (BowVector::addWeight in feature order + normalize(L1); FeatureVector::addFeature,
    TemplatedVocabulary.h:1151-1217): test helper only — ORBVocabulary.transform() runs the device kernels."""
    bow, fv = {}, {}
    for i in range(len(word)):
        w = float(weight[i])
        if w > 0:
            bow[int(word[i])] = bow.get(int(word[i]), 0.0) + w if int(word[i]) in bow else w
            fv.setdefault(int(node[i]), []).append(i)
    ids = sorted(bow)
    vals = [bow[i] for i in ids]
    norm = 0.0
    for v in vals:
        norm += abs(v)
    if norm > 0.0:
        vals = [v / norm for v in vals]
    nodes = sorted(fv)
    start = [0]
    idx = []
    for nd in nodes:
        idx.extend(fv[nd])
        start.append(len(idx))
    return dict(bow_ids=np.array(ids, np.uint32), bow_vals=np.array(vals, np.float64), fv_nodes=np.array(nodes, np.uint32),
                fv_start=np.array(start, np.int32), fv_idx=np.array(idx, np.uint32))


# ------------------------------------------------------------------------------------------------
# Frame post-extraction steps (reference include/Frame.h:110-120,266-273): undistortion, RGB-D stereo
# coordinates, grid assignment
# ------------------------------------------------------------------------------------------------
class FrameCalib(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2", "k3", "bf")]

    @classmethod
    def from_dict(cls, d):
        return cls(*[float(d.get(n, 0.0)) for n, _ in cls._fields_])


TUM1_CALIB = dict(fx=517.306408, fy=516.469215, cx=318.643040, cy=255.313989, k1=0.262383, k2=-0.953104,
                  p1=-0.005358, p2=0.002628, k3=1.163314, bf=40.0)  # reference Examples/RGB-D/TUM1.yaml:8-26


def frame_image_bounds(calib, cols, rows):
    """Frame::ComputeImageBounds -> float32[4] mnMinX, mnMaxX, mnMinY, mnMaxY."""
    c = calib if isinstance(calib, FrameCalib) else FrameCalib.from_dict(calib)
    b = np.empty(4, np.float32)
    _check(lib().plslam_frame_image_bounds(C.byref(c), int(cols), int(rows), _vp(b)))
    return b


def frame_post_device(calib, bounds, d_keypoints, d_counts, d_depth, out=None, stream=None):
    """Batched UndistortKeyPoints + ComputeStereoFromRGBD + AssignFeaturesToGrid on device tensors.
    d_keypoints [B][cap][7] int32 view of cv::KeyPoint, d_counts [B] int32, d_depth [B or 1][H][W] float32."""
    import torch
    c = calib if isinstance(calib, FrameCalib) else FrameCalib.from_dict(calib)
    B, cap = d_keypoints.shape[0], d_keypoints.shape[1]
    assert d_depth.dtype == torch.float32 and d_depth.dim() == 3 and d_depth.stride(2) == 1
    H, W = d_depth.shape[1], d_depth.shape[2]
    dev = d_keypoints.device
    if out is None:
        out = dict(un_xy=torch.empty((B, cap, 2), dtype=torch.float32, device=dev),
                   uright=torch.empty((B, cap), dtype=torch.float32, device=dev),
                   depth=torch.empty((B, cap), dtype=torch.float32, device=dev),
                   grid_start=torch.empty((B, 64 * 48 + 1), dtype=torch.int32, device=dev),
                   grid_items=torch.empty((B, cap), dtype=torch.int32, device=dev))
    b = np.ascontiguousarray(bounds, np.float32)
    stride = d_depth.stride(0) if d_depth.shape[0] > 1 else 0
    _check(lib().plslam_frame_post_batch_device(C.byref(c), _vp(b), _vp(d_keypoints), _vp(d_counts), B, cap, _vp(d_depth),
                                                W, H, d_depth.stride(1), C.c_size_t(stride), _vp(out["un_xy"]),
                                                _vp(out["uright"]), _vp(out["depth"]), _vp(out["grid_start"]),
                                                _vp(out["grid_items"]), _stream_ptr(stream)))
    return out


def frame_post_host(calib, bounds, keypoints, depth):
    """One frame, numpy in / numpy out (KP_DTYPE keypoints, float32 depth map)."""
    c = calib if isinstance(calib, FrameCalib) else FrameCalib.from_dict(calib)
    kps = np.ascontiguousarray(keypoints, KP_DTYPE)
    depth = np.ascontiguousarray(depth, np.float32)
    n = len(kps)
    un, ur, z = np.empty((n, 2), np.float32), np.empty(n, np.float32), np.empty(n, np.float32)
    gs, gi = np.empty(64 * 48 + 1, np.int32), np.empty(max(n, 1), np.int32)
    b = np.ascontiguousarray(bounds, np.float32)
    _check(lib().plslam_frame_post_host(C.byref(c), _vp(b), _vp(kps), n, _vp(depth), depth.shape[1], depth.shape[0],
                                        depth.strides[0] // 4, _vp(un), _vp(ur), _vp(z), _vp(gs), _vp(gi)))
    return dict(un_xy=un, uright=ur, depth=z, grid_start=gs, grid_items=gi[:gs[-1]])


def search_by_projection_host(last, cur, cam, scale_factors, tcw_cur, tcw_last, th, mono=False, check_ori=True,
                              report_removed=False):
    """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) on host arrays (dict layout of tests/matchdata.py)
    through plslam_match_projection_host -> (match_cur, nmatches)."""
    keep = {k: np.ascontiguousarray(v) for k, v in list(last.items()) + [("c_" + k, v) for k, v in cur.items()]}
    sf = np.ascontiguousarray(scale_factors, np.float32)
    n1, n2 = len(last["desc"]), len(cur["desc"])
    m, n = np.empty(max(n2, 1), np.int32), np.zeros(1, np.int32)
    p = lambda a: a.ctypes.data
    j = ProjJob(p(keep["valid"]), p(keep["xyz"]), p(keep["desc"]), p(keep["octave"]), p(keep["angle"]), p(keep["obs"]),
                p(keep["c_xy"]), p(keep["c_octave"]), p(keep["c_angle"]), p(keep["c_desc"]), p(keep["c_uright"]),
                p(keep["c_taken"]), p(keep["c_grid_start"]), p(keep["c_grid_items"]), p(sf), p(m), p(n))
    j.cam[:] = np.asarray(cam, np.float32).tolist()
    j.tcw_cur[:] = np.asarray(tcw_cur, np.float32).ravel().tolist()
    j.tcw_last[:] = np.asarray(tcw_last, np.float32).ravel().tolist()
    j.th = float(th); j.n1 = n1; j.n2 = n2; j.mono = int(mono); j.check_orientation = int(check_ori)
    j.report_removed = int(report_removed)
    _check(lib().plslam_match_projection_host(C.byref(j), len(sf)))
    return m[:n2], int(n[0])


def search_by_projection_kf_host(kf, cur, cam, scale_factors, log_scale_factor, tcw_cur, th, orb_dist, check_ori=True):
    """ORBmatcher::SearchByProjection(Frame&, KeyFrame*, set<MapPoint*>&, th, ORBdist) (relocalisation) on host arrays (layout of
    tests/matchdata.py: relocalisation_case; cam = fx, fy, cx, cy, mnMinX, mnMaxX, mnMinY, mnMaxY, gridWInv, gridHInv) through
    plslam_match_projection_host with mode 1 -> (match_cur int32 [N2], nmatches)."""
    m, n2 = len(kf["desc"]), len(cur["desc"])
    keep = []
    def a(x, dt):
        x = np.ascontiguousarray(x, dt); keep.append(x); return x.ctypes.data
    match = np.empty(max(n2, 1), np.int32); nm = np.zeros(1, np.int32)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    j = ProjJob()
    j.last_valid, j.last_xyz, j.last_desc = a(kf["valid"], np.uint8), a(kf["xyz"], np.float32), a(kf["desc"], np.uint8)
    j.last_angle, j.last_dist_range = a(kf["angle"], np.float32), a(kf["dist_range"], np.float32)
    j.cur_xy, j.cur_octave, j.cur_angle = a(cur["xy"], np.float32), a(cur["octave"], np.int32), a(cur["angle"], np.float32)
    j.cur_desc, j.cur_taken = a(cur["desc"], np.uint8), a(cur["taken"], np.uint8)
    j.grid_start, j.grid_items, j.scale_factors = a(cur["grid_start"], np.int32), a(cur["grid_items"], np.int32), sf.ctypes.data
    j.match_cur, j.nmatches = match.ctypes.data, nm.ctypes.data
    c = np.asarray(cam, np.float32)
    j.cam = (C.c_float * 12)(c[0], c[1], c[2], c[3], 0.0, 0.0, c[4], c[5], c[6], c[7], c[8], c[9])
    j.tcw_cur = (C.c_float * 12)(*np.asarray(tcw_cur, np.float32).reshape(12))
    j.th, j.n1, j.n2, j.mono, j.check_orientation, j.report_removed = float(th), m, n2, 0, int(check_ori), 0
    j.mode, j.orb_dist, j.n_levels, j.log_scale_factor = 1, int(orb_dist), len(sf), float(log_scale_factor)
    _check(lib().plslam_match_projection_host(C.byref(j), len(sf)))
    return match[:n2], int(nm[0])


class KfProjJob(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("mp_valid", "mp_xyz", "mp_normal", "mp_dist_range", "mp_desc", "mp_level", "kf_xy", "kf_octave",
                                          "kf_desc", "kf_matched", "grid_start", "grid_items", "scale_factors", "match_kf", "nmatches")] + \
               [("scw", C.c_float * 12), ("cam", C.c_float * 4), ("bounds", C.c_int32 * 4), ("grid_width_inv", C.c_float),
                ("grid_height_inv", C.c_float), ("log_scale_factor", C.c_float), ("grid_cols", C.c_int32), ("grid_rows", C.c_int32),
                ("n_levels", C.c_int32), ("th", C.c_int32), ("m", C.c_int32), ("n", C.c_int32)]


def search_by_projection_sim3_host(kf, mp, scw, matched_in, th):
    """ORBmatcher::SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) on host arrays (layout: tests/matchdata.py
    loop_projection_case) through plslam_match_kf_projection_host -> (match_kf int32 [N]: map-point index newly assigned to each
    key-frame feature or -1, nmatches)."""
    m, n = len(mp["desc"]), len(kf["desc"])
    matched_in = np.asarray(matched_in, np.int32)
    found = np.zeros(m, bool)
    found[matched_in[matched_in >= 0]] = True
    keep = []
    def a(x, dt):
        x = np.ascontiguousarray(x, dt); keep.append(x); return x.ctypes.data
    sf = np.ascontiguousarray(kf["scale_factors"], np.float32)
    out = np.empty(max(n, 1), np.int32); nm = np.zeros(1, np.int32)
    j = KfProjJob()
    j.mp_valid = a((np.asarray(mp["state"]) == 1) & ~found, np.uint8)
    j.mp_xyz, j.mp_normal, j.mp_dist_range = a(mp["xyz"], np.float32), a(mp["normal"], np.float32), a(mp["dist_range"], np.float32)
    j.mp_desc = a(mp["desc"], np.uint8)
    j.kf_xy, j.kf_octave, j.kf_desc = a(kf["xy"], np.float32), a(kf["octave"], np.int32), a(kf["desc"], np.uint8)
    j.kf_matched = a(matched_in >= 0, np.uint8)
    j.grid_start, j.grid_items, j.scale_factors = a(kf["grid_start"], np.int32), a(kf["grid_items"], np.int32), sf.ctypes.data
    j.match_kf, j.nmatches = out.ctypes.data, nm.ctypes.data
    j.scw = (C.c_float * 12)(*np.asarray(scw, np.float32).reshape(12))
    j.cam = (C.c_float * 4)(*np.asarray(kf["cam4"], np.float32))
    j.bounds = (C.c_int32 * 4)(*[int(v) for v in kf["bounds4"]])
    j.grid_width_inv, j.grid_height_inv, j.log_scale_factor = float(kf["gwi"]), float(kf["ghi"]), float(kf["log_sf"])
    j.grid_cols, j.grid_rows, j.n_levels, j.th, j.m, j.n = 64, 48, len(sf), int(th), m, n
    _check(lib().plslam_match_kf_projection_host(C.byref(j)))
    return out[:n], int(nm[0])


class FuseJob(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("mp_valid", "mp_xyz", "mp_normal", "mp_dist_range", "mp_desc", "mp_level", "kf_xy", "kf_octave",
                                          "kf_uright", "kf_desc", "grid_start", "grid_items", "scale_factors", "inv_level_sigma2", "best_idx")] + \
               [("pose", C.c_float * 12), ("pose2", C.c_float * 12), ("ow", C.c_float * 3), ("cam", C.c_float * 5), ("bounds", C.c_int32 * 4), ("grid_width_inv", C.c_float),
                ("grid_height_inv", C.c_float), ("log_scale_factor", C.c_float), ("th", C.c_float), ("grid_cols", C.c_int32),
                ("grid_rows", C.c_int32), ("n_levels", C.c_int32), ("use_scw", C.c_int32), ("m", C.c_int32), ("n", C.c_int32)]


def fuse_search_host(kf, mp, th, scw=None):
    """Matching core of ORBmatcher::Fuse on host arrays (layout: tests/matchdata.py fuse_case): the rigid form (pKF's pose, chi-square
    tests) or, with scw, the similarity form -> best_idx int32 [M] (key-frame feature each map point would be fused into, -1 = none)."""
    m, n = len(mp["desc"]), len(kf["desc"])
    keep = []
    def a(x, dt):
        x = np.ascontiguousarray(x, dt); keep.append(x); return x.ctypes.data
    sf = np.ascontiguousarray(kf["scale_factors"], np.float32)
    out = np.empty(max(m, 1), np.int32)
    j = FuseJob()
    j.mp_valid = a(np.asarray(mp["state"]) == 1, np.uint8)
    j.mp_xyz, j.mp_normal, j.mp_dist_range = a(mp["xyz"], np.float32), a(mp["normal"], np.float32), a(mp["dist_range"], np.float32)
    j.mp_desc = a(mp["desc"], np.uint8)
    j.kf_xy, j.kf_octave, j.kf_desc = a(kf["xy"], np.float32), a(kf["octave"], np.int32), a(kf["desc"], np.uint8)
    j.kf_uright, j.inv_level_sigma2 = a(kf["uright"], np.float32), a(kf["inv_level_sigma2"], np.float32)
    j.grid_start, j.grid_items, j.scale_factors = a(kf["grid_start"], np.int32), a(kf["grid_items"], np.int32), sf.ctypes.data
    j.best_idx = out.ctypes.data
    j.pose = (C.c_float * 12)(*np.asarray(kf["tcw"] if scw is None else scw, np.float32).reshape(12))
    j.ow = (C.c_float * 3)(*np.asarray(kf["ow"], np.float32).reshape(3))
    j.cam = (C.c_float * 5)(*(list(np.asarray(kf["cam4"], np.float32)) + [float(kf["mbf"])]))
    j.bounds = (C.c_int32 * 4)(*[int(v) for v in kf["bounds4"]])
    j.grid_width_inv, j.grid_height_inv, j.log_scale_factor, j.th = float(kf["gwi"]), float(kf["ghi"]), float(kf["log_sf"]), float(th)
    j.grid_cols, j.grid_rows, j.n_levels, j.use_scw, j.m, j.n = 64, 48, len(sf), 0 if scw is None else 1, m, n
    _check(lib().plslam_match_fuse_search_host(C.byref(j)))
    return out[:m]


def sim3_transforms(s12, R12, t12):
    """sR12, sR21, t21 of ORBmatcher::SearchBySim3 with the reference's arithmetic (plslam_sim3_transforms)."""
    R = np.ascontiguousarray(R12, np.float32).reshape(9); t = np.ascontiguousarray(t12, np.float32).reshape(3)
    sR12, sR21, t21 = np.empty(9, np.float32), np.empty(9, np.float32), np.empty(3, np.float32)
    f = lib().plslam_sim3_transforms
    f.argtypes, f.restype = [C.c_float] + [C.c_void_p] * 5, None
    f(float(s12), _vp(R), _vp(t), _vp(sR12), _vp(sR21), _vp(t21))
    return sR12.reshape(3, 3), sR21.reshape(3, 3), t21


def search_by_sim3_host(kf1, kf2, mp1, mp2, s12, R12, t12, th, matched_in):
    """ORBmatcher::SearchBySim3 on host arrays (layout: tests/matchdata.py sim3_case): the two directions through
    plslam_match_fuse_search_host (use_scw = 2), the agreement pass here -> (match12 int32 [N1], nFound)."""
    n1, n2 = len(kf1["desc"]), len(kf2["desc"])
    matched_in = np.asarray(matched_in, np.int32)
    already1 = matched_in >= 0
    already2 = np.zeros(n2, bool)
    already2[matched_in[already1]] = True
    sR12, sR21, t21 = sim3_transforms(s12, R12, t12)
    t12 = np.asarray(t12, np.float32).reshape(3)
    def direction(kf_own, mp, already, kf_other, sR, tt):
        m, n = len(mp["desc"]), len(kf_other["desc"])
        keep = []
        def a(x, dt):
            x = np.ascontiguousarray(x, dt); keep.append(x); return x.ctypes.data
        sf = np.ascontiguousarray(kf_other["scale_factors"], np.float32)
        out = np.empty(max(m, 1), np.int32)
        j = FuseJob()
        j.mp_valid = a((np.asarray(mp["state"]) == 1) & ~already, np.uint8)
        j.mp_xyz, j.mp_dist_range, j.mp_desc = a(mp["xyz"], np.float32), a(mp["dist_range"], np.float32), a(mp["desc"], np.uint8)
        j.kf_xy, j.kf_octave, j.kf_desc = a(kf_other["xy"], np.float32), a(kf_other["octave"], np.int32), a(kf_other["desc"], np.uint8)
        j.grid_start, j.grid_items, j.scale_factors = a(kf_other["grid_start"], np.int32), a(kf_other["grid_items"], np.int32), sf.ctypes.data
        j.best_idx = out.ctypes.data
        j.pose = (C.c_float * 12)(*np.asarray(kf_own["tcw"], np.float32).reshape(12))
        j.pose2 = (C.c_float * 12)(*np.hstack([sR, np.asarray(tt, np.float32).reshape(3, 1)]).astype(np.float32).reshape(12))
        j.cam = (C.c_float * 5)(*(list(np.asarray(kf_other["cam4"], np.float32)) + [0.0]))
        j.bounds = (C.c_int32 * 4)(*[int(v) for v in kf_other["bounds4"]])
        j.grid_width_inv, j.grid_height_inv = float(kf_other["gwi"]), float(kf_other["ghi"])
        j.log_scale_factor, j.th = float(kf_other["log_sf"]), float(th)
        j.grid_cols, j.grid_rows, j.n_levels, j.use_scw, j.m, j.n = 64, 48, len(sf), 2, m, n
        _check(lib().plslam_match_fuse_search_host(C.byref(j)))
        return out[:m]
    m1 = direction(kf1, mp1, already1, kf2, sR21, t21)
    m2 = direction(kf2, mp2, already2, kf1, sR12, t12)
    match12 = np.full(n1, -1, np.int32)
    ok = (m1 >= 0)
    ok[ok] = m2[m1[ok]] == np.flatnonzero(ok)
    match12[ok] = m1[ok]
    return match12, int(ok.sum())


class FrustumJob(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("mp_xyz", "mp_normal", "mp_dist_range", "in_view", "proj", "level", "viewcos")] + \
               [("cam", C.c_float * 8), ("tcw", C.c_float * 12), ("ow", C.c_float * 3), ("mbf", C.c_float), ("log_scale_factor", C.c_float),
                ("viewing_cos_limit", C.c_float), ("n_levels", C.c_int32), ("m", C.c_int32)]


def is_in_frustum_host(xyz, normal, dist_range, cam8, tcw, ow, mbf, log_scale_factor, n_levels, cos_limit):
    """Frame::isInFrustum over M map points (host arrays) through plslam_frame_is_in_frustum_host ->
    dict(in_view uint8 [M], proj float32 [M,3], level int32 [M], viewcos float32 [M])."""
    m = len(xyz)
    x = np.ascontiguousarray(xyz, np.float32); nv = np.ascontiguousarray(normal, np.float32); dr = np.ascontiguousarray(dist_range, np.float32)
    out = dict(in_view=np.zeros(max(m, 1), np.uint8), proj=np.zeros((max(m, 1), 3), np.float32), level=np.zeros(max(m, 1), np.int32),
               viewcos=np.zeros(max(m, 1), np.float32))
    j = FrustumJob(x.ctypes.data, nv.ctypes.data, dr.ctypes.data, out["in_view"].ctypes.data, out["proj"].ctypes.data,
                   out["level"].ctypes.data, out["viewcos"].ctypes.data)
    j.cam = (C.c_float * 8)(*np.asarray(cam8, np.float32))
    j.tcw = (C.c_float * 12)(*np.asarray(tcw, np.float32).reshape(12))
    j.ow = (C.c_float * 3)(*np.asarray(ow, np.float32).reshape(3))
    j.mbf, j.log_scale_factor, j.viewing_cos_limit, j.n_levels, j.m = float(mbf), float(log_scale_factor), float(cos_limit), int(n_levels), m
    _check(lib().plslam_frame_is_in_frustum_host(C.byref(j)))
    return {k: v[:m] for k, v in out.items()}


class LineFrustumJob(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("ml_sp_ep", "ml_normal", "ml_dist_range", "in_view", "proj", "level", "viewcos")] + \
               [("cam", C.c_float * 8), ("tcw", C.c_float * 12), ("ow", C.c_float * 3), ("mbf", C.c_float), ("log_scale_factor", C.c_float),
                ("viewing_cos_limit", C.c_float), ("n_levels", C.c_int32), ("m", C.c_int32)]


def undistort_keylines_host(calib, xy4):
    """Frame::UndistortKeyLines (unpinned definition, include/plslam_b200.h): n x 4 end points -> n x 4."""
    xy4 = np.ascontiguousarray(xy4, np.float32).reshape(-1, 4)
    out = np.empty_like(xy4)
    c = calib if isinstance(calib, FrameCalib) else FrameCalib.from_dict(calib)
    _check(lib().plslam_frame_undistort_keylines_host(C.byref(c), _vp(xy4), len(xy4), _vp(out)))
    return out


def lines_in_area_host(queries7, lines4):
    """Frame::GetLinesInArea for a batch of queries (unpinned definition) -> list of int32 index arrays."""
    q = np.ascontiguousarray(queries7, np.float32).reshape(-1, 7)
    l4 = np.ascontiguousarray(lines4, np.float32).reshape(-1, 4)
    start = np.zeros(len(q) + 1, np.int32)
    items = np.empty(max(len(q) * max(len(l4), 1), 1), np.int32)
    _check(lib().plslam_frame_lines_in_area_host(_vp(q), len(q), _vp(l4), len(l4), _vp(start), _vp(items), len(items)))
    return [items[start[i]:start[i + 1]].copy() for i in range(len(q))]


def line_in_frustum_host(sp_ep, normal, dist_range, cam8, tcw, ow, mbf, log_scale_factor, n_levels, cos_limit):
    """Frame::isInFrustum(MapLine*, float) over M map lines (unpinned definition) -> dict(in_view, proj [M,6], level, viewcos)."""
    m = len(sp_ep)
    x = np.ascontiguousarray(sp_ep, np.float32); nv = np.ascontiguousarray(normal, np.float32); dr = np.ascontiguousarray(dist_range, np.float32)
    out = dict(in_view=np.zeros(max(m, 1), np.uint8), proj=np.zeros((max(m, 1), 6), np.float32), level=np.zeros(max(m, 1), np.int32),
               viewcos=np.zeros(max(m, 1), np.float32))
    j = LineFrustumJob(x.ctypes.data, nv.ctypes.data, dr.ctypes.data, out["in_view"].ctypes.data, out["proj"].ctypes.data,
                       out["level"].ctypes.data, out["viewcos"].ctypes.data)
    j.cam = (C.c_float * 8)(*np.asarray(cam8, np.float32))
    j.tcw = (C.c_float * 12)(*np.asarray(tcw, np.float32).reshape(12))
    j.ow = (C.c_float * 3)(*np.asarray(ow, np.float32).reshape(3))
    j.mbf, j.log_scale_factor, j.viewing_cos_limit, j.n_levels, j.m = float(mbf), float(log_scale_factor), float(cos_limit), int(n_levels), m
    _check(lib().plslam_frame_line_in_frustum_host(C.byref(j)))
    return {k: v[:m] for k, v in out.items()}


def predict_scale(max_distance, dist, log_scale_factor, n_levels):
    """MapPoint::PredictScale as the matcher kernels evaluate it."""
    f = lib().plslam_predict_scale
    f.argtypes, f.restype = [C.c_float, C.c_float, C.c_float, C.c_int], C.c_int
    return f(max_distance, dist, log_scale_factor, n_levels)


def bow_pairs_device(d_kps, d_desc, d_counts, fv, d_kf_valid=None, nnratio=0.7, check_ori=True, out=None, stream=None):
    """ORBmatcher::SearchByBoW on the frame pairs (2p, 2p+1) of a batch, everything device-resident.
    fv = ORBVocabulary.featvec_batch_device(...) -> dict(match [B/2][cap] int32, nmatches [B/2] int32)."""
    import torch
    B, cap = d_desc.shape[0], d_desc.shape[1]
    npairs = B // 2
    dev = d_desc.device
    if d_kf_valid is None:  # every keyframe feature has a good map point; reuse the buffer of a previous call
        d_kf_valid = out["_valid"] if out is not None else torch.ones((B, cap), dtype=torch.uint8, device=dev)
    if out is None:
        out = dict(match=torch.empty((npairs, cap), dtype=torch.int32, device=dev),
                   nmatches=torch.empty((npairs,), dtype=torch.int32, device=dev),
                   _angle=torch.empty((B, cap), dtype=torch.float32, device=dev),
                   _jobs=torch.empty((npairs, C.sizeof(BowJob)), dtype=torch.uint8, device=dev), _valid=d_kf_valid)
    _check(lib().plslam_match_bow_pairs_device(_vp(d_kps), _vp(d_desc), _vp(d_counts), cap, npairs, _vp(fv["fv_nodes"]),
                                               _vp(fv["fv_start"]), _vp(fv["fv_idx"]), _vp(fv["fv_count"]), _vp(d_kf_valid),
                                               C.c_float(nnratio), int(check_ori), _vp(out["_angle"]), _vp(out["_jobs"]),
                                               _vp(out["match"]), _vp(out["nmatches"]), _stream_ptr(stream)))
    return out


def search_by_bow_kfkf_host(kf1, kf2, nnratio=0.75, check_ori=True):
    """ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) on host arrays (dicts desc, angle, valid, nodes, start, idx)
    through plslam_match_bow_kfkf_host -> (match12 int32 [N1], nmatches)."""
    n1, n2 = len(kf1["desc"]), len(kf2["desc"])
    keep = []
    def a(x, dt):
        x = np.ascontiguousarray(x, dt); keep.append(x); return x.ctypes.data
    mf = np.empty(max(n2, 1), np.int32); m12 = np.empty(max(n1, 1), np.int32); nm = np.zeros(1, np.int32)
    j = BowJob(a(kf1["desc"], np.uint8), a(kf1["angle"], np.float32), a(kf1["valid"], np.uint8), a(kf1["nodes"], np.int32),
               a(kf1["start"], np.int32), a(kf1["idx"], np.int32), a(kf2["desc"], np.uint8), a(kf2["angle"], np.float32),
               a(kf2["nodes"], np.int32), a(kf2["start"], np.int32), a(kf2["idx"], np.int32), mf.ctypes.data, nm.ctypes.data,
               n1, n2, len(kf1["nodes"]), len(kf2["nodes"]), float(nnratio), int(check_ori), a(kf2["valid"], np.uint8), 1)
    _check(lib().plslam_match_bow_kfkf_host(C.byref(j), _vp(m12)))
    return m12[:n1], int(nm[0])


class LocalJob(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("mp_valid", "mp_proj", "mp_level", "mp_viewcos", "mp_desc", "mp_obs", "f_xy", "f_octave",
                                          "f_desc", "f_uright", "f_taken", "grid_start", "grid_items", "scale_factors", "match_f",
                                          "nmatches")] + \
               [("cam", C.c_float * 4), ("th", C.c_float), ("nnratio", C.c_float), ("m", C.c_int32), ("n", C.c_int32)]


assert C.sizeof(LocalJob) == 160


def local_points_batch_device(jobs, max_n, device, stream=None):
    """ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th) for a list of LocalJob (device pointers)."""
    jd = _jobs_to_device(jobs, device)
    _check(lib().plslam_match_local_points_batch_device(_vp(jd), len(jobs), int(max_n), _stream_ptr(stream)))
    return jd


def search_local_points_host(mp, fr, cam4, scale_factors, th, nnratio=0.8):
    """Host-array form (dict layout of tests/matchdata.local_points_case) -> (match_f, nmatches)."""
    keep = {k: np.ascontiguousarray(v) for k, v in list(mp.items()) + [("f_" + k, v) for k, v in fr.items()]}
    sf = np.ascontiguousarray(scale_factors, np.float32)
    m, n = len(mp["desc"]), len(fr["desc"])
    out, cnt = np.empty(max(n, 1), np.int32), np.zeros(1, np.int32)
    p = lambda a: a.ctypes.data
    j = LocalJob(p(keep["valid"]), p(keep["proj"]), p(keep["level"]), p(keep["viewcos"]), p(keep["desc"]), p(keep["obs"]),
                 p(keep["f_xy"]), p(keep["f_octave"]), p(keep["f_desc"]), p(keep["f_uright"]), p(keep["f_taken"]),
                 p(keep["f_grid_start"]), p(keep["f_grid_items"]), p(sf), p(out), p(cnt))
    j.cam[:] = np.asarray(cam4, np.float32).tolist()
    j.th = float(th); j.nnratio = float(nnratio); j.m = m; j.n = n
    _check(lib().plslam_match_local_points_host(C.byref(j), len(sf)))
    return out[:n], int(cnt[0])


# ---- on-disk formats either side of the path (host only; include/plslam_b200.h, last section) ----
TUM_NAME_STRIDE = 256


class InitJob(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("f1_octave", "f1_angle", "f1_desc", "f2_xy", "f2_angle", "f2_octave", "f2_desc", "grid_start",
                                          "grid_items", "prev_matched", "match12", "nmatches")] + \
               [("cam", C.c_float * 4), ("nnratio", C.c_float), ("window_size", C.c_int32), ("n1", C.c_int32), ("n2", C.c_int32),
                ("check_orientation", C.c_int32)]


def search_for_initialization_host(f1, f2, cam4, prev_matched, window_size=100, nnratio=0.9, check_ori=True):
    """ORBmatcher::SearchForInitialization on host arrays (dict layout of tests/matchdata.py) through
    plslam_match_initialization_host -> (vnMatches12, nmatches, updated vbPrevMatched)."""
    keep = {k: np.ascontiguousarray(v) for k, v in list(f1.items()) + [("2_" + k, v) for k, v in f2.items()]}
    n1, n2 = len(f1["desc"]), len(f2["desc"])
    prev = np.ascontiguousarray(prev_matched, np.float32).copy()
    m, n = np.empty(max(n1, 1), np.int32), np.zeros(1, np.int32)
    p = lambda a: a.ctypes.data
    j = InitJob(p(keep["octave"]), p(keep["angle"]), p(keep["desc"]), p(keep["2_xy"]), p(keep["2_angle"]), p(keep["2_octave"]), p(keep["2_desc"]),
                p(keep["2_grid_start"]), p(keep["2_grid_items"]), p(prev), p(m), p(n))
    j.cam[:] = np.asarray(cam4, np.float32).tolist()
    j.nnratio = float(nnratio); j.window_size = int(window_size); j.n1 = n1; j.n2 = n2; j.check_orientation = int(check_ori)
    _check(lib().plslam_match_initialization_host(C.byref(j)))
    return m[:n1], int(n[0]), prev


def LoadImages(association_file):
    """Examples/RGB-D/rgbd_tum.cc:151 LoadImages -> (vstrImageFilenamesRGB, vstrImageFilenamesD, vTimestamps)."""
    path = os.fsencode(association_file)
    n = C.c_int(0)
    _check(lib().plslam_tum_load_associations(path, None, None, None, TUM_NAME_STRIDE, 0, C.byref(n)))
    n = n.value
    ts = np.empty(n, np.float64)
    rgb, dep = C.create_string_buffer(max(n, 1) * TUM_NAME_STRIDE), C.create_string_buffer(max(n, 1) * TUM_NAME_STRIDE)
    m = C.c_int(0)
    _check(lib().plslam_tum_load_associations(path, _vp(ts), rgb, dep, TUM_NAME_STRIDE, n, C.byref(m)))
    cut = lambda buf: [buf[i * TUM_NAME_STRIDE:(i + 1) * TUM_NAME_STRIDE].split(b"\0", 1)[0].decode("latin-1") for i in range(n)]
    return cut(rgb.raw), cut(dep.raw), ts


def trajectory_line(timestamp, Tcw):
    """One line of System::SaveTrajectoryTUM for the camera pose Tcw (3x4 or 4x4 float32)."""
    T = np.ascontiguousarray(np.asarray(Tcw, np.float32)[:3, :4])
    buf, n = C.create_string_buffer(256), C.c_int(0)
    _check(lib().plslam_tum_pose_to_line(C.c_double(timestamp), _vp(T), buf, 256, C.byref(n)))
    return buf.raw[:n.value].decode("ascii")


def SaveTrajectoryTUM(filename, timestamps, poses_Tcw):
    """System::SaveTrajectoryTUM's file for per-frame camera poses Tcw [n, 3|4, 4]."""
    ts = np.ascontiguousarray(timestamps, np.float64)
    T = np.ascontiguousarray(np.asarray(poses_Tcw, np.float32)[:, :3, :4])
    assert len(ts) == len(T)
    _check(lib().plslam_tum_save_trajectory(os.fsencode(filename), _vp(ts), _vp(T), len(ts)))


# ---- ORBmatcher::SearchForTriangulation (include/plslam_b200.h: plslam_tri_job_t) ----
def epipole(R2w, t2w, Cw, fx, fy, cx, cy):
    """The epipole of KF1's camera centre in KF2 with the reference's rounding sequence (plslam_match_epipole)."""
    R, t, c = (np.ascontiguousarray(x, np.float32).ravel() for x in (R2w, t2w, Cw))
    ex, ey = C.c_float(), C.c_float()
    _check(lib().plslam_match_epipole(_vp(R), _vp(t), _vp(c), C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy),
                                      C.byref(ex), C.byref(ey)))
    return ex.value, ey.value


def _tri_job(kf1, kf2, F12, ex, ey, scale_factors, level_sigma2, only_stereo, check_ori, ptr):
    """TriJob over the arrays of two key-frame dicts (layout of tests/matchdata.py: triangulation_case); returns (job, keep)."""
    g = lambda d, k, t: np.ascontiguousarray(d[k], t)
    keep = dict(a_desc=g(kf1, "desc", np.uint8), a_xy=g(kf1, "xy", np.float32), a_angle=g(kf1, "angle", np.float32),
                a_uright=g(kf1, "uright", np.float32), a_has_mp=g(kf1, "has_mp", np.uint8), a_nodes=g(kf1, "nodes", np.int32),
                a_start=g(kf1, "start", np.int32), a_idx=g(kf1, "idx", np.int32),
                b_desc=g(kf2, "desc", np.uint8), b_xy=g(kf2, "xy", np.float32), b_angle=g(kf2, "angle", np.float32),
                b_octave=g(kf2, "octave", np.int32), b_uright=g(kf2, "uright", np.float32), b_has_mp=g(kf2, "has_mp", np.uint8),
                b_nodes=g(kf2, "nodes", np.int32), b_start=g(kf2, "start", np.int32), b_idx=g(kf2, "idx", np.int32),
                sf=np.ascontiguousarray(scale_factors, np.float32), sg=np.ascontiguousarray(level_sigma2, np.float32))
    n1, n2 = len(keep["a_desc"]), len(keep["b_desc"])
    keep["m"], keep["n"] = np.full(max(n1, 1), -9, np.int32), np.zeros(1, np.int32)
    p = ptr
    j = TriJob(p(keep["a_desc"]), p(keep["a_xy"]), p(keep["a_angle"]), p(keep["a_uright"]), p(keep["a_has_mp"]), p(keep["a_nodes"]),
               p(keep["a_start"]), p(keep["a_idx"]), p(keep["b_desc"]), p(keep["b_xy"]), p(keep["b_angle"]), p(keep["b_octave"]),
               p(keep["b_uright"]), p(keep["b_has_mp"]), p(keep["b_nodes"]), p(keep["b_start"]), p(keep["b_idx"]), p(keep["sf"]),
               p(keep["sg"]), p(keep["m"]), p(keep["n"]))
    j.F12[:] = np.asarray(F12, np.float32).ravel().tolist()
    j.ex, j.ey = float(ex), float(ey)
    j.n1, j.n2, j.n1_nodes, j.n2_nodes = n1, n2, len(keep["a_nodes"]), len(keep["b_nodes"])
    j.only_stereo, j.check_orientation = int(only_stereo), int(check_ori)
    return j, keep


def search_for_triangulation_host(kf1, kf2, F12, ex, ey, scale_factors, level_sigma2, only_stereo=False, check_ori=True):
    """ORBmatcher::SearchForTriangulation on host arrays through plslam_match_triangulation_host
    -> (vMatches12 int32 [N1], nmatches, vMatchedPairs [(i1, i2)])."""
    j, keep = _tri_job(kf1, kf2, F12, ex, ey, scale_factors, level_sigma2, only_stereo, check_ori, lambda a: a.ctypes.data)
    _check(lib().plslam_match_triangulation_host(C.byref(j), len(keep["sf"])))
    m = keep["m"][:j.n1].copy()
    return m, int(keep["n"][0]), [(int(i), int(v)) for i, v in enumerate(m) if v >= 0]
