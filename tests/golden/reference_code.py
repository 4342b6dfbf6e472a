"""Executes leaf functions of the reference's own machine code (lib/libORB_SLAM2.so) in this process.

The library cannot be dlopen'ed here (its OpenCV 3.3 / Pangolin / g2o dependencies are absent: SURVEY.md 8c), but four
functions on the matcher path are self-contained leaves — no calls, only rip-relative loads of .rodata constants that lie in the
same LOAD segment (file offset == virtual address, `readelf -l`): mapping the file read+execute and calling them through ctypes
with hand-built argument structs runs the reference's arithmetic itself, FMA contractions included:

  ORBmatcher::RadiusByViewingCos(const float&)                                   @0x79b60
  ORBmatcher::CheckDistEpipolarLine(KeyPoint const&, KeyPoint const&, Mat const&, KeyFrame const*)   @0x79b90
  ORBmatcher::ComputeThreeMaxima(vector<int>*, int, int&, int&, int&)            @0x79c40
  ORBmatcher::DescriptorDistance(Mat const&, Mat const&)                         @0x79d20

Only the build container has /root/reference; tests/golden/make_golden.py uses this module to write
tests/golden/reference_code.npz, which the tests read.  Needs an x86-64 CPU with AVX2/FMA (the binary was built -march=native).
"""
import ctypes as C
import os

import numpy as np

SO = "/root/reference/lib/libORB_SLAM2.so"
SHA256 = None  # filled by make_golden into the fixture


class RefCode:
    def __init__(self, path=SO):
        libc = C.CDLL(None, use_errno=True)
        libc.mmap.restype = C.c_void_p
        libc.mmap.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_long]
        self.size = os.path.getsize(path)
        fd = os.open(path, os.O_RDONLY)
        PROT_READ, PROT_EXEC, MAP_PRIVATE = 1, 4, 2
        self.base = libc.mmap(None, self.size, PROT_READ | PROT_EXEC, MAP_PRIVATE, fd, 0)
        os.close(fd)
        assert self.base not in (None, C.c_void_p(-1).value), "mmap failed"
        f = lambda ret, addr, *args: C.CFUNCTYPE(ret, *args)(self.base + addr)
        self._radius = f(C.c_float, 0x79b60, C.c_void_p, C.POINTER(C.c_float))
        self._check = f(C.c_uint8, 0x79b90, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p)
        self._maxima = f(None, 0x79c40, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int))
        self._dist = f(C.c_int, 0x79d20, C.c_void_p, C.c_void_p)

    # cv::Mat of OpenCV 3.x, 96 bytes: flags, dims, rows, cols, data @0x10, ..., size.p @0x40, step.p @0x48, step.buf @0x50
    @staticmethod
    def _mat(arr):
        arr = np.ascontiguousarray(arr)
        m = (C.c_uint64 * 12)()
        base = C.addressof(m)
        m[0] = (2 << 32) | 0x42ff4000          # flags (unused by the callees), dims = 2
        m[1] = (arr.shape[1] << 32) | arr.shape[0]
        m[2] = arr.ctypes.data                 # data
        m[8] = base + 8                        # size.p -> rows
        m[9] = base + 0x50                     # step.p -> step.buf
        m[10] = arr.strides[0]
        m[11] = arr.itemsize
        return m, arr

    def radius_by_viewing_cos(self, v):
        x = C.c_float(v)
        return float(self._radius(None, C.byref(x)))

    def descriptor_distance(self, a, b):
        ma, ka = self._mat(np.asarray(a, np.uint8).reshape(1, 32))
        mb, kb = self._mat(np.asarray(b, np.uint8).reshape(1, 32))
        return int(self._dist(C.addressof(ma), C.addressof(mb)))

    def check_dist_epipolar_line(self, kp1_xy, kp2_xy, kp2_octave, F12, level_sigma2):
        kp = np.zeros(2, np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                                   ("octave", "<i4"), ("class_id", "<i4")]))
        kp[0]["x"], kp[0]["y"] = kp1_xy
        kp[1]["x"], kp[1]["y"], kp[1]["octave"] = kp2_xy[0], kp2_xy[1], kp2_octave
        mF, keepF = self._mat(np.asarray(F12, np.float32).reshape(3, 3))
        sig = np.ascontiguousarray(level_sigma2, np.float32)
        kf = (C.c_uint64 * 0x62)()             # KeyFrame: only mvLevelSigma2's begin pointer @0x300 is read
        kf[0x300 // 8] = sig.ctypes.data
        return bool(self._check(None, kp.ctypes.data, kp.ctypes.data + 28, C.addressof(mF), C.addressof(kf)) & 1)

    def compute_three_maxima(self, sizes):
        """sizes: the 30 bin populations (the function only reads vector sizes)"""
        L = len(sizes)
        store = [np.zeros(max(int(s), 1), np.int32) for s in sizes]
        vec = (C.c_uint64 * (3 * L))()
        for i, (s, a) in enumerate(zip(sizes, store)):
            vec[3 * i] = a.ctypes.data
            vec[3 * i + 1] = a.ctypes.data + 4 * int(s)
            vec[3 * i + 2] = a.ctypes.data + 4 * len(a)
        i1, i2, i3 = C.c_int(-7), C.c_int(-7), C.c_int(-7)
        self._maxima(None, C.addressof(vec), L, C.byref(i1), C.byref(i2), C.byref(i3))
        return i1.value, i2.value, i3.value
