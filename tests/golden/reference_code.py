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


# ----------------------------------------------------------------------------------------------------------------------
# The whole library, loaded.  dlopen() of lib/libORB_SLAM2.so fails only because its NEEDED libraries (OpenCV 3.3, Pangolin,
# g2o, DBoW2, libGL) are absent.  build_stubs() generates, from the library's own dynamic symbol table, one stub object that
# defines every undefined non-system symbol (functions: return 0; objects: zero bytes) plus an empty shared object for each
# missing SONAME; with those loaded first the dynamic loader accepts the reference library as it is.  Every function whose
# call graph stays inside libORB_SLAM2.so + libstdc++/libm/libc then runs unmodified: the ORBextractor constructor and
# ORBextractor::DistributeOctTree (the quad-tree) are used below.  Nothing is patched and no reference byte is copied.
# ----------------------------------------------------------------------------------------------------------------------
STUB_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "oracle", "_ref", "refstub")
_SYSTEM = ("libstdc++.so.6", "libm.so.6", "libgcc_s.so.1", "libpthread.so.0")


BUMP_C = r"""
/* operator new / delete with strictly increasing addresses and no reuse.  ORBextractor::DistributeOctTree sorts
 * pair<int, ExtractorNode*>: equal node sizes are ordered by POINTER VALUE, so the reference's result depends on the state of
 * the heap.  Under this allocator the pointer order is the allocation order, which makes the reference deterministic and is
 * the instance of its behaviour the oracle restates (its `seq` tie-break). */
#include <stddef.h>
#include <sys/mman.h>
static char* base;
static size_t off;
void* _Znwm(size_t n) {
  if (!base) base = mmap(NULL, 1UL << 34, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  n = (n + 15) & ~(size_t)15;
  void* p = base + off;
  off += n ? n : 16;
  return p;
}
void* _Znam(size_t n) { return _Znwm(n); }
void _ZdlPv(void* p) { (void)p; }
void _ZdaPv(void* p) { (void)p; }
void _ZdlPvm(void* p, size_t n) { (void)p; (void)n; }
"""


def load_monotonic_new(out=STUB_DIR):
    """Must run before anything loads libstdc++ with RTLD_GLOBAL: the first global definition of operator new wins."""
    import subprocess
    os.makedirs(out, exist_ok=True)
    src = os.path.join(out, "bump.c")
    open(src, "w").write(BUMP_C)
    subprocess.check_call(["gcc", "-O1", "-shared", "-fPIC", "-o", os.path.join(out, "libbumpnew.so"), src])
    return C.CDLL(os.path.join(out, "libbumpnew.so"), mode=C.RTLD_GLOBAL)


CVSHIM_CC = r"""
// ABI-exact stand-ins for the five OpenCV 3.3 entry points ORBextractor::ComputeKeyPointsOctTree / computeOrientation call
// (objdump of 0x75fa0-0x76da0 and 0x6fb10-0x70383): the sub-matrix constructor, cv::FAST, cv::fastAtan2 and the two release
// helpers.  FAST and fastAtan2 are the oracle's restatements, which are pinned bit for bit against cv2 4.13
// (tests/test_oracle_cv2.py, tests/golden/cv2_primitives.npz).  Symbol names are given literally, no OpenCV header exists here.
#include <climits>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <new>
extern "C" int oracle_fast9(const uint8_t* img, int w, int h, int step, int th, int* out, int cap);
extern "C" float oracle_fast_atan2(float y, float x);
namespace {
struct Range { int start, end; };
struct Mat {  // cv::Mat of OpenCV 3.x, 96 bytes
  int flags, dims, rows, cols;
  uint8_t* data;
  const uint8_t *datastart, *dataend, *datalimit;
  void* allocator;
  void* u;
  int* sizep;
  size_t* stepp;
  size_t stepbuf[2];
};
static_assert(sizeof(Mat) == 96, "cv::Mat layout");
struct KeyPoint { float x, y, size, angle, response; int octave, class_id; };
struct InputArray { int flags; void* obj; int w, h; };
struct KpVector { KeyPoint *b, *e, *c; };
const int CONTINUOUS_FLAG = 1 << 14, SUBMATRIX_FLAG = 1 << 15;
}
extern "C" {
void shim_fastFree(void*) asm("_ZN2cv8fastFreeEPv");
void shim_fastFree(void*) {}
void shim_deallocate(Mat*) asm("_ZN2cv3Mat10deallocateEv");
void shim_deallocate(Mat*) {}
// Mat::Mat(const Mat& m, const Range& rowRange, const Range& colRange), 2-D case
void shim_mat_ranges(Mat* self, const Mat* m, const Range* rr, const Range* cr) asm("_ZN2cv3MatC1ERKS0_RKNS_5RangeES5_");
void shim_mat_ranges(Mat* self, const Mat* m, const Range* rr, const Range* cr) {
  *self = *m;
  self->sizep = &self->rows;
  self->stepp = self->stepbuf;
  self->stepbuf[0] = m->stepp[0];
  self->stepbuf[1] = m->stepp[1];
  const bool allR = rr->start == INT_MIN && rr->end == INT_MAX, allC = cr->start == INT_MIN && cr->end == INT_MAX;
  if (!allR) {
    self->rows = rr->end - rr->start;
    self->data += self->stepbuf[0] * (size_t)rr->start;
    self->flags |= SUBMATRIX_FLAG;
  }
  if (!allC) {
    self->cols = cr->end - cr->start;
    self->data += (size_t)cr->start * self->stepbuf[1];
    if (self->cols < m->cols) self->flags &= ~CONTINUOUS_FLAG;
    self->flags |= SUBMATRIX_FLAG;
  }
  if (self->rows == 1) self->flags |= CONTINUOUS_FLAG;
  if (self->rows <= 0 || self->cols <= 0) self->rows = self->cols = 0;
}
// void cv::FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression)
void shim_FAST(const InputArray* img, KpVector* kps, int threshold, bool nonmax) asm("_ZN2cv4FASTERKNS_11_InputArrayERSt6vectorINS_8KeyPointESaIS4_EEib");
void shim_FAST(const InputArray* img, KpVector* kps, int threshold, bool nonmax) {
  const Mat* m = (const Mat*)img->obj;
  (void)nonmax;  // the reference passes true at both call sites
  const int cap = m->rows * m->cols + 1;
  int* tmp = (int*)::operator new(sizeof(int) * 3 * (size_t)cap);
  const int n = m->rows > 0 && m->cols > 0 ? oracle_fast9(m->data, m->cols, m->rows, (int)m->stepp[0], threshold, tmp, cap) : 0;
  KeyPoint* out = (KeyPoint*)::operator new(sizeof(KeyPoint) * (size_t)(n ? n : 1));
  for (int i = 0; i < n; ++i) out[i] = KeyPoint{(float)tmp[3 * i], (float)tmp[3 * i + 1], 7.f, -1.f, (float)tmp[3 * i + 2], 0, -1};
  ::operator delete(tmp);
  if (kps->b) ::operator delete(kps->b);
  kps->b = out;
  kps->e = out + n;
  kps->c = out + (n ? n : 1);
}
float shim_fastAtan2(float y, float x) asm("_ZN2cv9fastAtan2Eff");
float shim_fastAtan2(float y, float x) { return oracle_fast_atan2(y, x); }
}
"""


def load_cv_shims(out=STUB_DIR):
    """The five OpenCV entry points of ComputeKeyPointsOctTree, backed by the oracle's cv2-pinned primitives; loaded with
    RTLD_GLOBAL before the generic stubs so that these definitions win."""
    import subprocess
    os.makedirs(out, exist_ok=True)
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    oracle_dir = os.path.join(root, "oracle", "_build")
    src = os.path.join(out, "cvshim.cc")
    open(src, "w").write(CVSHIM_CC)
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", os.path.join(out, "libcvshim.so"), src,
                           "-L" + oracle_dir, "-loracle", "-Wl,-rpath," + oracle_dir])
    return C.CDLL(os.path.join(out, "libcvshim.so"), mode=C.RTLD_GLOBAL)


def build_stubs(so=SO, out=STUB_DIR):
    import re
    import subprocess
    os.makedirs(out, exist_ok=True)
    for lib in _SYSTEM:
        C.CDLL(lib, mode=C.RTLD_GLOBAL)
    me = C.CDLL(None)
    funcs, objs = [], []
    for line in subprocess.run(["readelf", "--dyn-syms", "-W", so], capture_output=True, text=True, check=True).stdout.splitlines():
        p = line.split()
        if len(p) < 8 or p[6] != "UND" or p[4] == "WEAK":
            continue
        name = p[7].split("@")[0]
        if hasattr(me, name):
            continue  # the process already provides it (libc, libm, libstdc++ ...)
        (objs if p[3] == "OBJECT" else funcs).append(name)
    with open(os.path.join(out, "stub.s"), "w") as f:
        f.write(".text\nrefstub_fn:\n xorl %eax,%eax\n ret\n")
        for n in funcs:
            f.write(".globl %s\n.type %s,@function\n.set %s, refstub_fn\n" % (n, n, n))
        f.write(".data\n.align 64\n")
        for n in objs:
            f.write(".globl %s\n.type %s,@object\n.size %s,1024\n%s:\n .zero 1024\n" % (n, n, n, n))
        f.write('.section .note.GNU-stack,"",@progbits\n')
    open(os.path.join(out, "empty.c"), "w").write("")
    dyn = subprocess.run(["readelf", "-d", so], capture_output=True, text=True, check=True).stdout
    needed = [l.split("[")[1].rstrip("]") for l in dyn.splitlines() if "(NEEDED)" in l]
    needed = [n for n in needed if not re.match(r"lib(stdc\+\+|m|gcc_s|pthread|c)\.so", n)]
    subprocess.check_call(["gcc", "-shared", "-o", os.path.join(out, "librefstub.so"), os.path.join(out, "stub.s"),
                           "-Wl,-soname,librefstub.so"])
    for n in needed:
        subprocess.check_call(["gcc", "-shared", "-o", os.path.join(out, n), os.path.join(out, "empty.c"), "-Wl,-soname," + n])
    return needed


class RefLibrary:
    """lib/libORB_SLAM2.so dlopen'ed over the stubs; call build_stubs() first (make_golden does)."""

    KP = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4"),
                   ("class_id", "<i4")])

    def __init__(self, so=SO, monotonic_new=True):
        if monotonic_new:
            self._bump = load_monotonic_new()
        self._shims = load_cv_shims()
        needed = build_stubs(so)
        C.CDLL(os.path.join(STUB_DIR, "librefstub.so"), mode=C.RTLD_GLOBAL)
        for n in needed:
            C.CDLL(os.path.join(STUB_DIR, n), mode=C.RTLD_GLOBAL)
        self.lib = C.CDLL(so, mode=C.RTLD_GLOBAL)
        self._ctor = getattr(self.lib, "_ZN9ORB_SLAM212ORBextractorC1Eifiii")
        self._ctor.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        self._ctor.restype = None
        self._dist = getattr(self.lib, "_ZN9ORB_SLAM212ORBextractor17DistributeOctTreeERKSt6vectorIN2cv8KeyPointESaIS3_EERKiS9_S9_S9_S9_S9_")
        self._dist.argtypes = [C.c_void_p] * 9
        self._dist.restype = C.c_void_p
        self._ckp = getattr(self.lib, "_ZN9ORB_SLAM212ORBextractor23ComputeKeyPointsOctTreeERSt6vectorIS1_IN2cv8KeyPointESaIS3_EESaIS5_EE")
        self._ckp.argtypes = [C.c_void_p, C.c_void_p]
        self._ckp.restype = None

    @staticmethod
    def _vec(buf, off, dtype):
        b, e = buf[off // 8], buf[off // 8 + 1]
        n = (e - b) // np.dtype(dtype).itemsize
        return np.ctypeslib.as_array(C.cast(b, C.POINTER(C.c_uint8)), ((e - b),)).view(dtype)[:n].copy() if n else np.empty(0, dtype)

    def extractor(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        """Runs the reference constructor; returns (object memory, tables).  Layout from include/ORBextractor.h:85-110:
        mvImagePyramid @0, pattern @0x18, nfeatures @0x30, scaleFactor (double) @0x38, nlevels @0x40, iniThFAST @0x44,
        minThFAST @0x48, mnFeaturesPerLevel @0x50, umax @0x68, mvScaleFactor @0x80, mvInvScaleFactor @0x98, mvLevelSigma2 @0xb0,
        mvInvLevelSigma2 @0xc8."""
        obj = (C.c_uint64 * 64)()
        self._ctor(C.addressof(obj), nfeatures, scale_factor, nlevels, ini_th, min_th)
        t = dict(quota=self._vec(obj, 0x50, np.int32), umax=self._vec(obj, 0x68, np.int32), scale=self._vec(obj, 0x80, np.float32),
                 inv_scale=self._vec(obj, 0x98, np.float32), sigma2=self._vec(obj, 0xb0, np.float32),
                 inv_sigma2=self._vec(obj, 0xc8, np.float32), pattern=self._vec(obj, 0x18, np.int32).reshape(-1, 2),
                 nfeatures=int(np.frombuffer(obj, np.int32, 1, 0x30)[0]), scale_factor=float(np.frombuffer(obj, np.float64, 1, 0x38)[0]),
                 nlevels=int(np.frombuffer(obj, np.int32, 1, 0x40)[0]))
        return obj, t

    def distribute(self, obj, xyr, minX, maxX, minY, maxY, N, level=0):
        """ORBextractor::DistributeOctTree on (x, y, response) rows; returns the input indices in output order."""
        xyr = np.asarray(xyr)
        keys = np.zeros(len(xyr), self.KP)
        keys["x"], keys["y"], keys["response"] = xyr[:, 0], xyr[:, 1], xyr[:, 2]
        keys["size"], keys["angle"], keys["class_id"] = 7, -1, np.arange(len(xyr))
        vec = (C.c_uint64 * 3)(keys.ctypes.data, keys.ctypes.data + keys.nbytes, keys.ctypes.data + keys.nbytes)
        ret = (C.c_uint64 * 3)()
        ints = [C.c_int(v) for v in (minX, maxX, minY, maxY, N, level)]
        self._dist(C.addressof(ret), C.addressof(obj), C.addressof(vec), *[C.addressof(i) for i in ints])
        n = (ret[1] - ret[0]) // 28
        out = np.ctypeslib.as_array(C.cast(ret[0], C.POINTER(C.c_uint8)), (n * 28,)).view(self.KP).copy() if n else np.empty(0, self.KP)
        return out["class_id"].astype(np.int32), out

    def compute_keypoints_oct_tree(self, obj, level_images):
        """ORBextractor::ComputeKeyPointsOctTree(allKeypoints) on a pyramid supplied by the caller (one contiguous uint8 image
        per level, written into the object's mvImagePyramid as header-only cv::Mat).  cv::FAST / cv::fastAtan2 are the shims
        above.  Returns one KeyPoint array per level: level coordinates, size, octave and IC angle as the reference sets them."""
        mats_begin = obj[0]
        keep = []
        for l, img in enumerate(level_images):
            img = np.ascontiguousarray(img, np.uint8)
            keep.append(img)
            m = (C.c_uint64 * 12).from_address(mats_begin + 96 * l)
            base = mats_begin + 96 * l
            m[0] = (2 << 32) | (0x42FF0000 | (1 << 14))
            m[1] = (img.shape[1] << 32) | img.shape[0]
            m[2] = img.ctypes.data
            m[3] = img.ctypes.data
            m[4] = img.ctypes.data + img.nbytes
            m[5] = img.ctypes.data + img.nbytes
            m[6] = 0
            m[7] = 0
            m[8] = base + 8
            m[9] = base + 0x50
            m[10] = img.strides[0]
            m[11] = 1
        allk = (C.c_uint64 * 3)()
        self._ckp(C.addressof(obj), C.addressof(allk))
        n = (allk[1] - allk[0]) // 24
        out = []
        for l in range(n):
            v = (C.c_uint64 * 3).from_address(allk[0] + 24 * l)
            cnt = (v[1] - v[0]) // 28
            out.append(np.ctypeslib.as_array(C.cast(v[0], C.POINTER(C.c_uint8)), (cnt * 28,)).view(self.KP).copy() if cnt
                       else np.empty(0, self.KP))
        return out
