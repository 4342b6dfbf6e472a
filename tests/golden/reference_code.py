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


def _build_atomically(cmd, target):
    """Link to a temporary name and rename: another process may have the previous file mapped (rewriting a mapped shared
    object in place corrupts that process)."""
    import subprocess
    tmp = "%s.tmp.%d" % (target, os.getpid())
    subprocess.check_call(cmd + ["-o", tmp])
    os.replace(tmp, target)


def load_monotonic_new(out=STUB_DIR):
    """Must run before anything loads libstdc++ with RTLD_GLOBAL: the first global definition of operator new wins."""
    import subprocess
    os.makedirs(out, exist_ok=True)
    src = os.path.join(out, "bump.c")
    open(src, "w").write(BUMP_C)
    _build_atomically(["gcc", "-O1", "-shared", "-fPIC", src], os.path.join(out, "libbumpnew.so"))
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
#include <map>
#include <set>
#include <cmath>
#include <vector>
#include <cstdio>
#include <cstdlib>
static const bool g_trace = std::getenv("REFSHIM_TRACE") != nullptr;
#define TRACE(...) do { if (g_trace) { std::fprintf(stderr, __VA_ARGS__); std::fputc(10, stderr); } } while (0)
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
  TRACE("Mat(m, rows %d..%d, cols %d..%d) of %dx%d", rr->start, rr->end, cr->start, cr->end, m->rows, m->cols);
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
  TRACE("FAST th %d", threshold);
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

// ---- the further entry points of ORBextractor::operator() (0x76da0-0x782ee) and ComputePyramid (0x70430-0x70c53) ----
// Every Mat made here has u == NULL: the reference's inlined release() then never calls deallocate(), memory is simply
// not returned (the process is a fixture generator).
static const int MAGIC = 0x42FF0000, TYPE_MASK = 0xFFF, KIND_MASK = 31 << 16, KIND_MAT = 1 << 16;
static void mat_init_empty(Mat* m) {
  std::memset(m, 0, sizeof(Mat));
  m->flags = MAGIC;
  m->sizep = &m->rows;
  m->stepp = m->stepbuf;
}
static void mat_create(Mat* m, int rows, int cols, int type) {
  if (m->data && m->dims == 2 && (m->flags & TYPE_MASK) == type && m->rows == rows && m->cols == cols) return;
  static const size_t depth_size[8] = {1, 1, 2, 2, 4, 4, 8, 2};
  const size_t esz = depth_size[type & 7] * (size_t)(((type >> 3) & 511) + 1);  // CV_8UC1, CV_32FC1, CV_32FC2 occur
  mat_init_empty(m);
  m->flags = MAGIC | CONTINUOUS_FLAG | type;
  m->dims = 2;
  m->rows = rows;
  m->cols = cols;
  m->stepbuf[0] = (size_t)cols * esz;
  m->stepbuf[1] = esz;
  const size_t total = (size_t)rows * cols * esz;
  m->data = (uint8_t*)::operator new(total ? total : 1);
  std::memset(m->data, 0, total);
  m->datastart = m->data;
  m->dataend = m->datalimit = m->data + total;
}
static Mat* arr_mat(const InputArray* a) { return (a->flags & KIND_MASK) == KIND_MAT ? (Mat*)a->obj : nullptr; }

void shim_mat_dtor(Mat*) asm("_ZN2cv3MatD1Ev");
void shim_mat_dtor(Mat*) {}
void shim_copySize(Mat*, const Mat*) asm("_ZN2cv3Mat8copySizeERKS0_");
void shim_copySize(Mat*, const Mat*) {}
void shim_mat_create(Mat* m, int d, const int* sizes, int type) asm("_ZN2cv3Mat6createEiPKii");
void shim_mat_create(Mat* m, int d, const int* sizes, int type) {
  TRACE("Mat::create %d x %d type %d", sizes[0], sizes[1], type);
  if (d != 2) __builtin_trap();
  mat_create(m, sizes[0], sizes[1], type & TYPE_MASK);
}
struct Rect { int x, y, w, h; };
void shim_mat_rect(Mat* self, const Mat* m, const Rect* r) asm("_ZN2cv3MatC1ERKS0_RKNS_5Rect_IiEE");
void shim_mat_rect(Mat* self, const Mat* m, const Rect* r) {
  TRACE("Mat(m, Rect %d %d %d %d) of %dx%d", r->x, r->y, r->w, r->h, m->rows, m->cols);
  *self = *m;
  self->sizep = &self->rows;
  self->stepp = self->stepbuf;
  self->stepbuf[0] = m->stepp[0];
  self->stepbuf[1] = m->stepp[1];
  self->rows = r->h;
  self->cols = r->w;
  self->data += (size_t)r->y * self->stepbuf[0] + (size_t)r->x * self->stepbuf[1];
  if (r->w < m->cols) self->flags &= ~CONTINUOUS_FLAG;
  if (r->h == 1) self->flags |= CONTINUOUS_FLAG;
  if (r->w < m->cols || r->h < m->rows) self->flags |= SUBMATRIX_FLAG;
  if (self->rows <= 0 || self->cols <= 0) self->rows = self->cols = 0;
}
int shim_kind(const InputArray* a) asm("_ZNK2cv11_InputArray4kindEv");
int shim_kind(const InputArray* a) {
  TRACE("kind() flags %x", a->flags); return a->flags & KIND_MASK; }
bool shim_empty(const InputArray* a) asm("_ZNK2cv11_InputArray5emptyEv");
bool shim_empty(const InputArray* a) {
  TRACE("empty() flags %x", a->flags);
  const Mat* m = arr_mat(a);
  return !m || !m->data || m->rows == 0 || m->cols == 0;
}
void shim_getMat(Mat* ret, const InputArray* a, int idx) asm("_ZNK2cv11_InputArray7getMat_Ei");
void shim_getMat(Mat* ret, const InputArray* a, int idx) {
  TRACE("getMat_(%d) flags %x", idx, a->flags);
  (void)idx;
  const Mat* m = arr_mat(a);
  if (!m) { mat_init_empty(ret); return; }
  *ret = *m;
  ret->sizep = &ret->rows;
  ret->stepp = ret->stepbuf;
  ret->stepbuf[0] = m->stepp[0];
  ret->stepbuf[1] = m->stepp[1];
}
void shim_out_create(const InputArray* a, int rows, int cols, int type, int i, bool allowT, int mask) asm("_ZNK2cv12_OutputArray6createEiiiibi");
void shim_out_create(const InputArray* a, int rows, int cols, int type, int i, bool allowT, int mask) {
  TRACE("OutputArray::create %d x %d type %d", rows, cols, type);
  (void)i; (void)allowT; (void)mask;
  Mat* m = arr_mat(a);
  if (!m) __builtin_trap();
  mat_create(m, rows, cols, type & TYPE_MASK);
}
void shim_out_release(const InputArray* a) asm("_ZNK2cv12_OutputArray7releaseEv");
void shim_out_release(const InputArray* a) {
  TRACE("OutputArray::release");
  Mat* m = arr_mat(a);
  if (m) mat_init_empty(m);
}
void shim_copyTo(const Mat* src, const InputArray* dst) asm("_ZNK2cv3Mat6copyToERKNS_12_OutputArrayE");
void shim_copyTo(const Mat* src, const InputArray* dst) {
  TRACE("copyTo %dx%d", src->rows, src->cols);
  Mat* d = arr_mat(dst);
  if (!d) __builtin_trap();
  mat_create(d, src->rows, src->cols, src->flags & TYPE_MASK);
  if (d->data == src->data) return;
  for (int r = 0; r < src->rows; ++r)
    std::memcpy(d->data + (size_t)r * d->stepp[0], src->data + (size_t)r * src->stepp[0], (size_t)src->cols * src->stepp[1]);
}
extern "C" void oracle_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstep, uint8_t* dst, int dw, int dh, int dstep);
extern "C" void oracle_blur7_u8(const uint8_t* src, int w, int h, int sstep, uint8_t* dst, int dstep, const int* k);
// void cv::resize(InputArray src, OutputArray dst, Size dsize, double fx, double fy, int interpolation).  cv::Size_ has a
// user-provided copy constructor in OpenCV 3.x, so the Itanium ABI passes it by invisible reference: a pointer to {w, h}
void shim_resize(const InputArray* src, const InputArray* dst, const int* dsize, double fx, double fy, int interp) asm("_ZN2cv6resizeERKNS_11_InputArrayERKNS_12_OutputArrayENS_5Size_IiEEddi");
void shim_resize(const InputArray* src, const InputArray* dst, const int* dsize, double fx, double fy, int interp) {
  TRACE("resize -> %d x %d interp %d", dsize[0], dsize[1], interp);
  (void)fx; (void)fy;
  if (interp != 1) __builtin_trap();  // INTER_LINEAR at the one call site (@0x70b07)
  const Mat* s = arr_mat(src);
  Mat* d = arr_mat(dst);
  const int dw = dsize[0], dh = dsize[1];
  mat_create(d, dh, dw, s->flags & TYPE_MASK);  // same size already: the level's view into its bordered buffer is kept
  oracle_resize_linear_u8(s->data, s->cols, s->rows, (int)s->stepp[0], d->data, dw, dh, (int)d->stepp[0]);
}
static int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
  return p;
}
// void cv::copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType, const Scalar&)
void shim_copyMakeBorder(const InputArray* src, const InputArray* dst, int top, int bottom, int left, int right, int borderType, const void* value)
    asm("_ZN2cv14copyMakeBorderERKNS_11_InputArrayERKNS_12_OutputArrayEiiiiiRKNS_7Scalar_IdEE");
void shim_copyMakeBorder(const InputArray* src, const InputArray* dst, int top, int bottom, int left, int right, int borderType, const void* value) {
  TRACE("copyMakeBorder %d %d %d %d type %d", top, bottom, left, right, borderType);
  (void)value;
  if ((borderType & ~16) != 4) __builtin_trap();  // BORDER_REFLECT_101, optionally | BORDER_ISOLATED
  const Mat s = *arr_mat(src);                    // header copy: dst may be the buffer src is a view of
  const size_t sstep = arr_mat(src)->stepp[0];
  Mat* d = arr_mat(dst);
  mat_create(d, s.rows + top + bottom, s.cols + left + right, s.flags & TYPE_MASK);
  const size_t dstep = d->stepp[0];
  uint8_t* centre = d->data + (size_t)top * dstep + left;
  if (centre != s.data)
    for (int r = 0; r < s.rows; ++r) std::memmove(centre + (size_t)r * dstep, s.data + (size_t)r * sstep, (size_t)s.cols);
  for (int r = 0; r < s.rows; ++r) {
    uint8_t* row = centre + (size_t)r * dstep;
    for (int c = -left; c < 0; ++c) row[c] = row[reflect101(c, s.cols)];
    for (int c = s.cols; c < s.cols + right; ++c) row[c] = row[reflect101(c, s.cols)];
  }
  for (int r = -top; r < s.rows + bottom; ++r) {
    if (r >= 0 && r < s.rows) continue;
    std::memcpy(d->data + (size_t)(r + top) * dstep, d->data + (size_t)(reflect101(r, s.rows) + top) * dstep, (size_t)(s.cols + left + right));
  }
}
// void cv::GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY, int borderType)
void shim_GaussianBlur(const InputArray* src, const InputArray* dst, const int* ksize, double sx, double sy, int borderType)
    asm("_ZN2cv12GaussianBlurERKNS_11_InputArrayERKNS_12_OutputArrayENS_5Size_IiEEddi");
void shim_GaussianBlur(const InputArray* src, const InputArray* dst, const int* ksize, double sx, double sy, int borderType) {
  TRACE("GaussianBlur ksize %d x %d sigma %g %g border %d", ksize[0], ksize[1], sx, sy, borderType);
  if (ksize[0] != 7 || ksize[1] != 7 || sx != 2.0 || sy != 2.0 || borderType != 4) __builtin_trap();  // the one call (@0x77487)
  const Mat* s = arr_mat(src);
  Mat* d = arr_mat(dst);
  static const int k[7] = {18, 34, 48, 56, 48, 34, 18};  // getGaussianKernel(7, 2) in 8-bit fixed point, pinned against cv2
  const size_t n = (size_t)s->rows * s->cols;
  uint8_t* tmp = (uint8_t*)::operator new(n ? n : 1);
  for (int r = 0; r < s->rows; ++r) std::memcpy(tmp + (size_t)r * s->cols, s->data + (size_t)r * s->stepp[0], (size_t)s->cols);
  const int rows = s->rows, cols = s->cols, type = s->flags & TYPE_MASK;
  mat_create(d, rows, cols, type);
  oracle_blur7_u8(tmp, cols, rows, cols, d->data, (int)d->stepp[0], k);
  ::operator delete(tmp);
}
// MatExpr cv::Mat::zeros(int rows, int cols, int type) and the MatOp whose assign() the caller invokes through vtable slot 3
struct MatExpr { const void* op; int flags; int pad; Mat a, b, c; double alpha, beta; double s[4]; };
static_assert(sizeof(MatExpr) == 352 && offsetof(MatExpr, c) == 0xd0, "cv::MatExpr layout");
static void op_dtor(void*) {}
static bool op_elementwise(const void*, const MatExpr*) { return false; }
static void op_assign(const void*, const MatExpr* e, Mat* m, int type) {
  TRACE("MatOp::assign into %dx%d", m->rows, m->cols);
  (void)type;
  mat_create(m, (int)e->alpha, (int)e->beta, e->flags & TYPE_MASK);
  for (int r = 0; r < m->rows; ++r) std::memset(m->data + (size_t)r * m->stepp[0], 0, (size_t)m->cols);
}
static const void* const op_vtable[4] = {(const void*)op_dtor, (const void*)op_dtor, (const void*)op_elementwise, (const void*)op_assign};
static const void* const op_object[1] = {op_vtable};
void shim_zeros(MatExpr* ret, int rows, int cols, int type) asm("_ZN2cv3Mat5zerosEiii");
// helper for the harness (not an OpenCV symbol): a DBoW2::FeatureVector (= std::map<unsigned, std::vector<unsigned>>, same
// libstdc++ layout in GCC 5 and today) built in place from CSR arrays
void refshim_build_featvec(void* where, const int* nodes, const int* start, const int* idx, int n);
void refshim_build_featvec(void* where, const int* nodes, const int* start, const int* idx, int n) {
  auto* m = new (where) std::map<unsigned int, std::vector<unsigned int>>();
  for (int k = 0; k < n; ++k) (*m)[(unsigned)nodes[k]].assign(idx + start[k], idx + start[k + 1]);
}
void shim_zeros(MatExpr* ret, int rows, int cols, int type) {
  TRACE("Mat::zeros %d x %d type %d", rows, cols, type);
  std::memset(ret, 0, sizeof(MatExpr));
  ret->op = op_object;
  ret->flags = type & TYPE_MASK;
  mat_init_empty(&ret->a);
  mat_init_empty(&ret->b);
  mat_init_empty(&ret->c);
  ret->alpha = rows;
  ret->beta = cols;
}

// ---- the matrix expressions of ORBmatcher::SearchByProjection(Frame&, const Frame&, ...) (@0x80d00): -Rcw.t()*tcw,
// Rlw*twc+tlw, Rcw*x3Dw+tcw.  Expressions stay lazy as in OpenCV (MatExpr flags: 'T' = a^T * alpha, 'G' = alpha*op(a)*b + beta*c);
// the evaluation in assign() is cv::gemm's for CV_32F: with no transpose flag and an inner dimension of 2..4 the small-matrix
// path (float products summed left to right in float, then (double)sum*alpha + (double)c*beta), otherwise GEMMSingleMul (double
// accumulator, (float)(sum*alpha [+ c*beta])).  The small-matrix arithmetic is pinned against cv2 4.13 (tests/test_tum_io_cpu.py).
static float& at(const Mat* m, int r, int c) { return *(float*)(m->data + (size_t)r * m->stepp[0] + (size_t)c * 4); }
static void hdr_copy(Mat* d, const Mat* s) {
  *d = *s;
  d->sizep = &d->rows;
  d->stepp = d->stepbuf;
  d->stepbuf[0] = s->stepp[0];
  d->stepbuf[1] = s->stepp[1];
}
static void expr_assign(const void*, const MatExpr* e, Mat* m, int type) {
  (void)type;
  if ((e->flags >> 8) == 'Z') { op_assign(nullptr, e, m, type); return; }
  if ((e->flags >> 8) == 'S') {  // a - b, element by element in float (cv::subtract on CV_32F)
    const Mat *A = &e->a, *B = &e->b;
    if (A->rows != B->rows || A->cols != B->cols) __builtin_trap();
    Mat D;
    mat_init_empty(&D);
    mat_create(&D, A->rows, A->cols, 5);
    for (int i = 0; i < A->rows; ++i)
      for (int j = 0; j < A->cols; ++j) at(&D, i, j) = at(A, i, j) - at(B, i, j);
    mat_create(m, A->rows, A->cols, 5);
    for (int i = 0; i < A->rows; ++i) std::memcpy(m->data + (size_t)i * m->stepp[0], D.data + (size_t)i * D.stepp[0], (size_t)A->cols * 4);
    return;
  }
  if ((e->flags >> 8) == 'A' || (e->flags >> 8) == 'T') {
    // 'A': s * a or -a = MatOp_AddEx(a, alpha): convertTo with a scale, dst = src * (float)alpha + (float)0.
    // 'T': alpha * a.t() = MatOp_T: transpose, then (alpha != 1) the same scaled conversion.
    const Mat* A = &e->a;
    const bool tr = (e->flags >> 8) == 'T';
    const int rows = tr ? A->cols : A->rows, cols = tr ? A->rows : A->cols;
    const float scale = (float)e->alpha;
    Mat D;
    mat_init_empty(&D);
    mat_create(&D, rows, cols, 5);
    for (int i = 0; i < rows; ++i)
      for (int j = 0; j < cols; ++j) {
        const float v = tr ? at(A, j, i) : at(A, i, j);
        if (tr && e->alpha == 1) { at(&D, i, j) = v; continue; }
        volatile float prod = v * scale;
        at(&D, i, j) = prod + 0.0f;
      }
    mat_create(m, rows, cols, 5);
    for (int i = 0; i < rows; ++i) std::memcpy(m->data + (size_t)i * m->stepp[0], D.data + (size_t)i * D.stepp[0], (size_t)cols * 4);
    return;
  }
  if ((e->flags >> 8) == 'D') {  // a / s = MatOp_AddEx(a, alpha = 1./s): convertTo with a scale, cvtScale_<float, float, float>:
    const Mat* A = &e->a;         // dst = src * (float)alpha + (float)0
    const float scale = (float)e->alpha, shift = 0.0f;
    Mat D;
    mat_init_empty(&D);
    mat_create(&D, A->rows, A->cols, 5);
    for (int i = 0; i < A->rows; ++i)
      for (int j = 0; j < A->cols; ++j) {
        volatile float prod = at(A, i, j) * scale;  // two roundings, as the SIMD mul + add of OpenCV 3.3's cvtScale
        at(&D, i, j) = prod + shift;
      }
    mat_create(m, A->rows, A->cols, 5);
    for (int i = 0; i < A->rows; ++i) std::memcpy(m->data + (size_t)i * m->stepp[0], D.data + (size_t)i * D.stepp[0], (size_t)A->cols * 4);
    return;
  }
  if ((e->flags >> 8) != 'G') __builtin_trap();
  const bool tA = (e->flags & 1) != 0;
  const Mat *A = &e->a, *B = &e->b, *Cm = e->c.data ? &e->c : nullptr;
  const int rows = tA ? A->cols : A->rows, len = tA ? A->rows : A->cols, cols = B->cols;
  Mat D;
  mat_init_empty(&D);
  mat_create(&D, rows, cols, 5);
  const bool small = !tA && len >= 2 && len <= 4 && (len == cols || len == rows);
  for (int i = 0; i < rows; ++i)
    for (int j = 0; j < cols; ++j) {
      if (small) {
        float t = at(A, i, 0) * at(B, 0, j);
        for (int k = 1; k < len; ++k) t = t + at(A, i, k) * at(B, k, j);
        at(&D, i, j) = (float)((double)t * e->alpha + (Cm ? (double)at(Cm, i, j) : 0.0) * (Cm ? e->beta : 0.0));
      } else {
        double s0 = 0;
        for (int k = 0; k < len; ++k) s0 += (double)(tA ? at(A, k, i) : at(A, i, k)) * (double)at(B, k, j);
        at(&D, i, j) = Cm ? (float)(s0 * e->alpha + (double)at(Cm, i, j) * e->beta) : (float)(s0 * e->alpha);
      }
    }
  mat_create(m, rows, cols, 5);
  for (int i = 0; i < rows; ++i) std::memcpy(m->data + (size_t)i * m->stepp[0], D.data + (size_t)i * D.stepp[0], (size_t)cols * 4);
}
static const void* const expr_vtable[4] = {(const void*)op_dtor, (const void*)op_dtor, (const void*)op_elementwise, (const void*)expr_assign};
static const void* const expr_object[1] = {expr_vtable};
static void expr_init(MatExpr* e, int kind, int tA) {
  std::memset(e, 0, sizeof(MatExpr));
  e->op = expr_object;
  e->flags = (kind << 8) | tA;
  mat_init_empty(&e->a);
  mat_init_empty(&e->b);
  mat_init_empty(&e->c);
  e->alpha = 1;
}
void shim_t(MatExpr* ret, const Mat* m) asm("_ZNK2cv3Mat1tEv");
void shim_t(MatExpr* ret, const Mat* m) {
  TRACE("Mat::t()");
  expr_init(ret, 'T', 1);
  hdr_copy(&ret->a, m);
}
void shim_neg(MatExpr* ret, const MatExpr* e) asm("_ZN2cvngERKNS_7MatExprE");
void shim_neg(MatExpr* ret, const MatExpr* e) {
  TRACE("operator-(MatExpr)");
  if ((e->flags >> 8) != 'T') __builtin_trap();
  expr_init(ret, 'T', 1);
  hdr_copy(&ret->a, &e->a);
  ret->alpha = -e->alpha;
}
void shim_mul_em(MatExpr* ret, const MatExpr* e, const Mat* m) asm("_ZN2cvmlERKNS_7MatExprERKNS_3MatE");
void shim_mul_em(MatExpr* ret, const MatExpr* e, const Mat* m) {
  TRACE("operator*(MatExpr, Mat)");
  // MatOp::matmul: a transposed operand becomes GEMM_A_T with scale = alpha; a scaled operand (-a, s * a) GEMM with scale = alpha
  if ((e->flags >> 8) != 'T' && (e->flags >> 8) != 'A') __builtin_trap();
  expr_init(ret, 'G', (e->flags >> 8) == 'T' ? 1 : 0);
  hdr_copy(&ret->a, &e->a);
  hdr_copy(&ret->b, m);
  ret->alpha = e->alpha;
}
void shim_mul_mm(MatExpr* ret, const Mat* a, const Mat* b) asm("_ZN2cvmlERKNS_3MatES2_");
void shim_mul_mm(MatExpr* ret, const Mat* a, const Mat* b) {
  TRACE("operator*(Mat, Mat)");
  expr_init(ret, 'G', 0);
  hdr_copy(&ret->a, a);
  hdr_copy(&ret->b, b);
}
void shim_add_em(MatExpr* ret, const MatExpr* e, const Mat* m) asm("_ZN2cvplERKNS_7MatExprERKNS_3MatE");
void shim_add_em(MatExpr* ret, const MatExpr* e, const Mat* m) {
  TRACE("operator+(MatExpr, Mat)");
  if ((e->flags >> 8) != 'G' || e->c.data) __builtin_trap();
  *ret = *e;
  hdr_copy(&ret->a, &e->a);
  hdr_copy(&ret->b, &e->b);
  hdr_copy(&ret->c, m);
  ret->beta = 1;
}
// ---- ORBmatcher::SearchByProjection(Frame&, KeyFrame*, set<MapPoint*>&, th, ORBdist) (@0x7e8c0): PO = x3Dw - Ow, cv::norm(PO) ----
void shim_sub_mm(MatExpr* ret, const Mat* a, const Mat* b) asm("_ZN2cvmiERKNS_3MatES2_");
void shim_sub_mm(MatExpr* ret, const Mat* a, const Mat* b) {
  TRACE("operator-(Mat, Mat)");
  expr_init(ret, 'S', 0);
  hdr_copy(&ret->a, a);
  hdr_copy(&ret->b, b);
}
static InputArray g_noarray = {0, nullptr, 0, 0};
const InputArray* shim_noArray() asm("_ZN2cv7noArrayEv");
const InputArray* shim_noArray() { return &g_noarray; }
// cv::norm(src, NORM_L2, noArray()) of a continuous CV_32F array: normL2_32f = squares accumulated in double in element
// order (the generic normL2Sqr<float, double> loop: v0*v0 + v1*v1 + v2*v2 + v3*v3 per group of four, then one by one), sqrt
double shim_norm(const InputArray* src, int normType, const InputArray* mask) asm("_ZN2cv4normERKNS_11_InputArrayEiS2_");
double shim_norm(const InputArray* src, int normType, const InputArray* mask) {
  const Mat* m = arr_mat(src);
  TRACE("norm type %d", normType);
  if (!m || normType != 4 || (mask && mask->obj) || (m->flags & TYPE_MASK) != 5) __builtin_trap();
  std::vector<float> v;
  for (int i = 0; i < m->rows; ++i)
    for (int j = 0; j < m->cols; ++j) v.push_back(at(m, i, j));
  double s = 0;
  size_t i = 0;
  for (; i + 4 <= v.size(); i += 4) {
    const double v0 = v[i], v1 = v[i + 1], v2 = v[i + 2], v3 = v[i + 3];
    s += v0 * v0 + v1 * v1 + v2 * v2 + v3 * v3;
  }
  for (; i < v.size(); ++i) {
    const double x = v[i];
    s += x * x;
  }
  return std::sqrt(s);
}
// Mat::dot of two continuous CV_32F arrays: dotProd_32f = products and sum in double, in element order (four per group, then
// one by one); Frame::isInFrustum (@0xf5d8d) takes PO.dot(Pn) of two 3x1 vectors
double shim_dot(const Mat* self, const InputArray* other) asm("_ZNK2cv3Mat3dotERKNS_11_InputArrayE");
double shim_dot(const Mat* self, const InputArray* other) {
  const Mat* m = arr_mat(other);
  TRACE("Mat::dot");
  if (!m || (self->flags & TYPE_MASK) != 5 || (m->flags & TYPE_MASK) != 5 || self->rows != m->rows || self->cols != m->cols) __builtin_trap();
  std::vector<float> a, b;
  for (int i = 0; i < m->rows; ++i)
    for (int j = 0; j < m->cols; ++j) { a.push_back(at(self, i, j)); b.push_back(at(m, i, j)); }
  double r = 0;
  size_t i = 0;
  for (; i + 4 <= a.size(); i += 4)
    r += (double)a[i] * b[i] + (double)a[i + 1] * b[i + 1] + (double)a[i + 2] * b[i + 2] + (double)a[i + 3] * b[i + 3];
  for (; i < a.size(); ++i) r += (double)a[i] * b[i];
  return r;
}
// ---- ORBmatcher::SearchBySim3 (@0x838b0): s12 * R12, (1.0 / s12) * R12.t(), -sR21 * t12 ----
void shim_mul_dm(MatExpr* ret, double s, const Mat* a) asm("_ZN2cvmlEdRKNS_3MatE");
void shim_mul_dm(MatExpr* ret, double s, const Mat* a) {
  TRACE("operator*(double %g, Mat)", s);
  expr_init(ret, 'A', 0);
  hdr_copy(&ret->a, a);
  ret->alpha = s;
}
void shim_mul_de(MatExpr* ret, double s, const MatExpr* e) asm("_ZN2cvmlEdRKNS_7MatExprE");
void shim_mul_de(MatExpr* ret, double s, const MatExpr* e) {
  TRACE("operator*(double %g, MatExpr)", s);
  if ((e->flags >> 8) != 'T') __builtin_trap();
  expr_init(ret, 'T', 1);
  hdr_copy(&ret->a, &e->a);
  ret->alpha = e->alpha * s;
}
void shim_neg_m(MatExpr* ret, const Mat* a) asm("_ZN2cvngERKNS_3MatE");
void shim_neg_m(MatExpr* ret, const Mat* a) {
  TRACE("operator-(Mat)");
  expr_init(ret, 'A', 0);
  hdr_copy(&ret->a, a);
  ret->alpha = -1;
}
// MapPoint::GetIndexInKeyFrame replaced for the harness: the index is a field of the faked map point (+0x3f4, -1 = not observed)
int stub_GetIndexInKeyFrame(char* self, char* kf) asm("_ZN9ORB_SLAM28MapPoint18GetIndexInKeyFrameEPNS_8KeyFrameE");
int stub_GetIndexInKeyFrame(char* self, char*) { return *(int*)(self + 0x3f4); }
// ---- ORBmatcher::SearchByProjection(KeyFrame*, cv::Mat Scw, vpPoints, vpMatched, th) (@0x880f0): sRcw / scw, tcw / scw ----
void shim_div_ms(MatExpr* ret, const Mat* a, double s) asm("_ZN2cvdvERKNS_3MatEd");
void shim_div_ms(MatExpr* ret, const Mat* a, double s) {
  TRACE("operator/(Mat, double %g)", s);
  expr_init(ret, 'D', 0);
  hdr_copy(&ret->a, a);
  ret->alpha = 1. / s;
}
// helper for the harness: KeyFrame::mGrid = std::vector<std::vector<std::vector<size_t>>> [cols][rows] built in place from CSR
void refshim_build_kfgrid(void* where, int cols, int rows, const int* start, const int* items);
void refshim_build_kfgrid(void* where, int cols, int rows, const int* start, const int* items) {
  auto* g = new (where) std::vector<std::vector<std::vector<size_t>>>((size_t)cols);
  for (int ix = 0; ix < cols; ++ix) {
    (*g)[ix].resize((size_t)rows);
    for (int iy = 0; iy < rows; ++iy)
      for (int k = start[ix * rows + iy]; k < start[ix * rows + iy + 1]; ++k) (*g)[ix][iy].push_back((size_t)items[k]);
  }
}
// ---- ORBmatcher::Fuse (@0x7a500, @0x7bb20): the map-graph side effects are the reference's CPU control plane; what is pinned is
// the matching decision.  The four functions below REPLACE the library's own (same mangled names; this object precedes the
// library in the global lookup order, and the library calls them through its PLT): they log the call and apply the minimum
// effect later iterations can observe (the key frame's slot, the bad flag, the observation count, "is in this key frame").
static std::vector<long> g_fuse_log;
static char* g_fuse_kf = nullptr;
void refshim_fuse_begin(void* kf) { g_fuse_kf = (char*)kf; g_fuse_log.clear(); }
int refshim_fuse_log(long* out, int cap) {
  const int n = (int)g_fuse_log.size();
  for (int i = 0; i < n && i < cap; ++i) out[i] = g_fuse_log[i];
  return n;
}
void stub_AddObservation(char* self, char* kf, unsigned long idx) asm("_ZN9ORB_SLAM28MapPoint14AddObservationEPNS_8KeyFrameEm");
void stub_AddObservation(char* self, char* kf, unsigned long idx) {
  g_fuse_log.push_back(1); g_fuse_log.push_back((long)self); g_fuse_log.push_back((long)idx);
  const float* ur = *(const float**)(kf + 0x188);
  *(int*)(self + 0x18) += (ur && ur[idx] >= 0) ? 2 : 1;
  self[0x3f0] = 1;  // harness flag: IsInKeyFrame(this key frame)
}
void stub_AddMapPoint(char* kf, char* mp, const unsigned long* idx) asm("_ZN9ORB_SLAM28KeyFrame11AddMapPointEPNS_8MapPointERKm");
void stub_AddMapPoint(char* kf, char* mp, const unsigned long* idx) {
  g_fuse_log.push_back(2); g_fuse_log.push_back((long)mp); g_fuse_log.push_back((long)*idx);
  (*(char***)(kf + 0x520))[*idx] = mp;
}
void stub_Replace(char* self, char* other) asm("_ZN9ORB_SLAM28MapPoint7ReplaceEPS0_");
void stub_Replace(char* self, char* other) {
  g_fuse_log.push_back(3); g_fuse_log.push_back((long)self); g_fuse_log.push_back((long)other);
  self[0x238] = 1;
  if (g_fuse_kf) {
    char** b = *(char***)(g_fuse_kf + 0x520);
    char** e = *(char***)(g_fuse_kf + 0x528);
    for (; b != e; ++b)
      if (*b == self) *b = other;
  }
}
bool stub_IsInKeyFrame(char* self, char* kf) asm("_ZN9ORB_SLAM28MapPoint12IsInKeyFrameEPNS_8KeyFrameE");
bool stub_IsInKeyFrame(char* self, char*) { return self[0x3f0] != 0; }
// helper for the harness (not an OpenCV symbol): a std::set<void*> built in place from an array of pointers
void refshim_build_ptrset(void* where, void* const* ptrs, int n);
void refshim_build_ptrset(void* where, void* const* ptrs, int n) {
  auto* st = new (where) std::set<void*>();
  for (int k = 0; k < n; ++k) st->insert(ptrs[k]);
}
// ---- Frame::UndistortKeyPoints (@0xf8630) / Frame::ComputeImageBounds (@0xf6010): Mat::reshape and cv::undistortPoints ----
void shim_reshape(Mat* ret, const Mat* m, int new_cn, int new_rows) asm("_ZNK2cv3Mat7reshapeEii");
void shim_reshape(Mat* ret, const Mat* m, int new_cn, int new_rows) {
  TRACE("Mat::reshape(%d, %d) of %dx%d type %d", new_cn, new_rows, m->rows, m->cols, m->flags & TYPE_MASK);
  if (new_rows != 0) __builtin_trap();
  hdr_copy(ret, m);
  const int type = m->flags & TYPE_MASK, depth = type & 7, cn = ((type >> 3) & 511) + 1;
  if (new_cn == 0) new_cn = cn;
  const int total_width = m->cols * cn;
  if (total_width % new_cn) __builtin_trap();
  static const size_t depth_size[8] = {1, 1, 2, 2, 4, 4, 8, 2};
  ret->cols = total_width / new_cn;
  ret->flags = (m->flags & ~TYPE_MASK) | depth | ((new_cn - 1) << 3);
  ret->stepbuf[1] = depth_size[depth] * (size_t)new_cn;
}
extern "C" void oracle_undistort_points(const float* calib10, const float* xy, int n, float* out_xy);
// void cv::undistortPoints(InputArray src, OutputArray dst, InputArray cameraMatrix, InputArray distCoeffs, InputArray R, InputArray P)
// as the two callers use it: N x 1 CV_32FC2 in place, K and P the same 3x3 CV_32F matrix, 4 or 5 float coefficients, R empty.
// The arithmetic is the oracle's restatement, pinned bit for bit against cv2 4.13 (tests/test_frame_cpu.py).
void shim_undistortPoints(const InputArray* src, const InputArray* dst, const InputArray* Km, const InputArray* dm, const InputArray* Rm, const InputArray* Pm)
    asm("_ZN2cv15undistortPointsERKNS_11_InputArrayERKNS_12_OutputArrayES2_S2_S2_S2_");
void shim_undistortPoints(const InputArray* src, const InputArray* dst, const InputArray* Km, const InputArray* dm, const InputArray* Rm, const InputArray* Pm) {
  const Mat *s = arr_mat(src), *K = arr_mat(Km), *D = arr_mat(dm), *P = arr_mat(Pm), *R = arr_mat(Rm);
  Mat* d = arr_mat(dst);
  TRACE("undistortPoints %dx%d type %d", s->rows, s->cols, s->flags & TYPE_MASK);
  if (!s || !d || !K || !D || !P || (R && R->data) || (s->flags & TYPE_MASK) != 13 || s->cols != 1) __builtin_trap();
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      if (at(K, r, c) != at(P, r, c)) __builtin_trap();
  const int nd = D->rows * D->cols;
  float calib[10] = {at(K, 0, 0), at(K, 1, 1), at(K, 0, 2), at(K, 1, 2), 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < nd && i < 5; ++i) calib[4 + i] = D->rows == 1 ? at(D, 0, i) : at(D, i, 0);
  if (calib[4] == 0.0f) __builtin_trap();  // both callers test k1 themselves; the oracle entry point copies when k1 == 0
  const int n = s->rows;
  float* in = (float*)::operator new(sizeof(float) * 2 * (size_t)(n ? n : 1));
  float* out = (float*)::operator new(sizeof(float) * 2 * (size_t)(n ? n : 1));
  for (int i = 0; i < n; ++i) std::memcpy(in + 2 * i, s->data + (size_t)i * s->stepp[0], 8);
  oracle_undistort_points(calib, in, n, out);
  mat_create(d, n, 1, 13);
  for (int i = 0; i < n; ++i) std::memcpy(d->data + (size_t)i * d->stepp[0], out + 2 * i, 8);
}
void shim_expr_dtor(MatExpr*) asm("_ZN2cv7MatExprD1Ev");
void shim_expr_dtor(MatExpr*) {}
void shim_expr_dtor2(MatExpr*) asm("_ZN2cv7MatExprD2Ev");
void shim_expr_dtor2(MatExpr*) {}
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
    _build_atomically(["/usr/bin/g++", "-O1", "-std=c++17", "-shared", "-fPIC", src, "-L" + oracle_dir, "-loracle",
                       "-Wl,-rpath," + oracle_dir], os.path.join(out, "libcvshim.so"))
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
    _build_atomically(["gcc", "-shared", os.path.join(out, "stub.s"), "-Wl,-soname,librefstub.so"], os.path.join(out, "librefstub.so"))
    for n in needed:
        _build_atomically(["gcc", "-shared", os.path.join(out, "empty.c"), "-Wl,-soname," + n], os.path.join(out, n))
    return needed


class RefLibrary:
    """lib/libORB_SLAM2.so dlopen'ed over the stubs and shims.

    Process hygiene: the stub and shim objects are loaded RTLD_GLOBAL and define cv:: symbols and operator new/delete.  Any other
    C++ library used LATER in the same process and bound lazily (cv2 above all) may resolve its calls to them.  Render inputs
    with cv2-based code before constructing this object, use it in a dedicated process (tests/golden/make_golden.py reflib), and
    never from the test-suite: the tests only read the fixtures it wrote."""

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
        self._call = getattr(self.lib, "_ZN9ORB_SLAM212ORBextractorclERKN2cv11_InputArrayES4_RSt6vectorINS1_8KeyPointESaIS6_EERKNS1_12_OutputArrayE")
        self._call.argtypes = [C.c_void_p] * 5
        self._call.restype = None

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

    def extract(self, obj, image):
        """ORBextractor::operator()(image, cv::Mat(), keypoints, descriptors) as Frame::ExtractORB calls it (Frame.h:67):
        the reference's own code from pyramid to descriptors, OpenCV entry points supplied by the shims above.
        Returns (keypoints, descriptors [n, 32] uint8)."""
        img, keep = self_mat = RefCode._mat(np.ascontiguousarray(image, np.uint8))
        img[0] = (2 << 32) | (0x42FF0000 | (1 << 14))
        empty = (C.c_uint64 * 12)()
        empty[0] = 0x42FF0000
        empty[8], empty[9] = C.addressof(empty) + 8, C.addressof(empty) + 0x50
        desc = (C.c_uint64 * 12)()
        desc[0] = 0x42FF0000
        desc[8], desc[9] = C.addressof(desc) + 8, C.addressof(desc) + 0x50
        arr = lambda flags, m: (C.c_uint64 * 3)(flags, C.addressof(m), 0)
        a_img, a_mask, a_desc = arr(0x01010000, img), arr(0x01010000, empty), arr(0x02010000, desc)
        kv = (C.c_uint64 * 3)()
        self._call(C.addressof(obj), C.addressof(a_img), C.addressof(a_mask), C.addressof(kv), C.addressof(a_desc))
        n = (kv[1] - kv[0]) // 28
        kps = np.ctypeslib.as_array(C.cast(kv[0], C.POINTER(C.c_uint8)), (n * 28,)).view(self.KP).copy() if n else np.empty(0, self.KP)
        rows, cols = desc[1] & 0xffffffff, desc[1] >> 32
        d = np.ctypeslib.as_array(C.cast(desc[2], C.POINTER(C.c_uint8)), (rows * cols,)).reshape(rows, cols).copy() if rows else np.empty((0, 32), np.uint8)
        return kps, d

    # ---- Frame::AssignFeaturesToGrid (@0xf9120), Frame::PosInGrid (@0xf5fa0), Frame::GetFeaturesInArea (@0xfbc60) ----
    # A Frame is faked as zeroed memory with the three members these functions read: N @0xec, mvKeysUn @0x120 (vector of
    # 28-byte cv::KeyPoint), mGrid @0x2c8 (std::vector<size_t>[64][48]); offsets from the disassembly of GetFeaturesInArea.
    def frame_grid_queries(self, kps_un, bounds, queries):
        """bounds = (mnMinX, mnMaxX, mnMinY, mnMaxY); queries = rows (x, y, r, minLevel, maxLevel).
        Returns (grid CSR start [64*48+1], items, [candidate index arrays per query])."""
        f32 = np.float32
        st = lambda name: C.c_float.in_dll(self.lib, name)
        minX, maxX, minY, maxY = (f32(v) for v in bounds)
        st("_ZN9ORB_SLAM25Frame6mnMinXE").value, st("_ZN9ORB_SLAM25Frame6mnMaxXE").value = minX, maxX
        st("_ZN9ORB_SLAM25Frame6mnMinYE").value, st("_ZN9ORB_SLAM25Frame6mnMaxYE").value = minY, maxY
        st("_ZN9ORB_SLAM25Frame21mfGridElementWidthInvE").value = f32(64.0) / (maxX - minX)    # Frame ctor, @0xfa2e8
        st("_ZN9ORB_SLAM25Frame22mfGridElementHeightInvE").value = f32(48.0) / (maxY - minY)
        keys = np.ascontiguousarray(kps_un, self.KP)
        frame = (C.c_uint64 * (0x13000 // 8))()
        base = C.addressof(frame)
        C.c_int32.from_address(base + 0xec).value = len(keys)
        frame[0x120 // 8], frame[0x128 // 8], frame[0x130 // 8] = keys.ctypes.data, keys.ctypes.data + keys.nbytes, keys.ctypes.data + keys.nbytes
        assign = getattr(self.lib, "_ZN9ORB_SLAM25Frame20AssignFeaturesToGridEv")
        assign.argtypes, assign.restype = [C.c_void_p], None
        assign(base)
        start, items = np.zeros(64 * 48 + 1, np.int32), []
        for c in range(64 * 48):
            b, e = frame[(0x2c8 + 24 * c) // 8], frame[(0x2c8 + 24 * c) // 8 + 1]
            n = (e - b) // 8
            if n:
                items += list(np.ctypeslib.as_array(C.cast(b, C.POINTER(C.c_uint64)), (n,)))
            start[c + 1] = len(items)
        area = getattr(self.lib, "_ZNK9ORB_SLAM25Frame17GetFeaturesInAreaERKfS2_S2_ii")
        area.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        area.restype = C.c_void_p
        res = []
        for x, y, r, lo, hi in queries:
            ret = (C.c_uint64 * 3)()
            fx, fy, fr = C.c_float(x), C.c_float(y), C.c_float(r)
            area(C.addressof(ret), base, C.addressof(fx), C.addressof(fy), C.addressof(fr), int(lo), int(hi))
            n = (ret[1] - ret[0]) // 8
            res.append(np.ctypeslib.as_array(C.cast(ret[0], C.POINTER(C.c_uint64)), (n,)).astype(np.int32) if n else np.empty(0, np.int32))
        return start, np.array(items, np.int32), res

    # ---- ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, float th) (@0x79f10) on faked objects ----
    # Offsets read from its disassembly and from MapPoint::isBad / Observations / GetDescriptor:
    #   MapPoint: nObs @0x18, mTrackProjX/Y/XR @0x1c/0x20/0x24, mnTrackScaleLevel @0x28, mTrackViewCos @0x2c, mbTrackInView @0x30,
    #             mDescriptor (cv::Mat) @0x1c8, mbBad @0x238, mutexes (zero bytes = unlocked)
    #   Frame:    N @0xec, mvKeysUn @0x120, mvuRight @0x138, mDescriptors (cv::Mat) @0x1c8, mvpMapPoints @0x288, mGrid @0x2c8,
    #             mvScaleFactors @0x12348
    @staticmethod
    def _mat_at(addr, arr):
        m = (C.c_uint64 * 12).from_address(addr)
        m[0] = (2 << 32) | (0x42FF0000 | (1 << 14))
        m[1] = (arr.shape[1] << 32) | arr.shape[0]
        m[2] = m[3] = arr.ctypes.data
        m[4] = m[5] = arr.ctypes.data + arr.nbytes
        m[6] = m[7] = 0
        m[8], m[9] = addr + 8, addr + 0x50
        m[10], m[11] = arr.strides[0], 1

    def search_local_points(self, mp, fr, cam4, scale_factors, th, nnratio):
        """Inputs in the layout of tests/matchdata.py: local_points_case.  Returns (match_f int32 [N] = map-point index assigned
        to each keypoint or -1, nmatches) as oracle.search_local_points does."""
        f32 = np.float32
        st = lambda name: C.c_float.in_dll(self.lib, name)
        st("_ZN9ORB_SLAM25Frame6mnMinXE").value, st("_ZN9ORB_SLAM25Frame6mnMinYE").value = f32(cam4[0]), f32(cam4[1])
        st("_ZN9ORB_SLAM25Frame21mfGridElementWidthInvE").value = f32(cam4[2])
        st("_ZN9ORB_SLAM25Frame22mfGridElementHeightInvE").value = f32(cam4[3])
        M, N = len(mp["desc"]), len(fr["desc"])
        keys = np.zeros(N, self.KP)
        keys["x"], keys["y"], keys["octave"] = fr["xy"][:, 0], fr["xy"][:, 1], fr["octave"]
        uright = np.ascontiguousarray(fr["uright"], np.float32)
        fdesc = np.ascontiguousarray(fr["desc"], np.uint8)
        mdesc = np.ascontiguousarray(mp["desc"], np.uint8)
        sf = np.ascontiguousarray(scale_factors, np.float32)
        mps = (C.c_uint8 * (0x400 * (M + 1)))()          # M map points + one "already tracked" point for fr["taken"]
        mbase = C.addressof(mps)
        for i in range(M + 1):
            a = mbase + 0x400 * i
            if i < M:
                C.c_int32.from_address(a + 0x18).value = 1 if mp["obs"][i] else 0
                C.c_float.from_address(a + 0x1c).value = f32(mp["proj"][i, 0])
                C.c_float.from_address(a + 0x20).value = f32(mp["proj"][i, 1])
                C.c_float.from_address(a + 0x24).value = f32(mp["proj"][i, 2])
                C.c_int32.from_address(a + 0x28).value = int(mp["level"][i])
                C.c_float.from_address(a + 0x2c).value = f32(mp["viewcos"][i])
                C.c_uint8.from_address(a + 0x30).value = 1 if mp["valid"][i] else 0
                self._mat_at(a + 0x1c8, mdesc[i:i + 1])
            else:
                C.c_int32.from_address(a + 0x18).value = 1
        vp = np.array([mbase + 0x400 * i for i in range(M)], np.uint64)
        fmp = np.array([mbase + 0x400 * M if t else 0 for t in fr["taken"]], np.uint64)
        frame = (C.c_uint64 * (0x12800 // 8))()
        fb = C.addressof(frame)
        C.c_int32.from_address(fb + 0xec).value = N
        setv = lambda off, arr: (frame.__setitem__(off // 8, arr.ctypes.data), frame.__setitem__(off // 8 + 1, arr.ctypes.data + arr.nbytes),
                                 frame.__setitem__(off // 8 + 2, arr.ctypes.data + arr.nbytes))
        setv(0x120, keys); setv(0x138, uright); setv(0x288, fmp); setv(0x12348, sf)
        self._mat_at(fb + 0x1c8, fdesc)
        assign = getattr(self.lib, "_ZN9ORB_SLAM25Frame20AssignFeaturesToGridEv")
        assign.argtypes, assign.restype = [C.c_void_p], None
        assign(fb)
        fn = getattr(self.lib, "_ZN9ORB_SLAM210ORBmatcher18SearchByProjectionERNS_5FrameERKSt6vectorIPNS_8MapPointESaIS5_EEf")
        fn.argtypes, fn.restype = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float], C.c_int
        matcher = (C.c_uint8 * 8)()
        C.c_float.from_address(C.addressof(matcher)).value = f32(nnratio)
        matcher[4] = 1
        vec = (C.c_uint64 * 3)(vp.ctypes.data, vp.ctypes.data + vp.nbytes, vp.ctypes.data + vp.nbytes)
        n = fn(C.addressof(matcher), fb, C.addressof(vec), f32(th))
        out = np.full(N, -1, np.int32)
        for i in range(N):
            p = int(fmp[i])
            if p and p != mbase + 0x400 * M:
                out[i] = (p - mbase) // 0x400
        return out, int(n)

    # ---- ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches) (@0x80150) ----
    # KeyFrame: mvKeysUn @0x170, mDescriptors @0x1b8, mFeatVec @0x248, mvpMapPoints @0x520 (KeyFrame::GetMapPointMatches
    # @0x9c4b0 copies it under the mutex @0x690).  Frame: N @0xec, mvKeys @0xf0, mFeatVec @0x198, mDescriptors @0x1c8.
    def search_by_bow(self, kf, f, nnratio=0.7, check_ori=True):
        """Inputs in the layout of oracle.search_by_bow.  Returns (match_f int32 [N2] = KF feature matched to each F feature or
        -1, nmatches)."""
        f32 = np.float32
        n1, n2 = len(kf["desc"]), len(f["desc"])
        k1, k2 = np.zeros(n1, self.KP), np.zeros(n2, self.KP)
        k1["angle"], k2["angle"] = kf["angle"], f["angle"]
        d1, d2 = np.ascontiguousarray(kf["desc"], np.uint8), np.ascontiguousarray(f["desc"], np.uint8)
        mps = (C.c_uint8 * (0x400 * max(n1, 1)))()   # one zeroed (good, not bad) MapPoint per valid KF feature
        mb = C.addressof(mps)
        vp = np.array([mb + 0x400 * i if kf["valid"][i] else 0 for i in range(n1)], np.uint64)
        kfo = (C.c_uint64 * (0x800 // 8))()
        kb = C.addressof(kfo)
        fro = (C.c_uint64 * (0x400 // 8))()
        fb = C.addressof(fro)
        def setv(obj, off, arr):
            obj[off // 8], obj[off // 8 + 1], obj[off // 8 + 2] = arr.ctypes.data, arr.ctypes.data + arr.nbytes, arr.ctypes.data + arr.nbytes
        setv(kfo, 0x170, k1); setv(kfo, 0x520, vp)
        self._mat_at(kb + 0x1b8, d1)
        C.c_int32.from_address(fb + 0xec).value = n2
        setv(fro, 0xf0, k2)
        self._mat_at(fb + 0x1c8, d2)
        build = self._shims.refshim_build_featvec
        build.argtypes, build.restype = [C.c_void_p] * 4 + [C.c_int], None
        arrs = [np.ascontiguousarray(x, np.int32) for x in (kf["nodes"], kf["start"], kf["idx"], f["nodes"], f["start"], f["idx"])]
        build(kb + 0x248, arrs[0].ctypes.data, arrs[1].ctypes.data, arrs[2].ctypes.data, len(arrs[0]))
        build(fb + 0x198, arrs[3].ctypes.data, arrs[4].ctypes.data, arrs[5].ctypes.data, len(arrs[3]))
        fn = getattr(self.lib, "_ZN9ORB_SLAM210ORBmatcher11SearchByBoWEPNS_8KeyFrameERNS_5FrameERSt6vectorIPNS_8MapPointESaIS7_EE")
        fn.argtypes, fn.restype = [C.c_void_p] * 4, C.c_int
        matcher = (C.c_uint8 * 8)()
        C.c_float.from_address(C.addressof(matcher)).value = f32(nnratio)
        matcher[4] = 1 if check_ori else 0
        res = (C.c_uint64 * 3)()
        n = fn(C.addressof(matcher), kb, fb, C.addressof(res))
        cnt = (res[1] - res[0]) // 8
        ptrs = np.ctypeslib.as_array(C.cast(res[0], C.POINTER(C.c_uint64)), (cnt,)) if cnt else np.empty(0, np.uint64)
        out = np.full(n2, -1, np.int32)
        for i, p in enumerate(ptrs):
            if p:
                out[i] = (int(p) - mb) // 0x400
        return out, int(n)

    # ---- ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12) (@0x82cc0, loop closing) ----
    def search_by_bow_kfkf(self, kf1, kf2, nnratio=0.75, check_ori=True):
        """kf1 / kf2: dict(desc, angle, valid, nodes, start, idx) (valid = map point exists and is not bad).  Returns
        (match12 int32 [N1] = KF2 feature whose map point was assigned to KF1 feature i1, or -1; nmatches)."""
        f32 = np.float32
        keep = []
        def make(kf):
            n = len(kf["desc"])
            k = np.zeros(n, self.KP)
            k["angle"] = kf["angle"]
            d = np.ascontiguousarray(kf["desc"], np.uint8)
            mps = (C.c_uint8 * (0x400 * max(n, 1)))()
            mb = C.addressof(mps)
            vp = np.array([mb + 0x400 * i if kf["valid"][i] else 0 for i in range(n)], np.uint64)
            o = (C.c_uint64 * (0x800 // 8))()
            b = C.addressof(o)
            o[0x170 // 8], o[0x170 // 8 + 1], o[0x170 // 8 + 2] = k.ctypes.data, k.ctypes.data + k.nbytes, k.ctypes.data + k.nbytes
            o[0x520 // 8], o[0x520 // 8 + 1], o[0x520 // 8 + 2] = vp.ctypes.data, vp.ctypes.data + vp.nbytes, vp.ctypes.data + vp.nbytes
            self._mat_at(b + 0x1b8, d)
            arrs = [np.ascontiguousarray(kf[x], np.int32) for x in ("nodes", "start", "idx")]
            build = self._shims.refshim_build_featvec
            build.argtypes, build.restype = [C.c_void_p] * 4 + [C.c_int], None
            build(b + 0x248, arrs[0].ctypes.data, arrs[1].ctypes.data, arrs[2].ctypes.data, len(arrs[0]))
            keep.extend([k, d, mps, vp, o, arrs])
            return b, mb, n
        b1, mb1, n1 = make(kf1)
        b2, mb2, n2 = make(kf2)
        fn = getattr(self.lib, "_ZN9ORB_SLAM210ORBmatcher11SearchByBoWEPNS_8KeyFrameES2_RSt6vectorIPNS_8MapPointESaIS5_EE")
        fn.argtypes, fn.restype = [C.c_void_p] * 4, C.c_int
        matcher = (C.c_uint8 * 8)()
        C.c_float.from_address(C.addressof(matcher)).value = f32(nnratio)
        matcher[4] = 1 if check_ori else 0
        res = (C.c_uint64 * 3)()
        n = fn(C.addressof(matcher), b1, b2, C.addressof(res))
        cnt = (res[1] - res[0]) // 8
        ptrs = np.ctypeslib.as_array(C.cast(res[0], C.POINTER(C.c_uint64)), (cnt,)) if cnt else np.empty(0, np.uint64)
        out = np.full(n1, -1, np.int32)
        for i, p in enumerate(ptrs):
            if p:
                out[i] = (int(p) - mb2) // 0x400
        return out, int(n)

    # ---- ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, float th, bool bMono) (@0x80d00) ----
    # Further Frame members: mbf @0xe0, mb @0xe4, mvKeys @0xf0, mvbOutlier (vector<bool>) @0x2a0, mTcw (cv::Mat 4x4) @0x122c8;
    # MapPoint::mWorldPos (cv::Mat 3x1) @0xd8 (MapPoint::GetWorldPos @0x91630).  Frame::fx/fy/cx/cy are static members.
    @staticmethod
    def _fmat_at(addr, arr):
        m = (C.c_uint64 * 12).from_address(addr)
        m[0] = (2 << 32) | (0x42FF0000 | (1 << 14) | 5)
        m[1] = (arr.shape[1] << 32) | arr.shape[0]
        m[2] = m[3] = arr.ctypes.data
        m[4] = m[5] = arr.ctypes.data + arr.nbytes
        m[6] = m[7] = 0
        m[8], m[9] = addr + 8, addr + 0x50
        m[10], m[11] = arr.strides[0], 4

    def search_by_projection(self, last, cur, cam, scale_factors, tcw_cur, tcw_last, th, mono=False, check_ori=True):
        """Inputs in the layout of oracle.search_by_projection (tests/matchdata.py: projection_case).  Returns (match_cur int32
        [N2] = last-frame index assigned to each current keypoint or -1, nmatches)."""
        f32 = np.float32
        st = lambda name: C.c_float.in_dll(self.lib, name)
        for name, v in zip(("2fx", "2fy", "2cx", "2cy"), cam[:4]):
            st("_ZN9ORB_SLAM25Frame%sE" % name).value = f32(v)
        st("_ZN9ORB_SLAM25Frame6mnMinXE").value, st("_ZN9ORB_SLAM25Frame6mnMaxXE").value = f32(cam[6]), f32(cam[7])
        st("_ZN9ORB_SLAM25Frame6mnMinYE").value, st("_ZN9ORB_SLAM25Frame6mnMaxYE").value = f32(cam[8]), f32(cam[9])
        st("_ZN9ORB_SLAM25Frame21mfGridElementWidthInvE").value = f32(cam[10])
        st("_ZN9ORB_SLAM25Frame22mfGridElementHeightInvE").value = f32(cam[11])
        n1, n2 = len(last["desc"]), len(cur["desc"])
        ldesc, cdesc = np.ascontiguousarray(last["desc"], np.uint8), np.ascontiguousarray(cur["desc"], np.uint8)
        xyz = np.ascontiguousarray(last["xyz"], np.float32)
        mps = (C.c_uint8 * (0x400 * (n1 + 1)))()
        mb_ = C.addressof(mps)
        for i in range(n1):
            a = mb_ + 0x400 * i
            C.c_int32.from_address(a + 0x18).value = 1 if last["obs"][i] else 0
            self._fmat_at(a + 0xd8, xyz[i].reshape(3, 1))
            self._mat_at(a + 0x1c8, ldesc[i:i + 1])
        C.c_int32.from_address(mb_ + 0x400 * n1 + 0x18).value = 1            # the "already tracked" point of cur["taken"]
        lmp = np.array([mb_ + 0x400 * i if last["valid"][i] else 0 for i in range(n1)], np.uint64)
        cmp_ = np.array([mb_ + 0x400 * n1 if t else 0 for t in cur["taken"]], np.uint64)
        lk = np.zeros(n1, self.KP)
        lk["octave"], lk["angle"] = last["octave"], last["angle"]
        ck = np.zeros(n2, self.KP)
        ck["x"], ck["y"], ck["octave"], ck["angle"] = cur["xy"][:, 0], cur["xy"][:, 1], cur["octave"], cur["angle"]
        uright = np.ascontiguousarray(cur["uright"], np.float32)
        sf = np.ascontiguousarray(scale_factors, np.float32)
        T = lambda t: np.ascontiguousarray(np.vstack([np.asarray(t, np.float32).reshape(3, 4), [[0, 0, 0, 1]]]).astype(np.float32))
        Tc, Tl = T(tcw_cur), T(tcw_last)
        outl = np.zeros((n1 + 63) // 64 + 1, np.uint64)
        fl, fc = (C.c_uint64 * (0x12800 // 8))(), (C.c_uint64 * (0x12800 // 8))()
        lb, cb = C.addressof(fl), C.addressof(fc)
        def setv(obj, off, arr):
            obj[off // 8], obj[off // 8 + 1], obj[off // 8 + 2] = arr.ctypes.data, arr.ctypes.data + arr.nbytes, arr.ctypes.data + arr.nbytes
        C.c_int32.from_address(lb + 0xec).value = n1
        setv(fl, 0xf0, lk); setv(fl, 0x120, lk); setv(fl, 0x288, lmp)
        fl[0x2a0 // 8] = outl.ctypes.data                                      # vector<bool>::_M_start._M_p, all bits clear
        self._fmat_at(lb + 0x122c8, Tl)
        C.c_int32.from_address(cb + 0xec).value = n2
        C.c_float.from_address(cb + 0xe0).value, C.c_float.from_address(cb + 0xe4).value = f32(cam[4]), f32(cam[5])
        setv(fc, 0xf0, ck); setv(fc, 0x120, ck); setv(fc, 0x138, uright); setv(fc, 0x288, cmp_); setv(fc, 0x12348, sf)
        self._mat_at(cb + 0x1c8, cdesc)
        self._fmat_at(cb + 0x122c8, Tc)
        assign = getattr(self.lib, "_ZN9ORB_SLAM25Frame20AssignFeaturesToGridEv")
        assign.argtypes, assign.restype = [C.c_void_p], None
        assign(cb)
        fn = getattr(self.lib, "_ZN9ORB_SLAM210ORBmatcher18SearchByProjectionERNS_5FrameERKS1_fb")
        fn.argtypes, fn.restype = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_bool], C.c_int
        matcher = (C.c_uint8 * 8)()
        C.c_float.from_address(C.addressof(matcher)).value = f32(0.9)
        matcher[4] = 1 if check_ori else 0
        n = fn(C.addressof(matcher), cb, lb, f32(th), bool(mono))
        out = np.full(n2, -1, np.int32)
        for i in range(n2):
            p = int(cmp_[i])
            if p and p != mb_ + 0x400 * n1:
                out[i] = (p - mb_) // 0x400
        return out, int(n)

    # ---- ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, th, ORBdist) ----
    # (@0x7e8c0, Tracking::Relocalization).  Further members: MapPoint::mbBad @0x238, mfMinDistance @0x248, mfMaxDistance @0x24c;
    # Frame::mnScaleLevels @0x12338, mfLogScaleFactor @0x12340; KeyFrame::mvpMapPoints @0x520 (GetMapPointMatches, mutex @0x690).
    def search_by_projection_kf(self, kf, cur, cam, scale_factors, log_scale_factor, tcw_cur, th, orb_dist, check_ori=True):
        """kf: dict(state uint8 [M] (0 = no map point, 1 = good, 2 = bad, 3 = in sAlreadyFound), xyz [M,3], desc [M,32],
        dist_range [M,2] (mfMinDistance, mfMaxDistance), angle [M]); cur: dict(xy, octave, angle, desc, taken (mvpMapPoints[i] != NULL));
        cam = (fx, fy, cx, cy, mnMinX, mnMaxX, mnMinY, mnMaxY, gridWInv, gridHInv).  Returns (match_cur int32 [N2], nmatches)."""
        f32 = np.float32
        st = lambda name: C.c_float.in_dll(self.lib, name)
        for name, v in zip(("2fx", "2fy", "2cx", "2cy"), cam[:4]):
            st("_ZN9ORB_SLAM25Frame%sE" % name).value = f32(v)
        st("_ZN9ORB_SLAM25Frame6mnMinXE").value, st("_ZN9ORB_SLAM25Frame6mnMaxXE").value = f32(cam[4]), f32(cam[5])
        st("_ZN9ORB_SLAM25Frame6mnMinYE").value, st("_ZN9ORB_SLAM25Frame6mnMaxYE").value = f32(cam[6]), f32(cam[7])
        st("_ZN9ORB_SLAM25Frame21mfGridElementWidthInvE").value = f32(cam[8])
        st("_ZN9ORB_SLAM25Frame22mfGridElementHeightInvE").value = f32(cam[9])
        m, n2 = len(kf["desc"]), len(cur["desc"])
        kdesc, cdesc = np.ascontiguousarray(kf["desc"], np.uint8), np.ascontiguousarray(cur["desc"], np.uint8)
        xyz = np.ascontiguousarray(kf["xyz"], np.float32)
        rng_ = np.ascontiguousarray(kf["dist_range"], np.float32)
        mps = (C.c_uint8 * (0x400 * (m + 1)))()
        mb_ = C.addressof(mps)
        for i in range(m):
            a = mb_ + 0x400 * i
            self._fmat_at(a + 0xd8, xyz[i].reshape(3, 1))
            self._mat_at(a + 0x1c8, kdesc[i:i + 1])
            C.c_uint8.from_address(a + 0x238).value = 1 if kf["state"][i] == 2 else 0
            C.c_float.from_address(a + 0x248).value = rng_[i, 0]
            C.c_float.from_address(a + 0x24c).value = rng_[i, 1]
        kmp = np.array([mb_ + 0x400 * i if kf["state"][i] else 0 for i in range(m)], np.uint64)
        found = np.array([mb_ + 0x400 * i for i in range(m) if kf["state"][i] == 3], np.uint64)
        cmp_ = np.array([mb_ + 0x400 * m if t else 0 for t in cur["taken"]], np.uint64)
        kk = np.zeros(m, self.KP)
        kk["angle"] = kf["angle"]
        ck = np.zeros(n2, self.KP)
        ck["x"], ck["y"], ck["octave"], ck["angle"] = cur["xy"][:, 0], cur["xy"][:, 1], cur["octave"], cur["angle"]
        sf = np.ascontiguousarray(scale_factors, np.float32)
        Tc = np.ascontiguousarray(np.vstack([np.asarray(tcw_cur, np.float32).reshape(3, 4), [[0, 0, 0, 1]]]).astype(np.float32))
        fk, fc = (C.c_uint64 * (0x800 // 8))(), (C.c_uint64 * (0x12800 // 8))()
        kb, cb = C.addressof(fk), C.addressof(fc)
        def setv(obj, off, arr):
            obj[off // 8], obj[off // 8 + 1], obj[off // 8 + 2] = arr.ctypes.data, arr.ctypes.data + arr.nbytes, arr.ctypes.data + arr.nbytes
        setv(fk, 0x170, kk); setv(fk, 0x520, kmp)
        C.c_int32.from_address(cb + 0xec).value = n2
        setv(fc, 0xf0, ck); setv(fc, 0x120, ck); setv(fc, 0x288, cmp_); setv(fc, 0x12348, sf)
        C.c_int32.from_address(cb + 0x12338).value = len(sf)
        C.c_float.from_address(cb + 0x12340).value = f32(log_scale_factor)
        self._mat_at(cb + 0x1c8, cdesc)
        self._fmat_at(cb + 0x122c8, Tc)
        assign = getattr(self.lib, "_ZN9ORB_SLAM25Frame20AssignFeaturesToGridEv")
        assign.argtypes, assign.restype = [C.c_void_p], None
        assign(cb)
        pset = (C.c_uint64 * 8)()
        bs = self._shims.refshim_build_ptrset
        bs.argtypes, bs.restype = [C.c_void_p, C.c_void_p, C.c_int], None
        bs(C.addressof(pset), found.ctypes.data if len(found) else None, len(found))
        fn = getattr(self.lib, "_ZN9ORB_SLAM210ORBmatcher18SearchByProjectionERNS_5FrameEPNS_8KeyFrameERKSt3setIPNS_8MapPointESt4lessIS7_ESaIS7_EEfi")
        fn.argtypes, fn.restype = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int], C.c_int
        matcher = (C.c_uint8 * 8)()
        C.c_float.from_address(C.addressof(matcher)).value = f32(0.9)
        matcher[4] = 1 if check_ori else 0
        n = fn(C.addressof(matcher), cb, kb, C.addressof(pset), f32(th), int(orb_dist))
        out = np.full(n2, -1, np.int32)
        for i in range(n2):
            p = int(cmp_[i])
            if p and p != mb_ + 0x400 * m:
                out[i] = (p - mb_) // 0x400
        return out, int(n)

    # ---- Frame::isInFrustum(MapPoint* pMP, float viewingCosLimit) (@0xf5190; Tracking::SearchLocalPoints calls it per local map point) ----
    # Frame: mRcw (3x3) @0x123a8, mtcw (3x1) @0x12408, mOw (3x1) @0x124c8, mbf @0xe0, mnScaleLevels @0x12338, mfLogScaleFactor @0x12340;
    # MapPoint: mWorldPos @0xd8, mNormalVector @0x168, mfMinDistance @0x248, mfMaxDistance @0x24c; outputs mTrackProjX @0x1c,
    # mTrackProjY @0x20, mTrackProjXR @0x24, mnTrackScaleLevel @0x28, mTrackViewCos @0x2c, mbTrackInView @0x30.
    def is_in_frustum(self, xyz, normal, dist_range, cam, tcw, ow, mbf, log_scale_factor, n_levels, cos_limit):
        """cam = (fx, fy, cx, cy, mnMinX, mnMaxX, mnMinY, mnMaxY).  Returns dict(in_view uint8 [M], proj float32 [M,3] (u, v, ur),
        level int32 [M], viewcos float32 [M]) as the reference's function leaves them in the map points (untouched fields read 0)."""
        f32 = np.float32
        st = lambda name: C.c_float.in_dll(self.lib, name)
        for name, v in zip(("2fx", "2fy", "2cx", "2cy"), cam[:4]):
            st("_ZN9ORB_SLAM25Frame%sE" % name).value = f32(v)
        st("_ZN9ORB_SLAM25Frame6mnMinXE").value, st("_ZN9ORB_SLAM25Frame6mnMaxXE").value = f32(cam[4]), f32(cam[5])
        st("_ZN9ORB_SLAM25Frame6mnMinYE").value, st("_ZN9ORB_SLAM25Frame6mnMaxYE").value = f32(cam[6]), f32(cam[7])
        m = len(xyz)
        xyz = np.ascontiguousarray(xyz, np.float32); normal = np.ascontiguousarray(normal, np.float32)
        rng_ = np.ascontiguousarray(dist_range, np.float32)
        T = np.asarray(tcw, np.float32).reshape(3, 4)
        R = np.ascontiguousarray(T[:, :3]); t = np.ascontiguousarray(T[:, 3:4]); O = np.ascontiguousarray(np.asarray(ow, np.float32).reshape(3, 1))
        fr = (C.c_uint64 * (0x12800 // 8))()
        fb = C.addressof(fr)
        self._fmat_at(fb + 0x123a8, R); self._fmat_at(fb + 0x12408, t); self._fmat_at(fb + 0x124c8, O)
        C.c_float.from_address(fb + 0xe0).value = f32(mbf)
        C.c_int32.from_address(fb + 0x12338).value = int(n_levels)
        C.c_float.from_address(fb + 0x12340).value = f32(log_scale_factor)
        fn = getattr(self.lib, "_ZN9ORB_SLAM25Frame11isInFrustumEPNS_8MapPointEf")
        fn.argtypes, fn.restype = [C.c_void_p, C.c_void_p, C.c_float], C.c_bool
        out = dict(in_view=np.zeros(m, np.uint8), proj=np.zeros((m, 3), np.float32), level=np.zeros(m, np.int32), viewcos=np.zeros(m, np.float32))
        mp = (C.c_uint8 * 0x400)()
        a = C.addressof(mp)
        for i in range(m):
            C.memset(a, 0, 0x400)
            self._fmat_at(a + 0xd8, xyz[i].reshape(3, 1))
            self._fmat_at(a + 0x168, normal[i].reshape(3, 1))
            C.c_float.from_address(a + 0x248).value = rng_[i, 0]
            C.c_float.from_address(a + 0x24c).value = rng_[i, 1]
            r = fn(fb, a, f32(cos_limit))
            assert bool(r) == bool(mp[0x30])
            out["in_view"][i] = mp[0x30]
            out["proj"][i] = (C.c_float.from_address(a + 0x1c).value, C.c_float.from_address(a + 0x20).value, C.c_float.from_address(a + 0x24).value)
            out["level"][i] = C.c_int32.from_address(a + 0x28).value
            out["viewcos"][i] = C.c_float.from_address(a + 0x2c).value
        return out

    # ---- ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, vector<MapPoint*>& vpMatched, int th) ----
    # (@0x880f0, LoopClosing::ComputeSim3 / SearchAndFuse).  KeyFrame members: mnGridCols @0x18, mnGridRows @0x1c,
    # mfGridElementWidthInv @0x20, mfGridElementHeightInv @0x24, fx fy cx cy @0x130.., N @0x154, mvKeysUn @0x170, mDescriptors @0x1b8,
    # mnScaleLevels @0x2d8, mfLogScaleFactor @0x2e0, mvScaleFactors @0x2e8, mnMinX mnMinY mnMaxX mnMaxY (int) @0x330.., mGrid @0x548
    # (KeyFrame::GetFeaturesInArea @0x96fe0, IsInImage @0x97480, MapPoint::PredictScale(dist, KeyFrame*) @0x8fb60).
    def make_keyframe(self, kf, keep):
        """kf: dict(xy, octave, desc, grid_start, grid_items, cam4, bounds4 (int: minx, miny, maxx, maxy), gwi, ghi, scale_factors,
        log_sf [, angle, uright]) -> address of a faked KeyFrame."""
        n = len(kf["desc"])
        k = np.zeros(n, self.KP)
        k["x"], k["y"], k["octave"] = kf["xy"][:, 0], kf["xy"][:, 1], kf["octave"]
        if "angle" in kf:
            k["angle"] = kf["angle"]
        d = np.ascontiguousarray(kf["desc"], np.uint8)
        sf = np.ascontiguousarray(kf["scale_factors"], np.float32)
        o = (C.c_uint64 * (0x800 // 8))()
        b = C.addressof(o)
        def setv(off, arr):
            o[off // 8], o[off // 8 + 1], o[off // 8 + 2] = arr.ctypes.data, arr.ctypes.data + arr.nbytes, arr.ctypes.data + arr.nbytes
        C.c_int32.from_address(b + 0x18).value, C.c_int32.from_address(b + 0x1c).value = 64, 48
        C.c_float.from_address(b + 0x20).value, C.c_float.from_address(b + 0x24).value = np.float32(kf["gwi"]), np.float32(kf["ghi"])
        for i in range(4):
            C.c_float.from_address(b + 0x130 + 4 * i).value = np.float32(kf["cam4"][i])
            C.c_int32.from_address(b + 0x330 + 4 * i).value = int(kf["bounds4"][i])
        C.c_int32.from_address(b + 0x154).value = n
        setv(0x170, k); setv(0x2e8, sf)
        self._mat_at(b + 0x1b8, d)
        C.c_int32.from_address(b + 0x2d8).value = len(sf)
        C.c_float.from_address(b + 0x2e0).value = np.float32(kf["log_sf"])
        gs, gi = np.ascontiguousarray(kf["grid_start"], np.int32), np.ascontiguousarray(kf["grid_items"], np.int32)
        bg = self._shims.refshim_build_kfgrid
        bg.argtypes, bg.restype = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p], None
        bg(b + 0x548, 64, 48, gs.ctypes.data, gi.ctypes.data)
        keep.extend([k, d, sf, o, gs, gi])
        if "uright" in kf:   # the members ORBmatcher::Fuse reads on top: mbf @0x148, mvuRight @0x188, mvInvLevelSigma2 @0x318, Tcw @0x3a0,
            ur = np.ascontiguousarray(kf["uright"], np.float32)           # Ow @0x460, mvpMapPoints @0x520
            inv = np.ascontiguousarray(kf["inv_level_sigma2"], np.float32)
            T = np.ascontiguousarray(np.vstack([np.asarray(kf["tcw"], np.float32).reshape(3, 4), [[0, 0, 0, 1]]]).astype(np.float32))
            O = np.ascontiguousarray(np.asarray(kf["ow"], np.float32).reshape(3, 1))
            C.c_float.from_address(b + 0x148).value = np.float32(kf["mbf"])
            setv(0x188, ur); setv(0x318, inv)
            self._fmat_at(b + 0x3a0, T); self._fmat_at(b + 0x460, O)
            keep.extend([ur, inv, T, O])
        return b

    def make_map_points(self, mp, keep):
        """mp: dict(state (1 good, 2 bad), xyz, normal, dist_range, desc) -> (base address of M faked MapPoints 0x400 bytes apart, buffer)."""
        m = len(mp["desc"])
        xyz = np.ascontiguousarray(mp["xyz"], np.float32); nrm = np.ascontiguousarray(mp["normal"], np.float32)
        rng_ = np.ascontiguousarray(mp["dist_range"], np.float32); desc = np.ascontiguousarray(mp["desc"], np.uint8)
        buf = (C.c_uint8 * (0x400 * (m + 1)))()
        base = C.addressof(buf)
        for i in range(m):
            a = base + 0x400 * i
            self._fmat_at(a + 0xd8, xyz[i].reshape(3, 1))
            self._fmat_at(a + 0x168, nrm[i].reshape(3, 1))
            self._mat_at(a + 0x1c8, desc[i:i + 1])
            C.c_uint8.from_address(a + 0x238).value = 1 if mp["state"][i] == 2 else 0
            C.c_float.from_address(a + 0x248).value = rng_[i, 0]
            C.c_float.from_address(a + 0x24c).value = rng_[i, 1]
        keep.extend([xyz, nrm, rng_, desc, buf])
        return base

    def search_by_projection_sim3(self, kf, mp, scw, matched_in, th):
        """kf / mp as in make_keyframe / make_map_points; scw: 3x4 (rows of Scw); matched_in int32 [N]: index into the map points
        already matched to a key-frame feature (-1 = none).  Returns (matched int32 [N] after the call, nmatches)."""
        keep = []
        kb = self.make_keyframe(kf, keep)
        base = self.make_map_points(mp, keep)
        m, n = len(mp["desc"]), len(kf["desc"])
        vp = np.array([base + 0x400 * i for i in range(m)], np.uint64)
        vm = np.array([base + 0x400 * int(j) if j >= 0 else 0 for j in matched_in], np.uint64)
        S = np.ascontiguousarray(np.vstack([np.asarray(scw, np.float32).reshape(3, 4), [[0, 0, 0, 1]]]).astype(np.float32))
        smat = (C.c_uint64 * 12)()
        self._fmat_at(C.addressof(smat), S)
        v1, v2 = (C.c_uint64 * 3)(), (C.c_uint64 * 3)()
        v1[0], v1[1], v1[2] = vp.ctypes.data, vp.ctypes.data + vp.nbytes, vp.ctypes.data + vp.nbytes
        v2[0], v2[1], v2[2] = vm.ctypes.data, vm.ctypes.data + vm.nbytes, vm.ctypes.data + vm.nbytes
        fn = getattr(self.lib, "_ZN9ORB_SLAM210ORBmatcher18SearchByProjectionEPNS_8KeyFrameEN2cv3MatERKSt6vectorIPNS_8MapPointESaIS7_EERS9_i")
        fn.argtypes, fn.restype = [C.c_void_p] * 5 + [C.c_int], C.c_int
        matcher = (C.c_uint8 * 8)()
        C.c_float.from_address(C.addressof(matcher)).value = np.float32(0.75)
        matcher[4] = 1
        nm = fn(C.addressof(matcher), kb, C.addressof(smat), C.addressof(v1), C.addressof(v2), int(th))
        out = np.array([(int(p) - base) // 0x400 if p else -1 for p in vm], np.int32)
        return out, int(nm)

    # ---- ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, float th) (@0x7a500, LocalMapping::SearchInNeighbors) ----
    def fuse(self, kf, mp, kf_points, th, scw=None):
        """kf as make_keyframe (with uright, inv_level_sigma2, tcw, ow, mbf); mp as make_map_points plus state 0 = NULL entry, 3 = already
        in this key frame, and nobs [M]; kf_points int32 [N]: index (into an extra pool of M2 = N key-frame points appended after the
        map points) of the map point each key-frame feature holds, -1 = none, with kf_nobs [N] their observation counts and kf_bad [N].
        Returns (nFused, log) with log = list of (kind, a, b): 1 AddObservation(mp a, idx b), 2 AddMapPoint(mp a, idx b),
        3 Replace(point a by point b); point indices >= M denote the key frame's own points (M + feature index)."""
        keep = []
        kb = self.make_keyframe(kf, keep)
        m, n = len(mp["desc"]), len(kf["desc"])
        # map points 0..M-1, then one faked point per key-frame feature (M + i)
        pool = dict(state=np.concatenate([np.where(np.asarray(mp["state"]) == 2, 2, 1), np.where(np.asarray(kf_points["bad"]), 2, 1)]).astype(np.uint8),
                    xyz=np.vstack([mp["xyz"], np.zeros((n, 3), np.float32)]), normal=np.vstack([mp["normal"], np.zeros((n, 3), np.float32)]),
                    dist_range=np.vstack([mp["dist_range"], np.ones((n, 2), np.float32)]), desc=np.vstack([mp["desc"], np.zeros((n, 32), np.uint8)]))
        base = self.make_map_points(pool, keep)
        for i in range(m):
            C.c_int32.from_address(base + 0x400 * i + 0x18).value = int(mp["nobs"][i])
            C.c_uint8.from_address(base + 0x400 * i + 0x3f0).value = 1 if mp["state"][i] == 3 else 0
        for i in range(n):
            C.c_int32.from_address(base + 0x400 * (m + i) + 0x18).value = int(kf_points["nobs"][i])
        vp = np.array([base + 0x400 * i if mp["state"][i] else 0 for i in range(m)], np.uint64)
        kmp = np.array([base + 0x400 * (m + i) if kf_points["has"][i] else 0 for i in range(n)], np.uint64)
        o = (C.c_uint64 * 3).from_address(kb + 0x520)
        o[0], o[1], o[2] = kmp.ctypes.data, kmp.ctypes.data + kmp.nbytes, kmp.ctypes.data + kmp.nbytes
        v1 = (C.c_uint64 * 3)()
        v1[0], v1[1], v1[2] = vp.ctypes.data, vp.ctypes.data + vp.nbytes, vp.ctypes.data + vp.nbytes
        begin = self._shims.refshim_fuse_begin
        begin.argtypes, begin.restype = [C.c_void_p], None
        begin(kb)
        matcher = (C.c_uint8 * 8)()
        C.c_float.from_address(C.addressof(matcher)).value = np.float32(0.6)
        matcher[4] = 1
        replace = None
        if scw is None:
            fn = getattr(self.lib, "_ZN9ORB_SLAM210ORBmatcher4FuseEPNS_8KeyFrameERKSt6vectorIPNS_8MapPointESaIS5_EEf")
            fn.argtypes, fn.restype = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float], C.c_int
            nf = fn(C.addressof(matcher), kb, C.addressof(v1), np.float32(th))
        else:
            # Fuse(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, float th, vector<MapPoint*>& vpReplacePoint) (@0x7bb20):
            # entries with state 4 alias a map point the key frame already holds (pKF->GetMapPoints() makes them "already found")
            for i in range(m):
                if mp["state"][i] == 4:
                    vp[i] = base + 0x400 * (m + int(mp["alias"][i]))
            S = np.ascontiguousarray(np.vstack([np.asarray(scw, np.float32).reshape(3, 4), [[0, 0, 0, 1]]]).astype(np.float32))
            smat = (C.c_uint64 * 12)()
            self._fmat_at(C.addressof(smat), S)
            rp = np.zeros(m, np.uint64)
            v2 = (C.c_uint64 * 3)()
            v2[0], v2[1], v2[2] = rp.ctypes.data, rp.ctypes.data + rp.nbytes, rp.ctypes.data + rp.nbytes
            fn = getattr(self.lib, "_ZN9ORB_SLAM210ORBmatcher4FuseEPNS_8KeyFrameEN2cv3MatERKSt6vectorIPNS_8MapPointESaIS7_EEfRS9_")
            fn.argtypes, fn.restype = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p], C.c_int
            nf = fn(C.addressof(matcher), kb, C.addressof(smat), C.addressof(v1), np.float32(th), C.addressof(v2))
            replace = np.array([(int(p) - base) // 0x400 if p else -1 for p in rp], np.int32)
        getlog = self._shims.refshim_fuse_log
        getlog.argtypes, getlog.restype = [C.c_void_p, C.c_int], C.c_int
        buf = (C.c_long * (9 * (m + 1)))()
        cnt = getlog(buf, len(buf))
        log = []
        for k in range(0, cnt, 3):
            kind, a_, b_ = buf[k], buf[k + 1], buf[k + 2]
            ai = (a_ - base) // 0x400
            bi = (b_ - base) // 0x400 if kind == 3 else b_
            log.append((int(kind), int(ai), int(bi)))
        return (int(nf), log) if scw is None else (int(nf), log, replace)

    # ---- ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th) (@0x838b0, LoopClosing::ComputeSim3) ----
    def search_by_sim3(self, kf1, kf2, mp1, mp2, s12, R12, t12, th, matched_in):
        """kf1 / kf2 as make_keyframe (with tcw); mp1 / mp2: the key frames' own map points, one per feature (state 0 = none, 1 good,
        2 bad); matched_in int32 [N1]: feature of KF2 whose map point is already in vpMatches12[i] (-1 = none).  Returns
        (vpMatches12 as indices into KF2's features, nFound)."""
        keep = []
        for kf in (kf1, kf2):
            kf.setdefault("uright", np.full(len(kf["desc"]), -1, np.float32))
            kf.setdefault("inv_level_sigma2", np.ones(len(kf["scale_factors"]), np.float32))
            kf.setdefault("ow", np.zeros(3, np.float32)); kf.setdefault("mbf", 0.0)
        b1, b2 = self.make_keyframe(kf1, keep), self.make_keyframe(kf2, keep)
        n1, n2 = len(kf1["desc"]), len(kf2["desc"])
        base1, base2 = self.make_map_points(mp1, keep), self.make_map_points(mp2, keep)
        for i in range(n1):
            C.c_int32.from_address(base1 + 0x400 * i + 0x3f4).value = -1
        for i in range(n2):
            C.c_int32.from_address(base2 + 0x400 * i + 0x3f4).value = i      # KF2's own points are observed in KF2 at their index
        v1 = np.array([base1 + 0x400 * i if mp1["state"][i] else 0 for i in range(n1)], np.uint64)
        v2 = np.array([base2 + 0x400 * i if mp2["state"][i] else 0 for i in range(n2)], np.uint64)
        for b, v in ((b1, v1), (b2, v2)):
            o = (C.c_uint64 * 3).from_address(b + 0x520)
            o[0], o[1], o[2] = v.ctypes.data, v.ctypes.data + v.nbytes, v.ctypes.data + v.nbytes
        vm = np.array([base2 + 0x400 * int(j) if j >= 0 else 0 for j in matched_in], np.uint64)
        vv = (C.c_uint64 * 3)()
        vv[0], vv[1], vv[2] = vm.ctypes.data, vm.ctypes.data + vm.nbytes, vm.ctypes.data + vm.nbytes
        Rm = np.ascontiguousarray(R12, np.float32).reshape(3, 3); tm = np.ascontiguousarray(t12, np.float32).reshape(3, 1)
        rmat, tmat = (C.c_uint64 * 12)(), (C.c_uint64 * 12)()
        self._fmat_at(C.addressof(rmat), Rm); self._fmat_at(C.addressof(tmat), tm)
        sv, thv = C.c_float(np.float32(s12)), np.float32(th)
        fn = getattr(self.lib, "_ZN9ORB_SLAM210ORBmatcher12SearchBySim3EPNS_8KeyFrameES2_RSt6vectorIPNS_8MapPointESaIS5_EERKfRKN2cv3MatESE_f")
        fn.argtypes, fn.restype = [C.c_void_p] * 7 + [C.c_float], C.c_int
        matcher = (C.c_uint8 * 8)()
        C.c_float.from_address(C.addressof(matcher)).value = np.float32(0.75)
        matcher[4] = 1
        nf = fn(C.addressof(matcher), b1, b2, C.addressof(vv), C.addressof(sv), C.addressof(rmat), C.addressof(tmat), thv)
        out = np.array([(int(p) - base2) // 0x400 if p else -1 for p in vm], np.int32)
        return out, int(nf)

    # ---- ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo) (@0x86b30) ----
    # KeyFrame: fx/fy/cx/cy @0x130..0x13c, N @0x154, mvKeysUn @0x170, mvuRight @0x188, mDescriptors @0x1b8, mFeatVec @0x248,
    # mvScaleFactors @0x2e8, mvLevelSigma2 @0x300, Tcw (cv::Mat 4x4) @0x3a0, Ow (3x1) @0x460, mvpMapPoints @0x520 (offsets from
    # the function and from KeyFrame::GetCameraCenter / GetRotation / GetTranslation / GetMapPoint).  cv::Mat F12 is passed by
    # invisible reference (non-trivial copy constructor).
    def search_for_triangulation(self, kf1, kf2, F12, pose, cam, scale_factors, level_sigma2, only_stereo=False, check_ori=True):
        """Inputs as tests/matchdata.py: triangulation_case returns them; pose = (R2w, t2w, Cw).  The epipole is computed by the
        reference itself.  Returns (vMatches12-equivalent int32 [N1], nmatches, pairs)."""
        f32 = np.float32
        R2w, t2w, Cw = pose
        n1, n2 = len(kf1["desc"]), len(kf2["desc"])
        keep = []
        def make_kf(kf, n, Tcw, Ow):
            o = (C.c_uint64 * (0x800 // 8))()
            b = C.addressof(o)
            keys = np.zeros(n, self.KP)
            keys["x"], keys["y"], keys["angle"] = kf["xy"][:, 0], kf["xy"][:, 1], kf["angle"]
            if "octave" in kf:
                keys["octave"] = kf["octave"]
            ur = np.ascontiguousarray(kf["uright"], np.float32)
            d = np.ascontiguousarray(kf["desc"], np.uint8)
            dummy = (C.c_uint8 * 0x400)()
            mp = np.array([C.addressof(dummy) if h else 0 for h in kf["has_mp"]], np.uint64)
            sfa, sga = np.ascontiguousarray(scale_factors, np.float32), np.ascontiguousarray(level_sigma2, np.float32)
            arrs = [np.ascontiguousarray(kf[k], np.int32) for k in ("nodes", "start", "idx")]
            def setv(off, arr):
                o[off // 8], o[off // 8 + 1], o[off // 8 + 2] = arr.ctypes.data, arr.ctypes.data + arr.nbytes, arr.ctypes.data + arr.nbytes
            for off, v in zip((0x130, 0x134, 0x138, 0x13c), cam):
                C.c_float.from_address(b + off).value = f32(v)
            C.c_int32.from_address(b + 0x154).value = n
            setv(0x170, keys); setv(0x188, ur); setv(0x520, mp); setv(0x2e8, sfa); setv(0x300, sga)
            self._mat_at(b + 0x1b8, d)
            self._fmat_at(b + 0x3a0, Tcw)
            self._fmat_at(b + 0x460, Ow)
            build = self._shims.refshim_build_featvec
            build.argtypes, build.restype = [C.c_void_p] * 4 + [C.c_int], None
            build(b + 0x248, arrs[0].ctypes.data, arrs[1].ctypes.data, arrs[2].ctypes.data, len(arrs[0]))
            keep.extend([o, keys, ur, d, dummy, mp, sfa, sga, arrs, Tcw, Ow])
            return b
        T2 = np.ascontiguousarray(np.vstack([np.hstack([np.asarray(R2w, np.float32), np.asarray(t2w, np.float32).reshape(3, 1)]),
                                             [[0, 0, 0, 1]]]).astype(np.float32))
        eye = np.ascontiguousarray(np.eye(4, dtype=np.float32))
        b1 = make_kf(kf1, n1, eye, np.ascontiguousarray(np.asarray(Cw, np.float32).reshape(3, 1)))
        b2 = make_kf(kf2, n2, T2, np.zeros((3, 1), np.float32))
        Fm = np.ascontiguousarray(np.asarray(F12, np.float32).reshape(3, 3))
        fmat = (C.c_uint64 * 12)()
        self._fmat_at(C.addressof(fmat), Fm)
        fn = getattr(self.lib, "_ZN9ORB_SLAM210ORBmatcher22SearchForTriangulationEPNS_8KeyFrameES2_N2cv3MatERSt6vectorISt4pairImmESaIS7_EEb")
        fn.argtypes, fn.restype = [C.c_void_p] * 5 + [C.c_bool], C.c_int
        matcher = (C.c_uint8 * 8)()
        C.c_float.from_address(C.addressof(matcher)).value = f32(0.6)
        matcher[4] = 1 if check_ori else 0
        res = (C.c_uint64 * 3)()
        n = fn(C.addressof(matcher), b1, b2, C.addressof(fmat), C.addressof(res), bool(only_stereo))
        cnt = (res[1] - res[0]) // 16
        pr = np.ctypeslib.as_array(C.cast(res[0], C.POINTER(C.c_uint64)), (cnt * 2,)).reshape(cnt, 2).astype(np.int64) if cnt else np.empty((0, 2), np.int64)
        m = np.full(n1, -1, np.int32)
        for i, j in pr:
            m[i] = j
        return m, int(n), [(int(i), int(j)) for i, j in pr]

    # ---- Frame::ComputeStereoFromRGBD(const cv::Mat& imDepth) (@0xf6860): mvuRight @0x138, mvDepth @0x150 from mvKeys @0xf0,
    # mvKeysUn @0x120, mbf @0xe0 and the float depth map ----
    def compute_stereo_from_rgbd(self, xy, un_xy, depth, mbf):
        n = len(xy)
        k, ku = np.zeros(n, self.KP), np.zeros(n, self.KP)
        k["x"], k["y"] = xy[:, 0], xy[:, 1]
        ku["x"], ku["y"] = un_xy[:, 0], un_xy[:, 1]
        d = np.ascontiguousarray(depth, np.float32)
        fr = (C.c_uint64 * (0x400 // 8))()
        b = C.addressof(fr)
        C.c_int32.from_address(b + 0xec).value = n
        C.c_float.from_address(b + 0xe0).value = np.float32(mbf)
        for off, arr in ((0xf0, k), (0x120, ku)):
            fr[off // 8], fr[off // 8 + 1], fr[off // 8 + 2] = arr.ctypes.data, arr.ctypes.data + arr.nbytes, arr.ctypes.data + arr.nbytes
        dm = (C.c_uint64 * 12)()
        self._fmat_at(C.addressof(dm), d)
        fn = getattr(self.lib, "_ZN9ORB_SLAM25Frame21ComputeStereoFromRGBDERKN2cv3MatE")
        fn.argtypes, fn.restype = [C.c_void_p, C.c_void_p], None
        fn(b, C.addressof(dm))
        get = lambda off: np.ctypeslib.as_array(C.cast(fr[off // 8], C.POINTER(C.c_float)), (n,)).copy()
        return get(0x138), get(0x150)

    # ---- ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) (@0x7db00) ----
    def search_for_initialization(self, f1, f2, cam4, prev_matched, window_size=100, nnratio=0.9, check_ori=True):
        f32 = np.float32
        st = lambda name: C.c_float.in_dll(self.lib, name)
        st("_ZN9ORB_SLAM25Frame6mnMinXE").value, st("_ZN9ORB_SLAM25Frame6mnMinYE").value = f32(cam4[0]), f32(cam4[1])
        st("_ZN9ORB_SLAM25Frame21mfGridElementWidthInvE").value = f32(cam4[2])
        st("_ZN9ORB_SLAM25Frame22mfGridElementHeightInvE").value = f32(cam4[3])
        keep = []
        def make(f):
            n = len(f["desc"])
            k = np.zeros(n, self.KP)
            k["x"], k["y"], k["octave"], k["angle"] = f["xy"][:, 0], f["xy"][:, 1], f["octave"], f["angle"]
            d = np.ascontiguousarray(f["desc"], np.uint8)
            o = (C.c_uint64 * (0x12800 // 8))()
            b = C.addressof(o)
            C.c_int32.from_address(b + 0xec).value = n
            for off in (0xf0, 0x120):
                o[off // 8], o[off // 8 + 1], o[off // 8 + 2] = k.ctypes.data, k.ctypes.data + k.nbytes, k.ctypes.data + k.nbytes
            self._mat_at(b + 0x1c8, d)
            keep.extend([k, d, o])
            return b
        b1, b2 = make(f1), make(f2)
        assign = getattr(self.lib, "_ZN9ORB_SLAM25Frame20AssignFeaturesToGridEv")
        assign.argtypes, assign.restype = [C.c_void_p], None
        assign(b2)
        prev = np.array(prev_matched, np.float32).reshape(-1, 2).copy()
        pv = (C.c_uint64 * 3)(prev.ctypes.data, prev.ctypes.data + prev.nbytes, prev.ctypes.data + prev.nbytes)
        mv = (C.c_uint64 * 3)()
        fn = getattr(self.lib, "_ZN9ORB_SLAM210ORBmatcher23SearchForInitializationERNS_5FrameES2_RSt6vectorIN2cv6Point_IfEESaIS6_EERS3_IiSaIiEEi")
        fn.argtypes, fn.restype = [C.c_void_p] * 5 + [C.c_int], C.c_int
        matcher = (C.c_uint8 * 8)()
        C.c_float.from_address(C.addressof(matcher)).value = f32(nnratio)
        matcher[4] = 1 if check_ori else 0
        n = fn(C.addressof(matcher), b1, b2, C.addressof(pv), C.addressof(mv), int(window_size))
        cnt = (mv[1] - mv[0]) // 4
        m = np.ctypeslib.as_array(C.cast(mv[0], C.POINTER(C.c_int32)), (cnt,)).copy()
        return m, int(n), prev

    # ---- Frame::UndistortKeyPoints() (@0xf8630) and Frame::ComputeImageBounds(const cv::Mat&) (@0xf6010) ----
    # Frame: mK (cv::Mat 3x3 CV_32F) @0x20, mDistCoef (4x1 or 5x1 CV_32F) @0x80, N @0xec, mvKeys @0xf0, mvKeysUn @0x120.
    def undistort_and_bounds(self, xy, calib, cols, rows):
        """calib: dict fx fy cx cy k1 k2 p1 p2 k3.  Returns (mvKeysUn xy float32 [n, 2], (mnMinX, mnMaxX, mnMinY, mnMaxY))."""
        n = len(xy)
        k = np.zeros(n, self.KP)
        k["x"], k["y"], k["size"], k["octave"] = xy[:, 0], xy[:, 1], 31, 2
        K = np.array([[calib["fx"], 0, calib["cx"]], [0, calib["fy"], calib["cy"]], [0, 0, 1]], np.float32)
        D = np.array([[calib[c]] for c in ("k1", "k2", "p1", "p2", "k3")], np.float32)
        fr = (C.c_uint64 * (0x400 // 8))()
        b = C.addressof(fr)
        C.c_int32.from_address(b + 0xec).value = n
        fr[0xf0 // 8], fr[0xf0 // 8 + 1], fr[0xf0 // 8 + 2] = k.ctypes.data, k.ctypes.data + k.nbytes, k.ctypes.data + k.nbytes
        self._fmat_at(b + 0x20, K)
        self._fmat_at(b + 0x80, D)
        fn = getattr(self.lib, "_ZN9ORB_SLAM25Frame18UndistortKeyPointsEv")
        fn.argtypes, fn.restype = [C.c_void_p], None
        fn(b)
        cnt = (fr[0x128 // 8] - fr[0x120 // 8]) // 28
        un = np.ctypeslib.as_array(C.cast(fr[0x120 // 8], C.POINTER(C.c_uint8)), (cnt * 28,)).view(self.KP).copy()
        im = (C.c_uint64 * 12)()
        dummy = np.zeros((rows, cols), np.uint8)
        self._mat_at(C.addressof(im), dummy)
        fb = getattr(self.lib, "_ZN9ORB_SLAM25Frame18ComputeImageBoundsERKN2cv3MatE")
        fb.argtypes, fb.restype = [C.c_void_p, C.c_void_p], None
        fb(b, C.addressof(im))
        st = lambda name: float(C.c_float.in_dll(self.lib, name).value)
        bounds = (st("_ZN9ORB_SLAM25Frame6mnMinXE"), st("_ZN9ORB_SLAM25Frame6mnMaxXE"), st("_ZN9ORB_SLAM25Frame6mnMinYE"),
                  st("_ZN9ORB_SLAM25Frame6mnMaxYE"))
        return un, bounds
