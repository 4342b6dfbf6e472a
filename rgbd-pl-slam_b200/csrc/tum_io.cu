// tum_io.cu — host-only: the TUM RGB-D association list and trajectory file, the two on-disk formats either side of the
// per-frame path.  No device code; lives in the C-ABI library so that one .so serves a port.
//
//   plslam_tum_load_associations  <- LoadImages(strAssociationFilename, vstrImageFilenamesRGB, vstrImageFilenamesD,
//                                    vTimestamps), reference Examples/RGB-D/rgbd_tum.cc:151-176
//   plslam_tum_pose_to_line / plslam_tum_save_trajectory
//                                 <- the output statement of ORB_SLAM2::System::SaveTrajectoryTUM (include/System.h:104;
//                                    lib/libORB_SLAM2.so@0x3df90: `fixed` @0x3e148, Rwc = Tcw.rowRange(0,3).colRange(0,3).t()
//                                    @0x3e732-0x3e794, twc = -Rwc * tcw @0x3e940-0x3e956, Converter::toQuaternion @0x3ea79,
//                                    precision 6 for the time stamp @0x3eb84, 9 for the seven pose values @0x3ebc0, every
//                                    value inserted as double, single blanks, endl @0x3ecc7)
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>

#include "common.cuh"

namespace plslam {
namespace {

struct Association {
  double t;
  std::string rgb, depth;
};

// One entry per non-empty line, whatever the line holds: the reference extracts "double, word, double, word" from a
// stringstream and pushes the results even when an extraction failed (0 / empty string; after the first failure nothing more
// is read from that line).  Same rules here on the raw bytes: a number is the longest prefix strtod accepts, a word runs to
// the next white space.
void parse_line(const std::string& line, Association& a) {
  const char* p = line.c_str();
  bool ok = true;
  auto skip = [&] {
    while (*p == ' ' || (*p >= '\t' && *p <= '\r')) ++p;
  };
  auto number = [&](double& v) {
    v = 0.0;
    if (!ok) return;
    skip();
    char* end = nullptr;
    const double x = std::strtod(p, &end);
    if (end == p) {
      ok = false;
      return;
    }
    v = x;
    p = end;
  };
  auto word = [&](std::string& w) {
    w.clear();
    if (!ok) return;
    skip();
    if (!*p) {
      ok = false;
      return;
    }
    const char* b = p;
    while (*p && !(*p == ' ' || (*p >= '\t' && *p <= '\r'))) ++p;
    w.assign(b, p);
  };
  double depth_time;
  number(a.t);
  word(a.rgb);
  number(depth_time);  // read and dropped: the RGB time stamp is the frame's
  word(a.depth);
}

int read_associations(const char* path, std::vector<Association>& out) {
  std::FILE* f = std::fopen(path, "rb");
  if (!f) {
    set_error("cannot open association file %s", path);
    return PLSLAM_ERR_INVALID;
  }
  std::string buf;
  char chunk[1 << 16];
  size_t got;
  while ((got = std::fread(chunk, 1, sizeof(chunk), f)) > 0) buf.append(chunk, got);
  std::fclose(f);
  for (size_t pos = 0; pos < buf.size();) {
    size_t e = buf.find('\n', pos);
    if (e == std::string::npos) e = buf.size();
    if (e > pos) {
      Association a;
      parse_line(buf.substr(pos, e - pos), a);
      out.push_back(a);
    }
    pos = e + 1;
  }
  return PLSLAM_OK;
}

// Eigen::Quaterniond(Matrix3d) as Converter::toQuaternion uses it: double arithmetic on the float rotation, result narrowed
// to float in the order x y z w
void rotation_to_quaternion(const double m[3][3], float q_xyzw[4]) {
  double q[4];  // x y z w
  double t = m[0][0] + m[1][1] + m[2][2];
  if (t > 0.0) {
    t = std::sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m[2][1] - m[1][2]) * t;
    q[1] = (m[0][2] - m[2][0]) * t;
    q[2] = (m[1][0] - m[0][1]) * t;
  } else {
    int i = 0;
    if (m[1][1] > m[0][0]) i = 1;
    if (m[2][2] > m[i][i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (m[k][j] - m[j][k]) * t;
    q[j] = (m[j][i] + m[i][j]) * t;
    q[k] = (m[k][i] + m[i][k]) * t;
  }
  for (int c = 0; c < 4; ++c) q_xyzw[c] = (float)q[c];
}

int pose_to_line(double timestamp, const float* T, char* line, int capacity, int* len) {
  // Rwc = Rcw^T; twc = -Rwc * tcw.  cv::gemm takes its small-matrix path for a 3x3 by 3x1 CV_32F product: float products
  // summed left to right in float, then (double)sum * alpha + 0 (pinned against cv2 4.13: tests/test_tum_io_cpu.py)
  double R[3][3];
  float twc[3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) R[r][c] = (double)T[c * 4 + r];
  for (int r = 0; r < 3; ++r) {
    float acc = T[0 * 4 + r] * T[0 * 4 + 3];
    acc = acc + T[1 * 4 + r] * T[1 * 4 + 3];
    acc = acc + T[2 * 4 + r] * T[2 * 4 + 3];
    twc[r] = (float)((double)acc * -1.0 + 0.0);
  }
  float q[4];
  rotation_to_quaternion(R, q);
  const int n = std::snprintf(line, (size_t)capacity, "%.6f %.9f %.9f %.9f %.9f %.9f %.9f %.9f\n", timestamp, (double)twc[0],
                              (double)twc[1], (double)twc[2], (double)q[0], (double)q[1], (double)q[2], (double)q[3]);
  if (n < 0 || n >= capacity) {
    set_error("trajectory line does not fit %d bytes", capacity);
    return PLSLAM_ERR_CAPACITY;
  }
  *len = n;
  return PLSLAM_OK;
}

}  // namespace
}  // namespace plslam

using namespace plslam;

extern "C" {

int plslam_tum_load_associations(const char* path, double* timestamps, char* rgb_names, char* depth_names, int name_stride,
                                 int capacity, int* count) {
  PL_CHECK_ARG(path && count && capacity >= 0 && (capacity == 0 || (timestamps && rgb_names && depth_names && name_stride > 1)));
  std::vector<Association> v;
  const int rc = read_associations(path, v);
  if (rc) return rc;
  *count = (int)v.size();
  if (capacity == 0) return PLSLAM_OK;  // size query
  if ((int)v.size() > capacity) {
    set_error("%d associations, capacity %d", (int)v.size(), capacity);
    return PLSLAM_ERR_CAPACITY;
  }
  for (size_t i = 0; i < v.size(); ++i) {
    if ((int)v[i].rgb.size() >= name_stride || (int)v[i].depth.size() >= name_stride) {
      set_error("file name longer than name_stride - 1 = %d", name_stride - 1);
      return PLSLAM_ERR_CAPACITY;
    }
    timestamps[i] = v[i].t;
    std::memset(rgb_names + i * (size_t)name_stride, 0, (size_t)name_stride);
    std::memset(depth_names + i * (size_t)name_stride, 0, (size_t)name_stride);
    std::memcpy(rgb_names + i * (size_t)name_stride, v[i].rgb.data(), v[i].rgb.size());
    std::memcpy(depth_names + i * (size_t)name_stride, v[i].depth.data(), v[i].depth.size());
  }
  return PLSLAM_OK;
}

int plslam_tum_pose_to_line(double timestamp, const float* Tcw_3x4, char* line, int capacity, int* length) {
  PL_CHECK_ARG(Tcw_3x4 && line && length && capacity >= 2);
  return pose_to_line(timestamp, Tcw_3x4, line, capacity, length);
}

int plslam_tum_save_trajectory(const char* path, const double* timestamps, const float* Tcw_3x4, int n) {
  PL_CHECK_ARG(path && n >= 0 && (n == 0 || (timestamps && Tcw_3x4)));
  std::FILE* f = std::fopen(path, "w");
  if (!f) {
    set_error("cannot open %s for writing", path);
    return PLSLAM_ERR_INVALID;
  }
  char line[256];
  for (int i = 0; i < n; ++i) {
    int len = 0;
    const int rc = pose_to_line(timestamps[i], Tcw_3x4 + (size_t)i * 12, line, (int)sizeof(line), &len);
    if (rc) {
      std::fclose(f);
      return rc;
    }
    std::fwrite(line, 1, (size_t)len, f);
  }
  std::fclose(f);
  return PLSLAM_OK;
}

}  // extern "C"
