// tum_oracle.cc — CPU restatement of the two on-disk formats around the path, written the way the reference writes them
// (iostream extraction / insertion), against which the product's byte-level parser and printf-style writer are compared.
//   LoadImages: reference Examples/RGB-D/rgbd_tum.cc:151-176.
//   Trajectory line: the output statement of ORB_SLAM2::System::SaveTrajectoryTUM (include/System.h:104; machine code
//   lib/libORB_SLAM2.so@0x3df90 — `fixed` @0x3e148, Rwc/twc @0x3e732-0x3e956, Converter::toQuaternion @0x3ea79,
//   setprecision(6) @0x3eb84, setprecision(9) @0x3ebc0, endl @0x3ecc7).  cv::gemm's small-matrix float path (float sum,
//   scaled in double) and Eigen's Quaterniond(Matrix3d) are third-party arithmetic, restated here and pinned against cv2 4.13 gemm and
//   scipy's Rotation in tests/test_tum_io_cpu.py.
//
// TEST INFRASTRUCTURE ONLY: never linked into the product.
#include <cmath>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <string>
#include <vector>

extern "C" {

// returns the number of entries; fills up to `capacity` of them (names NUL-terminated, name_stride bytes apart)
int oracle_tum_load(const char* path, double* timestamps, char* rgb, char* depth, int name_stride, int capacity) {
  std::ifstream in;
  in.open(path);
  if (!in.is_open()) return -1;
  int n = 0;
  while (!in.eof()) {
    std::string s;
    std::getline(in, s);
    if (!s.empty()) {
      std::stringstream ss;
      ss << s;
      double t = 0.0;  // the reference leaves it uninitialised: a white-space-only line prints an indeterminate stamp there
      std::string sRGB, sD;
      ss >> t;
      const double stamp = t;
      ss >> sRGB;
      ss >> t;
      ss >> sD;
      if (n < capacity) {
        timestamps[n] = stamp;
        std::memset(rgb + (size_t)n * name_stride, 0, name_stride);
        std::memset(depth + (size_t)n * name_stride, 0, name_stride);
        std::strncpy(rgb + (size_t)n * name_stride, sRGB.c_str(), name_stride - 1);
        std::strncpy(depth + (size_t)n * name_stride, sD.c_str(), name_stride - 1);
      }
      ++n;
    }
  }
  return n;
}

// Tcw: rows 0..2 of the pose, row-major 3x4 floats.  out7 = twc (3) + quaternion x y z w (4) as the floats that are printed.
void oracle_tum_pose(const float* T, float* out7) {
  double R[3][3];  // Rwc = Rcw^T (Eigen matrix of doubles built from the float entries)
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) R[r][c] = (double)T[c * 4 + r];
  for (int r = 0; r < 3; ++r) {  // -Rwc * tcw = gemm(Rwc, tcw, alpha = -1, beta = 0), small-matrix path: float sum, double scale
    const float p0 = T[0 * 4 + r] * T[0 * 4 + 3], p1 = T[1 * 4 + r] * T[1 * 4 + 3], p2 = T[2 * 4 + r] * T[2 * 4 + 3];
    const float s = (p0 + p1) + p2;
    out7[r] = (float)((double)s * -1.0 + 0.0);
  }
  double q[4];
  double t = R[0][0] + R[1][1] + R[2][2];
  if (t > 0.0) {
    t = std::sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R[2][1] - R[1][2]) * t;
    q[1] = (R[0][2] - R[2][0]) * t;
    q[2] = (R[1][0] - R[0][1]) * t;
  } else {
    int i = 0;
    if (R[1][1] > R[0][0]) i = 1;
    if (R[2][2] > R[i][i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(R[i][i] - R[j][j] - R[k][k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (R[k][j] - R[j][k]) * t;
    q[j] = (R[j][i] + R[i][j]) * t;
    q[k] = (R[k][i] + R[i][k]) * t;
  }
  for (int c = 0; c < 4; ++c) out7[3 + c] = (float)q[c];
}

int oracle_tum_pose_line(double timestamp, const float* T, char* line, int capacity) {
  float v[7];
  oracle_tum_pose(T, v);
  std::ostringstream f;
  f << std::fixed;
  f << std::setprecision(6) << timestamp << " " << std::setprecision(9) << v[0] << " " << v[1] << " " << v[2] << " " << v[3] << " "
    << v[4] << " " << v[5] << " " << v[6] << std::endl;
  const std::string s = f.str();
  if ((int)s.size() >= capacity) return -1;
  std::memcpy(line, s.c_str(), s.size() + 1);
  return (int)s.size();
}

}  // extern "C"
