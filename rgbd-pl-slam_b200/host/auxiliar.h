// auxiliar.h — helpers of the line front-end, mirroring the reference's include/auxiliar.h (comparators :30-51,
// sort_lines_by_response :67-72, compare_by_maxDepth :54-64, vector_mad :92-106).  Host-side and tiny; written
// against cv_compat.h so it builds with or without OpenCV.
#pragma once
#include <algorithm>
#include <cmath>
#include <utility>
#include <vector>

#include "cv_compat.h"

// ascending distance of the nearest neighbour
struct compare_descriptor_by_NN_dist {
  inline bool operator()(const std::vector<cv::DMatch>& a, const std::vector<cv::DMatch>& b) { return a[0].distance < b[0].distance; }
};
// ascending gap between the second-nearest and the nearest neighbour
struct compare_descriptor_by_NN12_dist {
  inline bool operator()(const std::vector<cv::DMatch>& a, const std::vector<cv::DMatch>& b) {
    return (a[1].distance - a[0].distance) < (b[1].distance - b[0].distance);
  }
};
struct sort_descriptor_by_queryIdx {
  inline bool operator()(const std::vector<cv::DMatch>& a, const std::vector<cv::DMatch>& b) { return a[0].queryIdx < b[0].queryIdx; }
};
struct compare_by_maxDepth {
  inline bool operator()(const std::pair<std::pair<float, float>, int>& a, const std::pair<std::pair<float, float>, int>& b) {
    return std::max(a.first.first, a.first.second) < std::max(b.first.first, b.first.second);
  }
};
struct sort_lines_by_response {
  inline bool operator()(const cv::line_descriptor::KeyLine& a, const cv::line_descriptor::KeyLine& b) { return a.response > b.response; }
};
// 1.4826 * median(|x - median(x)|), medians taken at index n/2 of the sorted sample
inline double vector_mad(std::vector<double> residues) {
  if (residues.empty()) return 0.0;
  const int n = (int)residues.size();
  std::sort(residues.begin(), residues.end());
  const double median = residues[n / 2];
  for (int i = 0; i < n; i++) residues[i] = std::fabs(residues[i] - median);
  std::sort(residues.begin(), residues.end());
  return 1.4826 * residues[n / 2];
}
