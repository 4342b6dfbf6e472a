// dropin_demo.cc — exercises the drop-in classes exactly as Frame::ExtractORB / Frame::ExtractLSD (reference
// include/Frame.h:67,70) and Tracking would: reads raw 8-bit frames, runs ORBextractor::operator(), LineSegment::
// ExtractLineSegment and LineSegmentMathch, and dumps the results for tests/test_dropin_gpu.py to compare with the
// C-ABI outputs.  usage: dropin_demo W H frameA.raw frameB.raw out.bin
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ExtractLineSegment.h"
#include "LSDmatcher.h"
#include "ORBextractor.h"
#include "ORBmatcher.h"

using namespace ORB_SLAM2;

static cv::Mat load(const char* path, int W, int H) {
  cv::Mat m(H, W, CV_8U);
  FILE* f = fopen(path, "rb");
  if (!f || fread(m.data, 1, (size_t)W * H, f) != (size_t)W * H) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
  fclose(f);
  return m;
}

int main(int argc, char** argv) {
  if (argc != 6) { fprintf(stderr, "usage: %s W H a.raw b.raw out.bin\n", argv[0]); return 2; }
  const int W = atoi(argv[1]), H = atoi(argv[2]);
  cv::Mat imA = load(argv[3], W, H), imB = load(argv[4], W, H);
  try {
    ORBextractor* mpORBextractorLeft = new ORBextractor(1000, 1.2f, 8, 20, 7);  // TUM1.yaml:42-55
    LineSegment* mpLineSegment = new LineSegment();
    std::vector<cv::KeyPoint> mvKeysA, mvKeysB;
    cv::Mat mDescriptorsA, mDescriptorsB, mLdescA, mLdescB;
    std::vector<KeyLine> mvKeylinesA, mvKeylinesB;
    std::vector<Vector3d> mvKeyLineFunctionsA, mvKeyLineFunctionsB;
    (*mpORBextractorLeft)(imA, cv::Mat(), mvKeysA, mDescriptorsA);  // Frame::ExtractORB
    (*mpORBextractorLeft)(imB, cv::Mat(), mvKeysB, mDescriptorsB);
    mpLineSegment->ExtractLineSegment(imA, mvKeylinesA, mLdescA, mvKeyLineFunctionsA);  // Frame::ExtractLSD
    mpLineSegment->ExtractLineSegment(imB, mvKeylinesB, mLdescB, mvKeyLineFunctionsB);
    mpLineSegment->LineSegmentMathch(mLdescA, mLdescB);
    mpLineSegment->LineDescriptorMAD();
    LSDmatcher lm(0.8f, true);
    std::vector<int> lmatch;
    const int nl = lm.MatchKNN(mLdescA, mLdescB, lmatch);
    const int d01 = ORBmatcher::DescriptorDistance(mDescriptorsA.row(0), mDescriptorsB.row(0));
    FILE* f = fopen(argv[5], "wb");
    int hdr[8] = {(int)mvKeysA.size(), (int)mvKeysB.size(), (int)mvKeylinesA.size(), (int)mvKeylinesB.size(), nl, d01,
                  mpORBextractorLeft->GetLevels(), (int)mpLineSegment->Matches().size()};
    fwrite(hdr, sizeof(hdr), 1, f);
    fwrite(mvKeysA.data(), sizeof(cv::KeyPoint), mvKeysA.size(), f);
    for (int i = 0; i < mDescriptorsA.rows; ++i) fwrite(mDescriptorsA.ptr(i), 1, 32, f);
    fwrite(mvKeylinesA.data(), sizeof(KeyLine), mvKeylinesA.size(), f);
    for (int i = 0; i < mLdescA.rows; ++i) fwrite(mLdescA.ptr(i), 1, 32, f);
    for (auto& v : mvKeyLineFunctionsA) fwrite(v.v, sizeof(double), 3, f);
    fwrite(lmatch.data(), sizeof(int), lmatch.size(), f);
    double mads[2] = {mpLineSegment->NNMad(), mpLineSegment->NN12Mad()};
    fwrite(mads, sizeof(mads), 1, f);
    fclose(f);
    printf("dropin_demo ok: %zu/%zu keypoints, %zu/%zu lines, %d line matches, MAD %.3f %.3f\n", mvKeysA.size(), mvKeysB.size(),
           mvKeylinesA.size(), mvKeylinesB.size(), nl, mads[0], mads[1]);
    delete mpLineSegment;
    delete mpORBextractorLeft;
  } catch (const std::exception& e) {
    fprintf(stderr, "dropin_demo failed: %s\n", e.what());
    return 1;
  }
  return 0;
}
