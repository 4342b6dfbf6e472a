// orb.cuh — device-side parameter block and host context of the ORB extractor.
#pragma once
#include <vector>

#include "common.cuh"

namespace plslam {

constexpr int ORB_MAXL = 12;
constexpr int ORB_EDGE = 19;       // EDGE_THRESHOLD (@0x705e1)
constexpr int ORB_MINB = 16;       // EDGE_THRESHOLD-3 (@0x760c6)
constexpr int ORB_HALF_PATCH = 15; // HALF_PATCH_SIZE

struct OrbLevel {
  int w, h, pitch;                 // level image (level 0 uses the caller's pitch)
  unsigned long long off;          // byte offset of the level inside one frame's pyramid slab (levels >= 1)
  int nCols, nRows, wCell, hCell;  // FAST cell grid
  int cellBase;                    // first flattened cell index of this level
  int maxBorderX, maxBorderY;
  int quota, nIni;                 // DistributeOctTree N and number of root nodes
  float hX;
  int candCap;                     // candidate capacity (records) of this level
  unsigned long long candOff;      // offset (records) of this level in one frame's candidate slab
  int kpOff;                       // offset of this level in one frame's per-level keypoint slab
  int tileBase, tilesX;            // blur tiles
  float scale;                     // mvScaleFactor[level]
  float patchSize;                 // (float)(int)(31 * scale)
  unsigned long long coefOff;      // offset (ints) of the resize tables of the transition level-1 -> level
};

struct OrbParams {
  OrbLevel lv[ORB_MAXL];
  int nlevels, iniTh, minTh;
  int totalCells, totalTiles;
  int maxKp;       // sum(quota + 3)
  int nodeCap;     // max quota + 8
  int patchPitch, patchRows;  // FAST per-warp shared tile geometry
  int umax[16];
  int blurk[7];
  unsigned long long pyrFrameStride;   // bytes
  unsigned long long candFrameStride;  // records
};

struct OrbImages {
  const uint8_t* img0;  // level 0 = the caller's frames
  int pitch0;
  unsigned long long stride0;
  uint8_t* pyr;      // levels >= 1
  uint8_t* blurred;  // all levels (level 0 included), dense slabs like pyr but with lv[0] at blurOff0
};

class OrbExtractor {
 public:
  OrbExtractor(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh);
  ~OrbExtractor();

  int set_blur_kernel(const int32_t k[7]);
  int extract_device(const uint8_t* d_images, int batch, int W, int H, int pitch, size_t frame_stride,
                     plslam_keypoint_t* d_kps, uint8_t* d_desc, int capacity, int32_t* d_counts, cudaStream_t st);
  int extract_host(const uint8_t* images, int batch, int W, int H, int pitch, size_t frame_stride,
                   plslam_keypoint_t* kps, uint8_t* desc, int capacity, int32_t* counts);
  int copy_level(int frame, int level, int which, uint8_t* out, size_t out_bytes);
  int copy_candidates(int frame, int level, int32_t* xyr, int capacity, int* n_out);
  int level_size(int level, int* w, int* h) const;
  int max_keypoints() const;
  int check_status(cudaStream_t st);  // synchronises st and maps the device status word to a return code
  StageTimer* timer = nullptr;        // optional per-stage event timing (owned by the caller)
  const int* device_status() const { return status.as<int>(); }

  // constructor tables (ORBextractor.h:102-110)
  int nfeatures, nlevels, iniThFAST, minThFAST;
  double scaleFactor;
  std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
  std::vector<int> mnFeaturesPerLevel, umax;
  int blurk[7] = {18, 34, 48, 56, 48, 34, 18};

 private:
  int configure(int W, int H, int batch);
  int device = -1;
  int cfgW = 0, cfgH = 0, cfgB = 0;
  OrbParams P{};
  // last-batch bookkeeping for the debug accessors
  const uint8_t* last_img0 = nullptr;
  int last_pitch0 = 0, last_batch = 0;
  size_t last_stride0 = 0;
  size_t blurOff0 = 0, blurFrameStride = 0;
  DevBuf pyr, blurred, coef, cand, candCount, knode, lvlKp, lvlCnt, status, tileTab;
  DevBuf stageIn, stageKps, stageDesc, stageCnt;
  cudaStream_t ownStream = nullptr;
  void* pinnedStatus = nullptr;
  uintptr_t mapsKey[4][6] = {};  // what the cached TMA descriptor sets of k_blur and k_resize_tma were encoded for
  alignas(64) unsigned char mapsCache[4][4096] = {};  // the encoded sets (host memory; passed to the kernels by value)
  int mapsNext = 0;
};

}  // namespace plslam
