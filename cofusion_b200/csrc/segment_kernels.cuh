// segment_kernels.cuh -- motion segmentation (Core/Segmentation/Segmentation.cpp:124-706) on the device.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace cfb {

struct SegLimits {
  static constexpr int kMaxModels = 15;  // labels = models + 1 "new" label
};

struct SegParams {  // Segmentation.h:135-149, values = GUI defaults (seg_default_params)
  int crfIterations;
  float scaleFeaturesRGB, scaleFeaturesDepth, scaleFeaturesPos;
  float weightAppearance, weightSmoothness;
  float unaryThresholdNew, unaryKError, unaryWeightError;
  float maxRelSizeNew, minRelSizeNew;
};
void seg_default_params(SegParams* p);

struct SegModelData {  // SegmentationResult::ModelData (Segmentation.h:41-67)
  unsigned id;
  unsigned superPixelCount;
  float avgConfidence, depthMean, depthStd;
  unsigned short top, right, bottom, left;
};
struct SegResultHeader {
  int numModelData, hasNewLabel;
};

// Owns the scratch of one segmentation instance (one per frame geometry).
class Segmentation {
 public:
  Segmentation(int W, int H);
  ~Segmentation();
  Segmentation(const Segmentation&) = delete;
  Segmentation& operator=(const Segmentation&) = delete;
  bool ok() const { return ok_; }

  // gSLICr restatement (Slic.cpp:62-80): labels + per super-pixel pixel counts
  cudaError_t slic(const uint8_t* rgb, cudaStream_t s);

  // Segmentation::performSegmentationCRF.  rgb (HxWx3 u8), depth (HxW f32 metres, the raw frame),
  // icpError[m] (HxW f32), vertConf4[m] (HxW float4, .w read) are DEVICE pointers in HOST arrays.
  // fullSeg (HxW u8, device) receives model ids / 255.  md_host receives up to numModels+1 entries.
  // Synchronises the stream (the caller branches on hasNew, as the reference does).
  cudaError_t performSegmentationCRF(const uint8_t* rgb, const float* depth, int numModels,
                                     const unsigned char* modelIds, const float* const* icpError,
                                     const float* const* vertConf4, unsigned char nextModelID, bool allowNew,
                                     const SegParams& prm, uint8_t* fullSeg, SegModelData* md_host, int* md_count,
                                     bool* hasNew, cudaStream_t s);

  int W, H, mx, my, N;
  int launches = 0;      // kernels launched by the last performSegmentationCRF
  bool useGraph = true;  // replay the launch sequence as a CUDA graph on capturable streams
  // The super-pixels depend on the frame alone: a caller may run slic() ahead (on another stream, ordered before
  // performSegmentationCRF by an event) and set this to the image it ran on; the next performSegmentationCRF on
  // that image then skips its own SLIC pass and clears the field.
  const uint8_t* slicAheadOf = nullptr;
  // device scratch (public: the tests read labels / unary / lowMap through the C ABI)
  int* labels = nullptr;
  float* centers = nullptr;  // [2][N][5] ping-pong
  int* slicSums = nullptr;   // [6 iterations][N][6]: x, y, c0, c1, c2, count
  unsigned *counts = nullptr, *dcounts = nullptr;
  float *sums = nullptr;     // raw per super-pixel sums [maps][N]
  float *low = nullptr;      // low-res maps [maps][N]: depth, then icp / conf per model
  float *unary = nullptr, *f6 = nullptr;  // unaries [N][L]; node records [N][8] (6 features + x, y)
  float *T2 = nullptr;                    // smoothness kernel by grid offset [my][mx]
  float *K2 = nullptr, *K6 = nullptr;      // Gaussian kernels [N][N], rebuilt every frame
  float *n2 = nullptr, *n6 = nullptr, *Q = nullptr;
  float *nq2 = nullptr, *nq6 = nullptr;  // [2][N][Lmax] ping-pong
  uint8_t* lowMap = nullptr;
  SegModelData* md = nullptr;
  SegResultHeader* hdr = nullptr;
  float* depthRange = nullptr;
  void* h_out = nullptr;

 private:
  cudaError_t enqueue(const uint8_t* rgb, const float* depth, int numModels, const unsigned char* modelIds,
                      const float* const* icpError, const float* const* vertConf4, unsigned char nextModelID,
                      bool allowNew, const SegParams& prm, uint8_t* fullSeg, cudaStream_t s);
  bool ok_ = false;
  static constexpr int kGraphSlots = 4;
  void* graphExec_[kGraphSlots] = {nullptr, nullptr, nullptr, nullptr};  // cudaGraphExec_t of cached launch sequences
  void* graphKey_[kGraphSlots] = {nullptr, nullptr, nullptr, nullptr};   // the arguments each was captured with
  int graphNext_ = 0;
};

}  // namespace cfb
