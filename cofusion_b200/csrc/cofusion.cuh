// cofusion.cuh -- cfb::CoFusion: the per-frame sequencing of CoFusion::processFrame
// (Core/CoFusion.cpp:171-545) for the models living on one device, GL-free.
//
// Kept from the reference: call order (track all -> predict -> indices/fuse/indices/clean -> predict),
// tick semantics, first-frame initialisation, fill-in for the background model only, external label
// masks (FrameData::mask).  Not here (out of scope): loop closure, ferns,
// re-detection, GUI, logging.  With enableMultipleModels the motion segmentation runs after tracking
// and drives spawn / deactivate exactly as CoFusion.cpp:227-299; without it the caller supplies the
// label mask (FrameData::mask) and spawns models itself.
#pragma once
#include <memory>
#include <vector>

#include "pipeline.cuh"
#include "segment_kernels.cuh"

namespace cfb {

struct CoFusionParams {   // constructor arguments / setters of CoFusion (CoFusion.h:47-66, :130-246)
  int timeDelta;          // 200 (GUI/MainController.cpp:186; effectively unused while openLoop)
  float depthCutoff;      // bilateral maxD, GUI default 5 (CoFusion::depthCutoff)
  float maxDepthProcessed;// 20 (CoFusion.cpp:51)
  float icpWeight;        // 10
  int pyramid, fastOdom, so3, frameToFrameRGB, rgbOnly;
  float confGlobalInit;   // 10  (MainController.cpp:175)
  float confObjectInit;   // 0.01
  float outlierCoefficient;  // 3 (GUI/Tools/GUI.h:208)
  unsigned maxSurfels;    // per model
  int predictBeforeFuse;  // 1: also run the predict() of CoFusion.cpp:347, whose images nobody reads
                          //    (loop closure is out of scope; identical results either way)
  int enableMultipleModels;   // 1: run performSegmentationCRF each frame (CoFusion.cpp:227-299)
  unsigned modelSpawnOffset;  // 20 (CoFusion.h:50): frames between two spawns
  SegParams seg;              // CRF parameters (GUI defaults)
};

class CoFusion {
 public:
  CoFusion(int device, int W, int H, float fx, float fy, float cx, float cy, const CoFusionParams& p);
  bool ok() const { return ctx.ok() && !models.empty() && models[0]->ok(); }
  // CoFusion::processFrame.  Buffers are host pointers (pinned or pageable) unless device_ptrs.
  // mask == nullptr: static scene, everything labelled background (CoFusion.cpp:190-197).
  cudaError_t processFrame(const uint8_t* rgb, const float* depth, const uint8_t* mask, bool device_ptrs,
                           float weightMultiplier);
  // spawn an object model (CoFusion::spawnObjectModel, CoFusion.cpp:588-597): created empty; it is
  // initialised by fusing the pixels labelled `id` of the current frame (CoFusion.cpp:265-276)
  cudaError_t spawnObjectModel(unsigned id, const float* initialPose /* 16 or null = background pose */);
  Model* model(size_t i) { return i < models.size() ? models[i].get() : nullptr; }
  size_t numModels() const { return models.size(); }
  int tick() const { return tick_; }
  cudaError_t predict();  // CoFusion::predict (CoFusion.cpp:533-545)

  bool batchedTracking = true;  // track all models of a frame in one persistent launch (gn_batched.cu)
  Context ctx;
  CoFusionParams params;
  std::vector<std::unique_ptr<Model>> models;
  std::vector<std::unique_ptr<Model>> inactiveModels;  // CoFusion::inactivateModel keeps the data
  // tracking statistics of model i after the last processFrame (waits for that frame's tracker)
  cudaError_t stats(size_t i, TrackStats* out) {
    if (i >= models.size()) return cudaErrorInvalidValue;
    cudaError_t e = models[i]->syncPose();
    if (e == cudaSuccess) *out = models[i]->odom.stats();
    return e;
  }

  // result of the last performSegmentation (CoFusion.cpp:232)
  std::unique_ptr<Segmentation> segmentation;
  std::vector<SegModelData> lastModelData;
  bool lastHasNewLabel = false;
  int lastSpawnedId = -1;       // id of the model spawned by the last frame, -1 = none
  int lastDeactivated = 0;      // number of models deactivated by the last frame

 private:
  cudaError_t segmentAndManageModels();  // CoFusion.cpp:227-299
  unsigned char takeNextModelID();       // getNextModelID(true)
  int tick_ = 1;
  unsigned spawnOffset_ = 0;
  unsigned char nextID_ = 1;  // id 0 went to the global model (CoFusion.cpp:70)
};

}  // namespace cfb
