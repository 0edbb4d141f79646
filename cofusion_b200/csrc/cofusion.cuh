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
#include "shard.cuh"

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

struct FrameInput {  // FrameData (Core/FrameData.h:25-50) + what the log readers do to it before processFrame
  const uint8_t* rgb = nullptr;       // H x W x 3, 8 bit
  const float* depth = nullptr;       // metric f32, or
  const uint16_t* depth16 = nullptr;  // raw sensor units (KlgLogReader.cpp:53-84), converted on the device
  float depthScale = 0.001f;          // metres per raw unit
  bool flipColors = false;            // FrameData::flipColors: the image arrives as BGR
  const uint8_t* mask = nullptr;      // external label image or null
  bool device_ptrs = false;
  int64_t timestamp = 0;              // logged with the poses (CoFusion.cpp:516)
};

class CoFusion {
 public:
  CoFusion(int device, int W, int H, float fx, float fy, float cx, float cy, const CoFusionParams& p);
  bool ok() const { return ctx.ok() && !models.empty() && models[0]->ok(); }
  // CoFusion::processFrame.  Buffers are host pointers (pinned or pageable) unless device_ptrs.
  // mask == nullptr: static scene, everything labelled background (CoFusion.cpp:190-197).
  cudaError_t processFrame(const uint8_t* rgb, const float* depth, const uint8_t* mask, bool device_ptrs,
                           float weightMultiplier);
  // processFrame(frame, inPose, weightMultiplier, bootstrap) (CoFusion.h:67-68): inPose == nullptr: regular
  // tracking; inPose && !bootstrap: the camera pose is overridden, nothing is tracked or segmented
  // (CoFusion.cpp:343-345); bootstrap: track, then right-multiply the camera pose by inPose (:219-222).
  cudaError_t processFrameEx(const FrameInput& in, const float* inPose16, bool bootstrap, float weightMultiplier);
  // ---- object sharding over the GPUs of a node: collective; afterwards every rank calls
  // processFrame* for every frame, only rank 0's frame pointers are read, the packed frame [rgb | depth f32 | mask]
  // reaches the other ranks by ONE NCCL broadcast.  A rank processes the models that live in ITS CoFusion object
  // (the application spawns object k on rank shard_owner(k, world)); ranks other than 0 do not process the camera
  // model unless processGlobalModel is set.
  cudaError_t shardInit(int rank, int world, const unsigned char id[128], const char** err);
  FrameShard shard;
  bool processGlobalModel = true;
  const char* lastShardError = "";
  // ---- export (CoFusion.cpp:646-783).  Pose logging must be enabled before the frames of interest.
  void enablePoseLogging(bool on) { poseLogging_ = on; }
  // entries of model i (list position): timestamps + 7 floats each (t.xyz, q.xyzw): camera -> world for the
  // first model, object -> world = cameraPose * modelPose^-1 for the others (CoFusion.cpp:503-518)
  cudaError_t poseLog(size_t i, std::vector<int64_t>* ts, std::vector<float>* p7);
  cudaError_t exportPoses(const char* dir);  // poses-<id>.txt, active and inactive models
  cudaError_t savePly(const char* dir);      // cloud-<id>.ply (binary little endian: x y z r g b nx ny nz radius)
  // spawn an object model (CoFusion::spawnObjectModel, CoFusion.cpp:588-597): created empty; it is
  // initialised by fusing the pixels labelled `id` of the current frame (CoFusion.cpp:265-276)
  cudaError_t spawnObjectModel(unsigned id, const float* initialPose /* 16 or null = background pose */);
  Model* model(size_t i) { return i < models.size() ? models[i].get() : nullptr; }
  size_t numModels() const { return models.size(); }
  int tick() const { return tick_; }
  cudaError_t predict();
  cudaError_t forkModels(const std::vector<Model*>& act);  // CoFusion::predict (CoFusion.cpp:533-545)

  bool batchedTracking = true;  // track all models of a frame in one persistent launch (gn_tiled.cu)
  Context ctx;
  CoFusionParams params;
  std::vector<std::unique_ptr<Model>> models;
  std::vector<std::unique_ptr<ArchivedModel>> inactiveModels;  // CoFusion::inactivateModel keeps the data
  std::vector<std::unique_ptr<Model>> spareModels;  // buffers of lost models (and one made ahead), reused by the next spawn
  cudaError_t acquireModel(unsigned id, float conf, std::unique_ptr<Model>* out);
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
  std::vector<Model*> processed();       // the models this rank processes (all of them unless sharded)
  cudaError_t segmentAndManageModels();  // CoFusion.cpp:227-299
  unsigned char takeNextModelID();       // getNextModelID(true)
  cudaError_t logPoses(int64_t timestamp);
  cudaError_t modelPoseLog(Model* m, std::vector<int64_t>* ts, std::vector<float>* p7);
  cudaError_t poseLogEntries(bool isCamera, const std::vector<int64_t>& lts, const std::vector<int>& lframe,
                             const std::vector<float>& lhost, std::vector<int64_t>* ts, std::vector<float>* p7);
  bool poseLogging_ = false;
  int logFrames_ = 0;  // frames logged so far (index into the camera model's log)
  int tick_ = 1;
  unsigned spawnOffset_ = 0;
  unsigned char nextID_ = 1;  // id 0 went to the global model (CoFusion.cpp:70)
};

}  // namespace cfb
