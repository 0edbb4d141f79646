// pipeline.cuh -- cfb::Context and cfb::Model: the class seam of the reference without OpenGL.
//
//   Context  <->  the per-frame shared state of CoFusion (Core/CoFusion.cpp:171-211): RGB, raw and
//                 bilateral-filtered metric depth, the depth pyramid shared by all models
//                 (Model::GPUSetup::depth_tmp, Model.h:79-80) and the label mask.
//   Model    <->  Core/Model/Model.{h,cpp}: pose, tracker (RGBDOdometry frameToModel), prediction
//                 images (ModelProjection render targets), ICP error map, surfel map.
// All device memory is allocated at construction; every operation is enqueued on the context's
// stream; the only host synchronisation per frame is the pose read-back at the end of tracking.
#pragma once
#include <stdint.h>

#include <vector>

#include "cfb_common.cuh"
#include "odometry.cuh"
#include "surfel_kernels.cuh"

namespace cfb {

class Context {
 public:
  Context(int device, int W, int H, float fx, float fy, float cx, float cy);
  ~Context();
  bool ok() const { return ok_; }

  // CoFusion::processFrame :179-184: upload + bilateral filter.  Host buffers (pinned or pageable).
  cudaError_t uploadFrame(const uint8_t* rgb_hwc, const float* depth, const uint8_t* mask);
  // same, inputs already resident in device memory
  // inputs_follow_stream: the device inputs were produced by work enqueued on `stream` (otherwise they are taken
  // to be complete when the call is made)
  cudaError_t setFrameDevice(const uint8_t* rgb_hwc, const float* depth, const uint8_t* mask, bool inputs_follow_stream = false);
  // Frame ingest on the device (KlgLogReader.cpp:53-84, FrameData.h:38-41): raw u16 depth (x depthScale) and / or an
  // image whose first and third channels are swapped travel as they are (1.54 MB instead of 2.15 MB per VGA frame) and
  // are converted by one kernel.  depth16 == nullptr: `depth` is metric f32 as in uploadFrame.
  cudaError_t uploadFrameRaw(const uint8_t* img_hwc, bool flipColors, const float* depth, const uint16_t* depth16,
                             float depthScale, const uint8_t* mask, bool device_ptrs, bool inputs_follow_stream = false);
  // filterDepth (CoFusion.cpp:567-574) + Model::generateCUDATextures (Model.cpp:319-348)
  cudaError_t preprocess(float depthCutoff);
  cudaError_t sync() { return cudaStreamSynchronize(stream); }

  int device, W, H;
  Intr K;
  cudaStream_t stream = nullptr;
  bool owns_stream = true;
  uint8_t* rgb = nullptr;           // H*W*3 -- the buffer of the CURRENT frame (one of rgbBuf[])
  float* depthRaw = nullptr;        // metric, 0 = invalid (one of depthBuf[])
  // uploadFrame() double-buffers the inputs and copies on its own stream: the H2D transfer of frame
  // t+1 overlaps the fuse / clean / predict kernels of frame t (the copy engine is otherwise idle)
  uint8_t* rgbBuf[2] = {nullptr, nullptr};
  float* depthBuf[2] = {nullptr, nullptr};
  uint8_t* h_rgbBuf[2] = {nullptr, nullptr};
  float* h_depthBuf[2] = {nullptr, nullptr};
  int cur = 0;
  cudaStream_t copyStream = nullptr;
  cudaEvent_t evCopied[2] = {nullptr, nullptr}, evBufferFree[2] = {nullptr, nullptr};
  // The frame side of a frame (device copies / ingest, bilateral filter, depth pyramid) depends on nothing the
  // models produce: it runs on preStream into double buffers, so that frame t+1's 80 us of filtering overlap the
  // many small surfel kernels of frame t, which leave most SMs idle.  `stream` joins at evPre.
  cudaStream_t preStream = nullptr;
  cudaEvent_t evInputs[2] = {nullptr, nullptr}, evPre[2] = {nullptr, nullptr}, evOrder = nullptr, evOrder2 = nullptr;
  // recorded after the tracker launch of a frame: the frame side of the NEXT frame starts behind it.  The tracker
  // needs every SM (one CTA each, all of the shared memory); a bilateral filter in flight at that moment holds it
  // up, while the surfel kernels after the tracker leave room for it.
  cudaEvent_t evTracked = nullptr;
  bool trackedRecorded = false;
  float* depthFilteredBuf[2] = {nullptr, nullptr};
  float* depthPyrBuf[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
  float* depthFiltered = nullptr;   // level 0 of the pyramid of the CURRENT frame
  float* depthPyr[3] = {nullptr, nullptr, nullptr};
  uint8_t* mask = nullptr;          // label image (model ids)
  bool maskIsZero = false;          // the label image is known to be all background (skips the per-frame clear)
  uint16_t* d16Buf[2] = {nullptr, nullptr};    // raw depth of the ingest path (allocated on first use)
  uint16_t* h_d16Buf[2] = {nullptr, nullptr};
  uint8_t* rawImgBuf[2] = {nullptr, nullptr};  // image before the channel swap
  uint8_t* h_rgb = nullptr;         // pinned staging for pageable callers
  float* h_depth = nullptr;
  uint8_t* h_mask = nullptr;
  int launches = 0;                 // kernels launched since the last reset (bench accounting)
  bool keepMask = false;            // true: a frame without mask keeps the previous labels (segmentation on)
  cudaEvent_t evFork = nullptr;     // fork point of the per-model streams
  void* batchScratch = nullptr;     // RGBDOdometry::tiledScratchBytes(), allocated on first multi-model frame

 private:
  cudaError_t beginFrame(bool inputs_follow_stream);
  bool ok_ = false;
};

class Model;
struct TrackParams;
// `for (auto model : models) model->performTracking(...)` (CoFusion.cpp:213-218): one persistent launch for
// up to RGBDOdometry::kMaxBatch models when the tracker mode allows it, the per-model path otherwise
// async: no host synchronisation -- poses stay in the models' device blocks, Model::syncPose() fetches them
cudaError_t trackModels(Context* ctx, Model* const* models, int n, const TrackParams& tp, bool async = false);

struct TrackParams {  // arguments of Model::performTracking (Model.h:128-129)
  int frameToFrameRGB, rgbOnly;
  float icpWeight;
  int pyramid, fastOdom, so3;
  float maxDepthProcessed;
  int force_host_loop;
};

struct ArchivedModel {  // see Model::archive
  unsigned id = 0, count = 0;
  Surfel* surfels = nullptr;  // device, `count` records
  float pose[16];
  float confidenceThreshold = 0.f;
  std::vector<int64_t> poseLogTs;
  std::vector<int> poseLogFrame;
  std::vector<float> poseLogHost;
  ~ArchivedModel();
};

class Model {
 public:
  Model(Context* ctx, unsigned id, float confidenceThreshold, unsigned maxSurfels, bool enableFillIn);
  ~Model();
  bool ok() const { return ok_ && odom.ok(); }

  // Install a model prediction rendered elsewhere (tests; also what combinedPredict produces):
  // AoS float4 vertex(+conf) / normal(+radius) maps in the camera frame and an RGBA8/RGB8 image.
  cudaError_t setPrediction(const float* v4, const float* n4, const uint8_t* img, int channels, bool device_ptrs);
  // RGBDOdometry::initFirstRGB on the current frame (CoFusion.cpp:205)
  cudaError_t initFirstRGB();
  // Model::performTracking (Model.cpp:369-389): initICP (:350-367) + getIncrementalTransformation
  cudaError_t performTracking(const TrackParams& tp);
  // the two halves of performTracking, for the batched tracker (trackModels): everything up to the
  // optimisation (prediction selection + pyramids), and the pose update after it
  cudaError_t prepareTracking(const TrackParams& tp, bool devicePose = false);
  void finishTracking(const float trans[3], const float rot[9]);

  // ---- surfel map (Model.cpp / ModelProjection.cpp; kernels in surfel_kernels.cu)
  cudaError_t initialise(int time, float maxDepthProcessed);                      // Model.cpp:227-272
  cudaError_t predictIndices(int time, float depthCutoff, int timeDelta);         // Model.h:155-157
  cudaError_t fuse(int time, float depthCutoff, float weightMultiplier);          // Model.cpp:408-563
  cudaError_t clean(int time, int timeDelta, float depthCutoff, float outlierCoefficient);  // Model.cpp:565-697
  cudaError_t combinedPredict(float depthCutoff, int time, int maxTime, int timeDelta);     // Model.h:151-153
  cudaError_t performFillIn(bool frameToFrameRGB, bool lost);                     // Model.cpp:901-909
  float computeFusionWeight(float weightMultiplier) const;                        // Model.cpp:391-406
  void refreshCountBound(int time);
  // Model::downloadMap (Model.cpp:868-899): synchronises, copies the live surfels to the host
  cudaError_t downloadMap(float* dst, size_t capacity_surfels, unsigned* count_out);
  cudaError_t uploadMap(const float* src, unsigned count);
  cudaError_t lastCount(unsigned* out);  // Model::lastCount (synchronises)
  const Surfel* surfels() const { return buf[target]; }
  SurfelGeom geom() const { return SurfelGeom{ctx->W, ctx->H, ctx->K.fx, ctx->K.fy, ctx->K.cx, ctx->K.cy}; }

  // ---- pose.  The device block `dpose` is what every kernel of the frame reads (PoseRef); the tracker's
  // epilogue refreshes it, so a frame needs no host synchronisation.  The host copies below are
  // brought up to date lazily by syncPose() (every accessor of the C ABI calls it).
  cudaError_t uploadPose();               // host pose / lastPose -> device block (after a host-side change)
  cudaError_t syncPose();                 // wait for the last asynchronous tracking step, refresh pose / lastPose / stats
  cudaError_t enqueuePoseReadback();      // after an asynchronous tracking launch: D2H of block + stats, event
  const PoseRef poseRef() const { return PoseRef(&dpose->pose); }
  const PoseRef invRef() const { return PoseRef(&dpose->inv); }

  cudaError_t recycle(unsigned id, float confidenceThreshold);  // pooled model -> fresh model (see pipeline.cu)
  cudaError_t archive(ArchivedModel* out);
  cudaError_t fork(cudaEvent_t after);  // per-model stream (see pipeline.cu)
  cudaError_t join();
  cudaStream_t work = nullptr;      // the stream this model's calls are enqueued on (the context's unless forked)
  cudaStream_t mstream = nullptr;
  cudaEvent_t evJoin = nullptr;

  Context* ctx;
  unsigned id;
  float pose[16], lastPose[16];
  PoseDev* dpose = nullptr;       // device
  struct PoseReadback {           // pinned
    PoseDev block;
    TrackStats stats;
  }* h_readback = nullptr;
  cudaEvent_t evPose = nullptr;
  bool poseStale = false;         // the device block is newer than pose / lastPose / odom.stats()
  int cleanTick = 0;              // tick of the last clean() enqueued (bounds the surfel count without a sync)
  // optional pose log (Model::poseLog, Model.h:230-242): one 3x4 pose per logged frame, kept on the device
  float* poseLogDev = nullptr;
  int poseLogCap = 0;
  std::vector<int64_t> poseLogTs;     // timestamp of entry k
  std::vector<int> poseLogFrame;      // index of the frame (camera-model entry) entry k belongs to
  std::vector<float> poseLogHost;     // entries already fetched (12 floats each)
  cudaError_t appendPoseLog(int64_t ts, int frame);
  cudaError_t fetchPoseLog();         // device entries -> poseLogHost (synchronises)
  float confidenceThreshold;
  float maxDepth;               // per-model depth limit (Model::setMaxDepth)
  bool allowsFillIn;
  bool usePrediction = false;   // true once combinedPredict has produced the tracker inputs
  RGBDOdometry odom;
  float* predVertex = nullptr;   // W*H float4: what the tracker consumes (selected prediction)
  float* predNormal = nullptr;   // W*H float4
  uint8_t* predImage = nullptr;  // W*H*4 RGBA8
  float* icpError = nullptr;     // W*H f32 (Model::icpError texture)

  unsigned capacity = 0;         // max surfels
  Surfel* buf[2] = {nullptr, nullptr};
  int target = 0, renderSource = 1;
  Surfel* unstable = nullptr;    // newUnstableBuffer (W*H)
  Surfel* candStaging = nullptr; // per eligible pixel candidate records
  uint32_t* candBest = nullptr;
  uint32_t* winner = nullptr;
  unsigned long long* keys = nullptr;
  IndexMaps indexMaps{};
  SplatMaps splat{};
  FillMaps fill{};
  ScanScratch scan{};
  MapCounters* counters = nullptr;   // device
  MapCounters* h_counters = nullptr; // pinned mirror
  unsigned count_ub = 0;             // host-side upper bound of counters->count

 private:
  std::vector<std::pair<void*, size_t>> zeroed_;  // device buffers the constructor zero-initialised
  bool ok_ = false;
};

}  // namespace cfb
