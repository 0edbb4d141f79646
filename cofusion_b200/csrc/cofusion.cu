// cofusion.cu -- cfb::CoFusion (see cofusion.cuh).
#include "cofusion.cuh"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <string>

namespace cfb {

#define RET_IF(e)                       \
  do {                                  \
    cudaError_t e__ = (e);              \
    if (e__ != cudaSuccess) return e__; \
  } while (0)

CoFusion::CoFusion(int device, int W, int H, float fx, float fy, float cx, float cy, const CoFusionParams& p)
    : ctx(device, W, H, fx, fy, cx, cy), params(p) {
  if (!ctx.ok()) return;
  // globalModel: id 0, fill-in enabled (CoFusion.cpp:70)
  models.emplace_back(new Model(&ctx, 0, p.confGlobalInit, p.maxSurfels, true));
  if (p.enableMultipleModels) {
    segmentation.reset(new Segmentation(W, H));
    ctx.keepMask = true;  // textures[MASK] persists between frames (CoFusion.cpp:233)
    // one object model made ahead: the first spawn of a sequence costs no allocation inside a frame
    std::unique_ptr<Model> spare(new Model(&ctx, 255, p.confObjectInit, p.maxSurfels, false));
    if (spare->ok()) spareModels.push_back(std::move(spare));
  }
}

// CoFusion::spawnObjectModel's `std::make_shared<Model>(...)` (CoFusion.cpp:590): from the pool when it has one
cudaError_t CoFusion::acquireModel(unsigned id, float conf, std::unique_ptr<Model>* out) {
  if (!spareModels.empty() && !getenv("CFB_NO_MODEL_POOL")) {  // (the switch is for the equivalence test)
    *out = std::move(spareModels.back());
    spareModels.pop_back();
    return (*out)->recycle(id, conf);
  }
  out->reset(new Model(&ctx, id, conf, params.maxSurfels, false));
  return (*out)->ok() ? cudaSuccess : cudaErrorMemoryAllocation;
}

unsigned char CoFusion::takeNextModelID() {  // CoFusion.cpp:628-645
  const unsigned char next = nextID_;
  while (true) {
    nextID_++;
    bool occupied = false;
    for (auto& m : models)
      if (nextID_ == m->id) occupied = true;
    if (!occupied) break;
  }
  return next;
}

static float seg_max_depth(const SegModelData& d) {  // getMaxDepth lambda (CoFusion.cpp:228)
  return (float)((double)d.depthMean + (double)d.depthStd * 1.2);
}

cudaError_t CoFusion::segmentAndManageModels() {
  if (!segmentation || !segmentation->ok()) return cudaErrorMemoryAllocation;
  if (spawnOffset_ < params.modelSpawnOffset) spawnOffset_++;
  const int n = (int)models.size();
  if (n > SegLimits::kMaxModels) return cudaErrorInvalidValue;
  // the label budget of the CRF kernels caps the number of live models (reference: 255)
  const bool allowNew = spawnOffset_ >= params.modelSpawnOffset && n < SegLimits::kMaxModels;
  unsigned char ids[SegLimits::kMaxModels];
  const float* icp[SegLimits::kMaxModels];
  const float* conf[SegLimits::kMaxModels];
  std::vector<Model*> owners(n);
  for (int i = 0; i < n; ++i) {
    ids[i] = (unsigned char)models[i]->id;
    icp[i] = models[i]->icpError;
    conf[i] = (const float*)models[i]->splat.vertexConf;  // Model::downloadVertexConfTexture (Model.h:189)
    owners[i] = models[i].get();
  }
  lastModelData.assign(n + 1, SegModelData{});
  int cnt = 0;
  bool hasNew = false;
  if (segmentation->slicAheadOf == ctx.rgb) RET_IF(cudaStreamWaitEvent(ctx.stream, ctx.evOrder2, 0));
  ctx.maskIsZero = false;  // the segmentation writes the label image
  RET_IF(segmentation->performSegmentationCRF(ctx.rgb, ctx.depthRaw, n, ids, icp, conf, nextID_, allowNew, params.seg,
                                              ctx.mask, lastModelData.data(), &cnt, &hasNew, ctx.stream));
  ctx.launches += segmentation->launches;
  lastModelData.resize(cnt);
  lastHasNewLabel = hasNew;
  lastSpawnedId = -1;
  lastDeactivated = 0;
  std::unique_ptr<Model> newModel;
  if (hasNew) {  // CoFusion.cpp:243-259, spawnObjectModel :588-597
    const unsigned char id = takeNextModelID();
    RET_IF(acquireModel(id, params.confObjectInit, &newModel));
    RET_IF(newModel->initFirstRGB());
    spawnOffset_ = 0;
    newModel->maxDepth = seg_max_depth(lastModelData.back());
    lastSpawnedId = id;
  }
  for (size_t i = 1; i < models.size(); ++i) models[i]->maxDepth = seg_max_depth(lastModelData[i]);
  if (hasNew) {  // CoFusion.cpp:265-281
    RET_IF(newModel->predictIndices(tick_, params.maxDepthProcessed, params.timeDelta));
    RET_IF(newModel->fuse(tick_, params.maxDepthProcessed, 100.f));
    RET_IF(newModel->clean(tick_, params.timeDelta, params.maxDepthProcessed, params.outlierCoefficient));
    models.push_back(std::move(newModel));
  }
  for (int k = 0; k < cnt && k < n; ++k) {  // lost models (CoFusion.cpp:284-291); unseenCount is never reset
    const SegModelData& m = lastModelData[k];
    if (m.superPixelCount <= 0 && m.id != 0) {
      for (size_t j = 0; j < models.size(); ++j)
        if (models[j].get() == owners[k]) {
          // the data stays (archive), the buffers serve the next spawn
          std::unique_ptr<ArchivedModel> a(new ArchivedModel());
          RET_IF(models[j]->archive(a.get()));
          inactiveModels.push_back(std::move(a));
          spareModels.push_back(std::move(models[j]));
          models.erase(models.begin() + j);
          lastDeactivated++;
          break;
        }
    }
  }
  // positional indexing into modelData AFTER the list changed, as the reference (CoFusion.cpp:294-298)
  for (size_t i = 1; i < models.size() && i < lastModelData.size(); ++i) {
    const float oldConf = models[i]->confidenceThreshold;
    const float a = lastModelData[i].avgConfidence;
    models[i]->confidenceThreshold = fminf(fmaxf(oldConf, a), 9.0f);
  }
  return cudaSuccess;
}

cudaError_t CoFusion::spawnObjectModel(unsigned id, const float* initialPose) {
  std::unique_ptr<Model> m;
  RET_IF(acquireModel(id, params.confObjectInit, &m));
  RET_IF(models[0]->syncPose());
  const float* src = initialPose ? initialPose : models[0]->pose;
  memcpy(m->pose, src, sizeof(m->pose));
  memcpy(m->lastPose, src, sizeof(m->lastPose));
  RET_IF(m->uploadPose());
  RET_IF(m->initFirstRGB());  // CoFusion.cpp:596
  // newModel->predictIndices / fuse (weight 100) / clean against the current frame (CoFusion.cpp:265-276)
  RET_IF(m->predictIndices(tick_, params.maxDepthProcessed, params.timeDelta));
  RET_IF(m->fuse(tick_, params.maxDepthProcessed, 100.f));
  RET_IF(m->clean(tick_, params.timeDelta, params.maxDepthProcessed, params.outlierCoefficient));
  models.push_back(std::move(m));
  return cudaSuccess;
}

std::vector<Model*> CoFusion::processed() {
  std::vector<Model*> v;
  for (size_t i = 0; i < models.size(); ++i)
    if (i > 0 || processGlobalModel) v.push_back(models[i].get());
  return v;
}

// The fuse / clean / predict stages of a frame are independent per model (CoFusion.cpp:465-488, :536-542): with
// more than one model each runs on its own stream from here to the end of the frame (joined after predict()).
cudaError_t CoFusion::forkModels(const std::vector<Model*>& act) {
  if (act.size() < 2 || !batchedTracking) return cudaSuccess;
  RET_IF(cudaEventRecord(ctx.evFork, ctx.stream));
  for (Model* m : act) RET_IF(m->fork(ctx.evFork));
  return cudaSuccess;
}

cudaError_t CoFusion::predict() {
  for (Model* m : processed()) {
    // lastFrameRecovery is never set without loop closure -> maxTime = tick (CoFusion.cpp:538)
    RET_IF(m->combinedPredict(params.maxDepthProcessed, tick_, tick_, params.timeDelta));
    RET_IF(m->performFillIn(params.frameToFrameRGB != 0, false));
  }
  return cudaSuccess;
}

// Optional timeline (tools only, CFB_TIMELINE=1): device time between four points of a frame and the
// host time spent enqueueing each section, averaged over 200 frames and printed to stderr.
namespace {
struct Timeline {
  bool on = false, init = false;
  cudaEvent_t ev[2][4];
  double host[4] = {0, 0, 0, 0}, devms[4] = {0, 0, 0, 0};
  double hprev = 0;
  int frames = 0, cur = 0;
  static double now() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
  }
};
Timeline g_tl;

// Finer device timeline (tools only, CFB_TIMELINE=2): events on the pipeline stream between the stages of a
// single-model frame; the mean device time of every section over 200 frames goes to stderr.
struct FineTimeline {
  static constexpr int kMax = 12, kRing = 64;  // the host may run dozens of frames ahead of the device
  bool on = false, init = false;
  cudaEvent_t ev[kRing][kMax];
  const char* name[kMax];
  int n[kRing] = {0};
  double sum[kMax] = {0}, gap = 0;
  int frames = 0, cur = 0;
  void begin() {
    if (!init) {
      init = true;
      const char* e = getenv("CFB_TIMELINE");
      on = e && atoi(e) == 2;
      if (on)
        for (int a = 0; a < kRing; ++a)
          for (int b = 0; b < kMax; ++b) cudaEventCreate(&ev[a][b]);
    }
    if (!on) return;
    cur = (cur + 1) % kRing;
    n[cur] = 0;
  }
  void mark(cudaStream_t s, const char* what) {
    if (!on || n[cur] >= kMax) return;
    name[n[cur]] = what;
    cudaEventRecord(ev[cur][n[cur]++], s);
  }
  void end() {
    if (!on) return;
    const int old = (cur + 1) % kRing, nxt = (cur + 2) % kRing;  // the oldest frames of the ring
    if (n[old] < 2 || n[nxt] < 1 || cudaEventQuery(ev[nxt][0]) != cudaSuccess) return;
    float ms;
    for (int k = 0; k + 1 < n[old]; ++k)
      if (cudaEventElapsedTime(&ms, ev[old][k], ev[old][k + 1]) == cudaSuccess) sum[k] += ms;
    if (cudaEventElapsedTime(&ms, ev[old][n[old] - 1], ev[nxt][0]) == cudaSuccess) gap += ms;
    if (++frames == 200) {
      fprintf(stderr, "[cfb timeline2, us]");
      double tot = 0;
      for (int k = 0; k + 1 < n[old]; ++k) {
        fprintf(stderr, " %s->%s %.1f |", name[k], name[k + 1], sum[k] * 5.0);
        tot += sum[k];
        sum[k] = 0;
      }
      fprintf(stderr, " to next frame %.1f | frame %.1f\n", gap * 5.0, (tot + gap) * 5.0);
      gap = 0;
      frames = 0;
    }
  }
};
FineTimeline g_ft;
}  // namespace

cudaError_t CoFusion::processFrame(const uint8_t* rgb, const float* depth, const uint8_t* mask, bool device_ptrs,
                                   float weightMultiplier) {
  FrameInput in;
  in.rgb = rgb;
  in.depth = depth;
  in.mask = mask;
  in.device_ptrs = device_ptrs;
  in.timestamp = (int64_t)(tick_ - 1) * 33;
  return processFrameEx(in, nullptr, false, weightMultiplier);
}

cudaError_t CoFusion::shardInit(int rank, int world, const unsigned char id[128], const char** err) {
  if (params.enableMultipleModels) {
    *err = "shard_init: the sharded path takes external label masks (enableMultipleModels = 0)";
    return cudaErrorInvalidValue;
  }
  RET_IF(cudaSetDevice(ctx.device));
  if (shard.init(rank, world, id, (size_t)ctx.W * ctx.H * 8, err) != 0) return cudaErrorUnknown;
  processGlobalModel = rank == 0;
  return cudaSuccess;
}

cudaError_t CoFusion::processFrameEx(const FrameInput& in_, const float* inPose, bool bootstrap, float weightMultiplier) {
  FrameInput in = in_;
  if (shard.active()) {
    // pack [rgb 3P | depth f32 4P | mask P] on the root, one broadcast, then every rank sees a device-resident frame
    const size_t P = (size_t)ctx.W * ctx.H;
    uint8_t* buf = nullptr;
    cudaEvent_t freeEvt = nullptr;
    RET_IF(shard.acquire(ctx.stream, &buf, &freeEvt));
    cudaEvent_t ready = nullptr;
    if (shard.rank() == 0) {
      if (!in.rgb || !in.depth || in.depth16 || in.flipColors) return cudaErrorInvalidValue;  // f32 RGB frames on the sharded path
      const cudaMemcpyKind k = in.device_ptrs ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
      RET_IF(cudaStreamWaitEvent(ctx.copyStream, freeEvt, 0));
      RET_IF(cudaMemcpyAsync(buf, in.rgb, P * 3, k, ctx.copyStream));
      RET_IF(cudaMemcpyAsync(buf + 3 * P, in.depth, P * 4, k, ctx.copyStream));
      if (in.mask)
        RET_IF(cudaMemcpyAsync(buf + 7 * P, in.mask, P, k, ctx.copyStream));
      else
        RET_IF(cudaMemsetAsync(buf + 7 * P, 0, P, ctx.copyStream));
      RET_IF(cudaEventRecord(ctx.evCopied[0], ctx.copyStream));
      ready = ctx.evCopied[0];
    }
    RET_IF(shard.broadcast(ready, ctx.stream, &lastShardError));
    in.rgb = buf;
    in.depth = (const float*)(buf + 3 * P);
    in.depth16 = nullptr;
    in.flipColors = false;
    in.mask = buf + 7 * P;
    in.device_ptrs = true;
    if (shard.rank() == 0 && !in_.device_ptrs) RET_IF(cudaEventSynchronize(ctx.evCopied[0]));  // host buffers are free again
  }
  if (!in.rgb || (!in.depth && !in.depth16) || (bootstrap && !inPose)) return cudaErrorInvalidValue;
  const bool device_ptrs = in.device_ptrs;
  Timeline& tl = g_tl;
  if (!tl.init) {
    tl.init = true;
    tl.on = getenv("CFB_TIMELINE") != nullptr && atoi(getenv("CFB_TIMELINE")) == 1;
    if (tl.on)
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 4; ++b) cudaEventCreate(&tl.ev[a][b]);
  }
  double h[4] = {0, 0, 0, 0};
  auto mark = [&](int k) {
    if (!tl.on) return;
    h[k] = Timeline::now();
    cudaEventRecord(tl.ev[tl.cur][k], ctx.stream);
  };
  mark(0);
  g_ft.begin();
  g_ft.mark(ctx.stream, "start");
  pdl_set(models.size() == 1 && !params.enableMultipleModels);  // see cfb_common.cuh
  // device inputs produced on the pipeline stream itself (the broadcast of the sharded path, a caller's own stream)
  // are ordered after it; otherwise the frame side starts right away, next to the previous frame's tail
  RET_IF(ctx.uploadFrameRaw(in.rgb, in.flipColors, in.depth, in.depth16, in.depthScale, in.mask, device_ptrs,
                            shard.active() || !ctx.owns_stream));
  RET_IF(ctx.preprocess(params.depthCutoff));
  if (params.enableMultipleModels && segmentation && tick_ > 1 && !(inPose && !bootstrap)) {
    // the super-pixels of this frame, ahead of time on the frame-side stream: they overlap the previous frame's
    // surfel kernels (or share the SMs with the tracker, whose CTAs leave most issue slots idle)
    RET_IF(segmentation->slic(ctx.rgb, ctx.preStream));
    RET_IF(cudaEventRecord(ctx.evOrder2, ctx.preStream));
    segmentation->slicAheadOf = ctx.rgb;
  }
  if (tick_ == 1) {
    if (processGlobalModel) {
      RET_IF(models[0]->initialise(tick_, params.maxDepthProcessed));
      RET_IF(models[0]->initFirstRGB());
    }
    // every later frame synchronises on its tracker (after the upload); the first one has none, and the
    // contract is that host buffers may be reused once the call returns
  } else if (inPose && !bootstrap) {
    // pose provided by the caller: Model::overridePose on the camera model, no tracking, no segmentation
    // (CoFusion.cpp:343-345); the fuse block below still runs (trackingOk stays true, :463)
    Model* g = models[0].get();
    RET_IF(g->syncPose());
    memcpy(g->pose, inPose, sizeof(g->pose));
    memcpy(g->lastPose, inPose, sizeof(g->lastPose));
    RET_IF(g->uploadPose());
    if (params.predictBeforeFuse) RET_IF(predict());
    if (!params.rgbOnly) {
      const std::vector<Model*> act = processed();
      RET_IF(forkModels(act));
      for (Model* m : act) RET_IF(m->predictIndices(tick_, params.maxDepthProcessed, params.timeDelta));
      for (Model* m : act) RET_IF(m->fuse(tick_, params.maxDepthProcessed, weightMultiplier));
      for (Model* m : act) RET_IF(m->predictIndices(tick_, params.maxDepthProcessed, params.timeDelta));
      for (Model* m : act) RET_IF(m->clean(tick_, params.timeDelta, params.maxDepthProcessed, params.outlierCoefficient));
    }
  } else {
    TrackParams tp;
    tp.frameToFrameRGB = params.frameToFrameRGB;
    tp.rgbOnly = params.rgbOnly;
    tp.icpWeight = params.icpWeight;
    tp.pyramid = params.pyramid;
    tp.fastOdom = params.fastOdom;
    tp.so3 = params.so3;
    tp.maxDepthProcessed = params.maxDepthProcessed;
    tp.force_host_loop = 0;
    mark(1);
    g_ft.mark(ctx.stream, "frame-side");
    {
      std::vector<Model*> ms = processed();
      if (ms.empty()) {
      } else if (batchedTracking) {
        // no host synchronisation: the poses stay on the device, stats / poses are fetched on demand
        RET_IF(trackModels(&ctx, ms.data(), (int)ms.size(), tp, true));
      } else {
        for (Model* m : ms) RET_IF(m->performTracking(tp));
      }
    }
    mark(2);
    RET_IF(cudaEventRecord(ctx.evTracked, ctx.stream));  // the next frame's frame side starts behind the tracker
    ctx.trackedRecorded = true;
    g_ft.mark(ctx.stream, "tracked");
    if (bootstrap) {  // globalModel->overridePose(globalModel->getPose() * inPose) (CoFusion.cpp:219-222)
      Model* g = models[0].get();
      RET_IF(g->syncPose());
      float r[16];
      for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
          float acc = 0.f;
          for (int k = 0; k < 4; ++k) acc += g->pose[i * 4 + k] * inPose[k * 4 + j];
          r[i * 4 + j] = acc;
        }
      memcpy(g->pose, r, sizeof(r));
      memcpy(g->lastPose, r, sizeof(r));
      RET_IF(g->uploadPose());
    }
    if (params.enableMultipleModels) RET_IF(segmentAndManageModels());
    // CoFusion.cpp:347: this prediction only feeds performSegmentation / the (dead) loop-closure
    // block; the fuse stage below uses the index maps and the frame, and the final predict()
    // overwrites every target -> skipped unless asked for.
    if (params.predictBeforeFuse) RET_IF(predict());
    if (!params.rgbOnly) {
      const std::vector<Model*> act = processed();
      RET_IF(forkModels(act));
      for (Model* m : act) RET_IF(m->predictIndices(tick_, params.maxDepthProcessed, params.timeDelta));
      g_ft.mark(ctx.stream, "indices");
      for (Model* m : act) RET_IF(m->fuse(tick_, params.maxDepthProcessed, weightMultiplier));
      g_ft.mark(ctx.stream, "fused");
      for (Model* m : act) RET_IF(m->predictIndices(tick_, params.maxDepthProcessed, params.timeDelta));
      g_ft.mark(ctx.stream, "indices2");
      for (Model* m : act) RET_IF(m->clean(tick_, params.timeDelta, params.maxDepthProcessed, params.outlierCoefficient));
      g_ft.mark(ctx.stream, "cleaned");
    }
  }
  RET_IF(predict());
  g_ft.mark(ctx.stream, "predicted");
  g_ft.end();
  for (auto& m : models) RET_IF(m->join());
  tick_++;
  if (poseLogging_) RET_IF(logPoses(in.timestamp));
  // nothing in this frame waited for the device; the contract is that host buffers may be reused once the
  // call returns, so wait for this frame's upload (it only depends on the frame before the previous one)
  if (!device_ptrs) RET_IF(cudaEventSynchronize(ctx.evCopied[ctx.cur]));
  mark(3);
  if (tl.on && tick_ > 3) {
    // the previous frame's events are complete by now (this frame synchronised after them)
    const int prev = tl.cur ^ 1;
    float ms;
    if (tl.hprev > 0) {
      for (int k = 0; k < 3; ++k)
        if (cudaEventElapsedTime(&ms, tl.ev[prev][k], tl.ev[prev][k + 1]) == cudaSuccess) tl.devms[k] += ms;
      if (cudaEventElapsedTime(&ms, tl.ev[prev][3], tl.ev[tl.cur][0]) == cudaSuccess) tl.devms[3] += ms;
      tl.host[3] += h[0] - tl.hprev;
      for (int k = 0; k < 3; ++k) tl.host[k] += h[k + 1] - h[k];
      tl.frames++;
    }
    tl.hprev = h[3];
    if (tl.frames == 200) {
      fprintf(stderr,
              "[cfb timeline, ms] device: start->track %.3f  track(+sync) %.3f  fuse..predict %.3f  frame gap %.3f | "
              "host: pre %.3f  track+sync %.3f  post %.3f  between calls %.3f\n",
              tl.devms[0] / 200, tl.devms[1] / 200, tl.devms[2] / 200, tl.devms[3] / 200, tl.host[0] / 200,
              tl.host[1] / 200, tl.host[2] / 200, tl.host[3] / 200);
      tl.frames = 0;
      for (int k = 0; k < 4; ++k) tl.devms[k] = tl.host[k] = 0;
    }
  }
  if (tl.on) tl.cur ^= 1;
  return cudaSuccess;
}

// ------------------------------------------------------------------------------------------------ export
cudaError_t CoFusion::logPoses(int64_t timestamp) {
  // CoFusion.cpp:503-518: one entry per active model and frame.  The 3x4 pose is copied device to device from the
  // model's pose block (the frame is not waited for); quaternions are formed when the log is read.
  for (auto& m : models) RET_IF(m->appendPoseLog(timestamp, logFrames_));
  logFrames_++;
  return cudaSuccess;
}

namespace {
void rigid_inverse(const float* T, float* Ti) {  // 3x4 row-major
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) Ti[r * 4 + c] = T[c * 4 + r];
  for (int r = 0; r < 3; ++r) Ti[r * 4 + 3] = -(Ti[r * 4] * T[3] + Ti[r * 4 + 1] * T[7] + Ti[r * 4 + 2] * T[11]);
}
void rigid_mul(const float* A, const float* B, float* C) {  // 3x4 * 3x4 (implicit last row 0 0 0 1)
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 4; ++c)
      C[r * 4 + c] = A[r * 4] * B[c] + A[r * 4 + 1] * B[4 + c] + A[r * 4 + 2] * B[8 + c] + (c == 3 ? A[r * 4 + 3] : 0.f);
  }
}
void quaternion_of(const float* T, float* q) {  // Eigen::Quaternionf(Matrix3f): x y z w
  const float m00 = T[0], m01 = T[1], m02 = T[2], m10 = T[4], m11 = T[5], m12 = T[6], m20 = T[8], m21 = T[9], m22 = T[10];
  float t = m00 + m11 + m22;
  if (t > 0.f) {
    t = sqrtf(t + 1.0f);
    q[3] = 0.5f * t;
    t = 0.5f / t;
    q[0] = (m21 - m12) * t;
    q[1] = (m02 - m20) * t;
    q[2] = (m10 - m01) * t;
  } else {
    const float M[3][3] = {{m00, m01, m02}, {m10, m11, m12}, {m20, m21, m22}};
    int i = 0;
    if (m11 > m00) i = 1;
    if (m22 > M[i][i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrtf(M[i][i] - M[j][j] - M[k][k] + 1.0f);
    q[i] = 0.5f * t;
    t = 0.5f / t;
    q[3] = (M[k][j] - M[j][k]) * t;
    q[j] = (M[j][i] + M[i][j]) * t;
    q[k] = (M[k][i] + M[i][k]) * t;
  }
}
}  // namespace

cudaError_t CoFusion::poseLogEntries(bool isCamera, const std::vector<int64_t>& lts, const std::vector<int>& lframe,
                                     const std::vector<float>& lhost, std::vector<int64_t>* ts, std::vector<float>* p7) {
  Model* g = models[0].get();
  RET_IF(g->fetchPoseLog());
  const size_t n = lts.size();
  if (ts) *ts = lts;
  if (p7) {
    p7->resize(n * 7);
    for (size_t k = 0; k < n; ++k) {
      const float* P = lhost.data() + k * 12;
      float T[12];
      if (isCamera) {
        memcpy(T, P, sizeof(T));
      } else {  // object -> world = cameraPose * modelPose^-1 of the same frame
        const int f = lframe[k];
        float inv[12];
        rigid_inverse(P, inv);
        rigid_mul(g->poseLogHost.data() + (size_t)f * 12, inv, T);
      }
      float* o = p7->data() + k * 7;
      o[0] = T[3];
      o[1] = T[7];
      o[2] = T[11];
      quaternion_of(T, o + 3);
    }
  }
  return cudaSuccess;
}

cudaError_t CoFusion::modelPoseLog(Model* m, std::vector<int64_t>* ts, std::vector<float>* p7) {
  RET_IF(m->fetchPoseLog());
  return poseLogEntries(m == models[0].get(), m->poseLogTs, m->poseLogFrame, m->poseLogHost, ts, p7);
}

cudaError_t CoFusion::poseLog(size_t i, std::vector<int64_t>* ts, std::vector<float>* p7) {
  if (i >= models.size()) return cudaErrorInvalidValue;
  return modelPoseLog(models[i].get(), ts, p7);
}

cudaError_t CoFusion::exportPoses(const char* dir) {  // CoFusion.cpp:758-783
  auto write = [&](unsigned id, const std::vector<int64_t>& ts, const std::vector<float>& p) -> cudaError_t {
    if (ts.empty()) return cudaSuccess;  // Model::isLoggingPoses
    const std::string fn = std::string(dir) + "/poses-" + std::to_string(id) + ".txt";
    FILE* f = fopen(fn.c_str(), "w");
    if (!f) return cudaErrorInvalidValue;
    for (size_t k = 0; k < ts.size(); ++k) {
      fprintf(f, "%lld", (long long)ts[k]);
      for (int q = 0; q < 7; ++q) fprintf(f, " %.9g", p[k * 7 + q]);
      fprintf(f, "\n");
    }
    fclose(f);
    return cudaSuccess;
  };
  std::vector<int64_t> ts;
  std::vector<float> p;
  for (auto& m : models) {
    RET_IF(modelPoseLog(m.get(), &ts, &p));
    RET_IF(write(m->id, ts, p));
  }
  for (auto& a : inactiveModels) {
    RET_IF(poseLogEntries(false, a->poseLogTs, a->poseLogFrame, a->poseLogHost, &ts, &p));
    RET_IF(write(a->id, ts, p));
  }
  return cudaSuccess;
}

cudaError_t CoFusion::savePly(const char* dir) {  // CoFusion.cpp:646-756
  Model* g = models[0].get();
  RET_IF(g->syncPose());
  for (auto& mp : models) {
    Model* m = mp.get();
    RET_IF(m->syncPose());
    unsigned n = 0;
    RET_IF(m->lastCount(&n));
    std::vector<float> map((size_t)n * 12);
    if (n) RET_IF(m->downloadMap(map.data(), n, &n));
    // Tp = globalPose * modelPose^-1 for the points; its inverse transpose (= its rotation) for the normals --
    // the reference initialises Tn from itself (CoFusion.cpp:703), the intended matrix is used here
    float inv[12], Tp[12];
    rigid_inverse(m->pose, inv);
    rigid_mul(g->pose, inv, Tp);
    size_t valid = 0;
    for (unsigned i = 0; i < n; ++i) valid += map[(size_t)i * 12 + 3] > m->confidenceThreshold;
    const std::string fn = std::string(dir) + "/cloud-" + std::to_string(m->id) + ".ply";
    FILE* f = fopen(fn.c_str(), "wb");
    if (!f) return cudaErrorInvalidValue;
    fprintf(f,
            "ply\nformat binary_little_endian 1.0\nelement vertex %zu\nproperty float x\nproperty float y\nproperty float z"
            "\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nproperty float nx\nproperty float ny"
            "\nproperty float nz\nproperty float radius\nend_header\n",
            valid);
    for (unsigned i = 0; i < n; ++i) {
      const float* s = map.data() + (size_t)i * 12;
      if (!(s[3] > m->confidenceThreshold)) continue;
      float rec[3], nor[3];
      for (int r = 0; r < 3; ++r) {
        rec[r] = Tp[r * 4] * s[0] + Tp[r * 4 + 1] * s[1] + Tp[r * 4 + 2] * s[2] + Tp[r * 4 + 3];
        nor[r] = -(Tp[r * 4] * s[8] + Tp[r * 4 + 1] * s[9] + Tp[r * 4 + 2] * s[10]);
      }
      const int col = (int)s[4];
      const unsigned char rgb[3] = {(unsigned char)(col >> 16 & 0xFF), (unsigned char)(col >> 8 & 0xFF), (unsigned char)(col & 0xFF)};
      fwrite(rec, sizeof(float), 3, f);
      fwrite(rgb, 1, 3, f);
      fwrite(nor, sizeof(float), 3, f);
      fwrite(&s[11], sizeof(float), 1, f);
    }
    fclose(f);
  }
  return cudaSuccess;
}

}  // namespace cfb
