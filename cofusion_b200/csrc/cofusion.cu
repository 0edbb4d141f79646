// cofusion.cu -- cfb::CoFusion (see cofusion.cuh).
#include "cofusion.cuh"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

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
  }
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
    newModel.reset(new Model(&ctx, id, params.confObjectInit, params.maxSurfels, false));
    if (!newModel->ok()) return cudaErrorMemoryAllocation;
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
          inactiveModels.push_back(std::move(models[j]));
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
  std::unique_ptr<Model> m(new Model(&ctx, id, params.confObjectInit, params.maxSurfels, false));
  if (!m->ok()) return cudaErrorMemoryAllocation;
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

cudaError_t CoFusion::predict() {
  for (auto& m : models) {
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
}  // namespace

cudaError_t CoFusion::processFrame(const uint8_t* rgb, const float* depth, const uint8_t* mask, bool device_ptrs,
                                   float weightMultiplier) {
  Timeline& tl = g_tl;
  if (!tl.init) {
    tl.init = true;
    tl.on = getenv("CFB_TIMELINE") != nullptr;
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
  if (device_ptrs)
    RET_IF(ctx.setFrameDevice(rgb, depth, mask));
  else
    RET_IF(ctx.uploadFrame(rgb, depth, mask));
  RET_IF(ctx.preprocess(params.depthCutoff));
  if (tick_ == 1) {
    RET_IF(models[0]->initialise(tick_, params.maxDepthProcessed));
    RET_IF(models[0]->initFirstRGB());
    // every later frame synchronises on its tracker (after the upload); the first one has none, and the
    // contract is that host buffers may be reused once the call returns
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
    {
      std::vector<Model*> ms;
      for (auto& m : models) ms.push_back(m.get());
      if (batchedTracking) {
        // no host synchronisation: the poses stay on the device, stats / poses are fetched on demand
        RET_IF(trackModels(&ctx, ms.data(), (int)ms.size(), tp, true));
      } else {
        for (Model* m : ms) RET_IF(m->performTracking(tp));
      }
    }
    mark(2);
    if (params.enableMultipleModels) RET_IF(segmentAndManageModels());
    // CoFusion.cpp:347: this prediction only feeds performSegmentation / the (dead) loop-closure
    // block; the fuse stage below uses the index maps and the frame, and the final predict()
    // overwrites every target -> skipped unless asked for.
    if (params.predictBeforeFuse) RET_IF(predict());
    if (!params.rgbOnly) {
      for (auto& m : models) RET_IF(m->predictIndices(tick_, params.maxDepthProcessed, params.timeDelta));
      for (auto& m : models) RET_IF(m->fuse(tick_, params.maxDepthProcessed, weightMultiplier));
      for (auto& m : models) RET_IF(m->predictIndices(tick_, params.maxDepthProcessed, params.timeDelta));
      for (auto& m : models)
        RET_IF(m->clean(tick_, params.timeDelta, params.maxDepthProcessed, params.outlierCoefficient));
    }
  }
  RET_IF(predict());
  tick_++;
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

}  // namespace cfb
