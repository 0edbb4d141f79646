// cofusion.cu -- cfb::CoFusion (see cofusion.cuh).
#include "cofusion.cuh"

#include <string.h>

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
  lastStats.resize(1);
}

cudaError_t CoFusion::spawnObjectModel(unsigned id, const float* initialPose) {
  std::unique_ptr<Model> m(new Model(&ctx, id, params.confObjectInit, params.maxSurfels, false));
  if (!m->ok()) return cudaErrorMemoryAllocation;
  const float* src = initialPose ? initialPose : models[0]->pose;
  memcpy(m->pose, src, sizeof(m->pose));
  memcpy(m->lastPose, src, sizeof(m->lastPose));
  RET_IF(m->initFirstRGB());  // CoFusion.cpp:596
  // newModel->predictIndices / fuse (weight 100) / clean against the current frame (CoFusion.cpp:265-276)
  RET_IF(m->predictIndices(tick_, params.maxDepthProcessed, params.timeDelta));
  RET_IF(m->fuse(tick_, params.maxDepthProcessed, 100.f));
  RET_IF(m->clean(tick_, params.timeDelta, params.maxDepthProcessed, params.outlierCoefficient));
  models.push_back(std::move(m));
  lastStats.resize(models.size());
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

cudaError_t CoFusion::processFrame(const uint8_t* rgb, const float* depth, const uint8_t* mask, bool device_ptrs,
                                   float weightMultiplier) {
  if (device_ptrs)
    RET_IF(ctx.setFrameDevice(rgb, depth, mask));
  else
    RET_IF(ctx.uploadFrame(rgb, depth, mask));
  RET_IF(ctx.preprocess(params.depthCutoff));
  if (tick_ == 1) {
    RET_IF(models[0]->initialise(tick_, params.maxDepthProcessed));
    RET_IF(models[0]->initFirstRGB());
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
    for (size_t i = 0; i < models.size(); ++i) {
      RET_IF(models[i]->performTracking(tp));
      lastStats[i] = models[i]->odom.stats();
    }
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
  return cudaSuccess;
}

}  // namespace cfb
