// pipeline.cu -- cfb::Context / cfb::Model (see pipeline.cuh).
#include "pipeline.cuh"

#include <stdlib.h>

#include <float.h>
#include <math.h>
#include <string.h>

#include "image_kernels.cuh"
#include "pose_math.cuh"

namespace cfb {

#define RET_IF(e)                       \
  do {                                  \
    cudaError_t e__ = (e);              \
    if (e__ != cudaSuccess) return e__; \
  } while (0)

namespace {
template <class T>
bool dalloc(T** p, size_t n, std::vector<std::pair<void*, size_t>>* reg = nullptr) {
  if (cudaMalloc((void**)p, n * sizeof(T)) != cudaSuccess || cudaMemset(*p, 0, n * sizeof(T)) != cudaSuccess) return false;
  if (reg) reg->push_back({(void*)*p, n * sizeof(T)});
  return true;
}
}  // namespace

Context::Context(int dev, int w, int h, float fx, float fy, float cx, float cy)
    : device(dev), W(w), H(h), K{fx, fy, cx, cy} {
  if (cudaSetDevice(dev) != cudaSuccess) return;
  if (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess) return;
  const size_t n = (size_t)W * H;
  bool good = dalloc(&rgbBuf[0], n * 3) && dalloc(&rgbBuf[1], n * 3) && dalloc(&depthBuf[0], n) && dalloc(&depthBuf[1], n) &&
              dalloc(&mask, n);
  for (int k = 0; k < 2; ++k) {
    good = good && dalloc(&depthFilteredBuf[k], n) && dalloc(&depthPyrBuf[k][1], n / 4) && dalloc(&depthPyrBuf[k][2], n / 16);
    depthPyrBuf[k][0] = depthFilteredBuf[k];
    good = good && cudaEventCreateWithFlags(&evInputs[k], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&evPre[k], cudaEventDisableTiming) == cudaSuccess;
  }
  good = good && cudaEventCreateWithFlags(&evOrder, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&evTracked, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&evOrder2, cudaEventDisableTiming) == cudaSuccess &&
         cudaStreamCreateWithFlags(&preStream, cudaStreamNonBlocking) == cudaSuccess;
  rgb = rgbBuf[0];
  depthRaw = depthBuf[0];
  depthFiltered = depthFilteredBuf[0];
  for (int i = 0; i < 3; ++i) depthPyr[i] = depthPyrBuf[0][i];
  for (int k = 0; k < 2; ++k)
    good = good && cudaMallocHost(&h_rgbBuf[k], n * 3) == cudaSuccess && cudaMallocHost(&h_depthBuf[k], n * 4) == cudaSuccess &&
           cudaEventCreateWithFlags(&evCopied[k], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&evBufferFree[k], cudaEventDisableTiming) == cudaSuccess;
  good = good && cudaEventCreateWithFlags(&evFork, cudaEventDisableTiming) == cudaSuccess;
  h_rgb = h_rgbBuf[0];
  h_depth = h_depthBuf[0];
  good = good && cudaMallocHost(&h_mask, n) == cudaSuccess &&
         cudaStreamCreateWithFlags(&copyStream, cudaStreamNonBlocking) == cudaSuccess;
  ok_ = good;
}

Context::~Context() {
  if (stream) cudaStreamSynchronize(stream);
  if (copyStream) cudaStreamSynchronize(copyStream);
  if (preStream) cudaStreamSynchronize(preStream);
  for (int k = 0; k < 2; ++k) {
    cudaFree(depthFilteredBuf[k]);
    cudaFree(depthPyrBuf[k][1]);
    cudaFree(depthPyrBuf[k][2]);
    if (evInputs[k]) cudaEventDestroy(evInputs[k]);
    if (evPre[k]) cudaEventDestroy(evPre[k]);
    cudaFree(d16Buf[k]);
    cudaFreeHost(h_d16Buf[k]);
    cudaFree(rawImgBuf[k]);
    cudaFree(rgbBuf[k]);
    cudaFree(depthBuf[k]);
    cudaFreeHost(h_rgbBuf[k]);
    cudaFreeHost(h_depthBuf[k]);
    if (evCopied[k]) cudaEventDestroy(evCopied[k]);
    if (evBufferFree[k]) cudaEventDestroy(evBufferFree[k]);
  }
  if (evFork) cudaEventDestroy(evFork);
  if (evOrder) cudaEventDestroy(evOrder);
  if (evTracked) cudaEventDestroy(evTracked);
  if (evOrder2) cudaEventDestroy(evOrder2);
  if (preStream) cudaStreamDestroy(preStream);
  cudaFree(mask);
  cudaFreeHost(h_mask);
  cudaFree(batchScratch);
  if (copyStream) cudaStreamDestroy(copyStream);
  if (stream && owns_stream) cudaStreamDestroy(stream);
}

// Start a frame: flip the double buffers.  Everything enqueued so far on the pipeline stream (the whole previous
// frame) is what still reads the buffers being left; the ones being entered were last read by the frame before
// that, whose end evBufferFree[] of the new index marks.
cudaError_t Context::beginFrame(bool inputs_follow_stream) {
  RET_IF(cudaEventRecord(evBufferFree[cur], stream));
  cur ^= 1;
  rgb = rgbBuf[cur];
  depthRaw = depthBuf[cur];
  depthFiltered = depthFilteredBuf[cur];
  for (int i = 0; i < 3; ++i) depthPyr[i] = depthPyrBuf[cur][i];
  RET_IF(cudaStreamWaitEvent(copyStream, evBufferFree[cur], 0));
  RET_IF(cudaStreamWaitEvent(preStream, evBufferFree[cur], 0));
  {
    // (with the segmentation in the loop the frame has a host synchronisation in its middle and the early start wins:
    //  933 vs 903 frames/s on the 4-object scene; without it 0.625 vs 0.633 ms per frame the other way round)
    if (trackedRecorded && !keepMask) RET_IF(cudaStreamWaitEvent(preStream, evTracked, 0));
  }
  if (inputs_follow_stream) {
    RET_IF(cudaEventRecord(evOrder, stream));
    RET_IF(cudaStreamWaitEvent(preStream, evOrder, 0));
  }
  return cudaSuccess;
}

cudaError_t Context::uploadFrame(const uint8_t* rgb_h, const float* depth_h, const uint8_t* mask_h) {
  const size_t n = (size_t)W * H;
  RET_IF(beginFrame(false));
  // Pinned callers are copied straight from their buffers; pageable ones are staged through the
  // context's pinned buffers so that the copy is truly asynchronous either way.
  cudaPointerAttributes a;
  auto pinned = [&](const void* p) {
    return cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type == cudaMemoryTypeHost;
  };
  const uint8_t* r = rgb_h;
  const float* d = depth_h;
  const bool pr = pinned(rgb_h), pd = pinned(depth_h);
  if (!pr || !pd) {
    cudaGetLastError();
    RET_IF(cudaEventSynchronize(evCopied[cur]));  // the staging buffers' previous transfer (two frames ago) is done
  }
  if (!pr) {
    memcpy(h_rgbBuf[cur], rgb_h, n * 3);
    r = h_rgbBuf[cur];
  }
  if (!pd) {
    memcpy(h_depthBuf[cur], depth_h, n * 4);
    d = h_depthBuf[cur];
  }
  RET_IF(cudaMemcpyAsync(rgb, r, n * 3, cudaMemcpyHostToDevice, copyStream));
  RET_IF(cudaMemcpyAsync(depthRaw, d, n * 4, cudaMemcpyHostToDevice, copyStream));
  RET_IF(cudaEventRecord(evCopied[cur], copyStream));
  RET_IF(cudaStreamWaitEvent(preStream, evCopied[cur], 0));
  RET_IF(cudaStreamWaitEvent(stream, evCopied[cur], 0));
  if (mask_h) {
    const uint8_t* m = mask_h;
    if (!pinned(mask_h)) {
      cudaGetLastError();
      RET_IF(cudaStreamSynchronize(stream));  // single staging buffer for the (small, optional) label image
      memcpy(h_mask, mask_h, n);
      m = h_mask;
    }
    RET_IF(cudaMemcpyAsync(mask, m, n, cudaMemcpyHostToDevice, stream));
    maskIsZero = false;
  } else if (!keepMask) {
    // static scene: everything is background (CoFusion.cpp:190-197)
    if (!maskIsZero) RET_IF(cudaMemsetAsync(mask, 0, n, stream));  // (still zero from the last frame otherwise)
    maskIsZero = true;
  }
  return cudaSuccess;
}

cudaError_t Context::uploadFrameRaw(const uint8_t* img, bool flip, const float* depth, const uint16_t* depth16, float scale,
                                    const uint8_t* mask_p, bool device_ptrs, bool inputs_follow_stream) {
  if (!depth16 && !flip)
    return device_ptrs ? setFrameDevice(img, depth, mask_p, inputs_follow_stream) : uploadFrame(img, depth, mask_p);
  const size_t n = (size_t)W * H;
  for (int k = 0; k < 2; ++k) {  // raw buffers of the ingest path, on first use
    if (depth16 && !d16Buf[k]) {
      RET_IF(cudaMalloc((void**)&d16Buf[k], n * 2));
      RET_IF(cudaMallocHost((void**)&h_d16Buf[k], n * 2));
    }
    if (flip && !rawImgBuf[k]) RET_IF(cudaMalloc((void**)&rawImgBuf[k], n * 3));
  }
  RET_IF(beginFrame(device_ptrs && inputs_follow_stream));
  cudaStream_t cs = device_ptrs ? preStream : copyStream;
  const cudaMemcpyKind kind = device_ptrs ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  cudaPointerAttributes a;
  auto pinned = [&](const void* q) { return cudaPointerGetAttributes(&a, q) == cudaSuccess && a.type == cudaMemoryTypeHost; };
  const void *src_img = img, *src_d = depth16 ? (const void*)depth16 : (const void*)depth;
  if (!device_ptrs) {
    const bool pi = pinned(img), pd = pinned(src_d);
    if (!pi || !pd) {
      cudaGetLastError();
      RET_IF(cudaEventSynchronize(evCopied[cur]));  // the staging buffers' previous transfer is done
    }
    if (!pi) {
      memcpy(h_rgbBuf[cur], img, n * 3);
      src_img = h_rgbBuf[cur];
    }
    if (!pd) {
      if (depth16) {
        memcpy(h_d16Buf[cur], depth16, n * 2);
        src_d = h_d16Buf[cur];
      } else {
        memcpy(h_depthBuf[cur], depth, n * 4);
        src_d = h_depthBuf[cur];
      }
    }
  }
  RET_IF(cudaMemcpyAsync(flip ? rawImgBuf[cur] : rgb, src_img, n * 3, kind, cs));
  if (depth16)
    RET_IF(cudaMemcpyAsync(d16Buf[cur], src_d, n * 2, kind, cs));
  else
    RET_IF(cudaMemcpyAsync(depthRaw, src_d, n * 4, kind, cs));
  if (!device_ptrs) {
    RET_IF(cudaEventRecord(evCopied[cur], copyStream));
    RET_IF(cudaStreamWaitEvent(preStream, evCopied[cur], 0));
  }
  RET_IF(launch_ingest(flip ? rawImgBuf[cur] : nullptr, depth16 ? d16Buf[cur] : nullptr, scale, flip ? 1 : 0, rgb, depthRaw,
                       (int)n, preStream));
  launches += 1;
  RET_IF(cudaEventRecord(evInputs[cur], preStream));
  RET_IF(cudaStreamWaitEvent(stream, evInputs[cur], 0));
  if (mask_p) {
    if (device_ptrs) {
      RET_IF(cudaMemcpyAsync(mask, mask_p, n, cudaMemcpyDeviceToDevice, stream));
      maskIsZero = false;
    } else {
      const uint8_t* m = mask_p;
      if (!pinned(mask_p)) {
        cudaGetLastError();
        RET_IF(cudaStreamSynchronize(stream));
        memcpy(h_mask, mask_p, n);
        m = h_mask;
      }
      RET_IF(cudaMemcpyAsync(mask, m, n, cudaMemcpyHostToDevice, stream));
      maskIsZero = false;
    }
  } else if (!keepMask) {
    if (!maskIsZero) RET_IF(cudaMemsetAsync(mask, 0, n, stream));  // (still zero from the last frame otherwise)
    maskIsZero = true;
  }
  return cudaSuccess;
}

cudaError_t Context::setFrameDevice(const uint8_t* rgb_d, const float* depth_d, const uint8_t* mask_d, bool inputs_follow_stream) {
  const size_t n = (size_t)W * H;
  RET_IF(beginFrame(inputs_follow_stream));
  RET_IF(cudaMemcpyAsync(rgb, rgb_d, n * 3, cudaMemcpyDeviceToDevice, preStream));
  RET_IF(cudaMemcpyAsync(depthRaw, depth_d, n * 4, cudaMemcpyDeviceToDevice, preStream));
  RET_IF(cudaEventRecord(evInputs[cur], preStream));
  RET_IF(cudaStreamWaitEvent(stream, evInputs[cur], 0));
  if (mask_d) {
    RET_IF(cudaMemcpyAsync(mask, mask_d, n, cudaMemcpyDeviceToDevice, stream));
    maskIsZero = false;
  } else if (!keepMask) {
    if (!maskIsZero) RET_IF(cudaMemsetAsync(mask, 0, n, stream));  // (still zero from the last frame otherwise)
    maskIsZero = true;
  }
  return cudaSuccess;
}

cudaError_t Context::preprocess(float depthCutoff) {
  RET_IF(launch_bilateral(depthRaw, (size_t)W * 4, W, H, depthCutoff, depthFiltered, (size_t)W * 4, preStream));
  {
    const void* src[1] = {depthPyr[0]};
    void* l1[1] = {depthPyr[1]};
    void* l2[1] = {depthPyr[2]};
    const int u8[1] = {0};
    RET_IF(launch_pyramid2(1, src, l1, l2, u8, W, H, preStream));
  }
  RET_IF(cudaEventRecord(evPre[cur], preStream));
  RET_IF(cudaStreamWaitEvent(stream, evPre[cur], 0));
  launches += 2;
  return cudaSuccess;
}

Model::Model(Context* c, unsigned id_, float conf, unsigned maxSurfels, bool enableFillIn)
    : ctx(c),
      id(id_),
      confidenceThreshold(conf),
      maxDepth(FLT_MAX),
      allowsFillIn(enableFillIn),
      odom(c->W, c->H, c->K.cx, c->K.cy, c->K.fx, c->K.fy),
      capacity(maxSurfels) {
  work = c->stream;
  for (int i = 0; i < 16; ++i) pose[i] = lastPose[i] = (i % 5 == 0) ? 1.f : 0.f;
  const size_t n = (size_t)c->W * c->H;
  bool good = dalloc(&predVertex, n * 4, &zeroed_) && dalloc(&predNormal, n * 4, &zeroed_) && dalloc(&predImage, n * 4, &zeroed_) &&
              dalloc(&icpError, n, &zeroed_);
  const size_t scanCap = (size_t)maxSurfels + n;
  good = good && dalloc(&buf[0], maxSurfels, &zeroed_) && dalloc(&buf[1], maxSurfels, &zeroed_) && dalloc(&unstable, n, &zeroed_) &&
         dalloc(&candStaging, n, &zeroed_) && dalloc(&candBest, n, &zeroed_) && dalloc(&winner, maxSurfels, &zeroed_) && dalloc(&keys, n, &zeroed_) &&
         dalloc(&indexMaps.index, n, &zeroed_) && dalloc(&indexMaps.vertConf, n, &zeroed_) && dalloc(&indexMaps.colorTime, n, &zeroed_) &&
         dalloc(&indexMaps.normRad, n, &zeroed_) && dalloc(&splat.image, n, &zeroed_) && dalloc(&splat.vertexConf, n, &zeroed_) &&
         dalloc(&splat.normalRad, n, &zeroed_) && dalloc(&splat.time, n, &zeroed_) && dalloc(&fill.image, n, &zeroed_) && dalloc(&fill.vertex, n, &zeroed_) &&
         dalloc(&fill.normal, n, &zeroed_) && dalloc(&scan.flags, scanCap, &zeroed_) && dalloc(&scan.ranks, scanCap, &zeroed_) &&
         dalloc(&scan.blockSums, 2 * (scanCap / 2048 + 4), &zeroed_) && dalloc(&counters, 1, &zeroed_);
  scan.capacity = scanCap;
  scan.host = new ScanHostState();
  // the projection passes take atomic minima into `keys`; the resolve passes leave it all ones again
  good = good && cudaMemset(keys, 0xFF, n * sizeof(unsigned long long)) == cudaSuccess;
  // likewise `winner` (atomic minima of candidate ordinals): the winners of a fuse restore their entries
  good = good && cudaMemset(winner, 0xFF, (size_t)maxSurfels * sizeof(uint32_t)) == cudaSuccess;
  good = good && cudaMallocHost(&h_counters, sizeof(MapCounters)) == cudaSuccess;
  if (good) memset(h_counters, 0, sizeof(MapCounters));
  // the pose block is followed by the tracker statistics of the frame: one read-back copy fetches both
  good = good && dalloc((PoseReadback**)&dpose, 1, &zeroed_) && cudaMallocHost(&h_readback, sizeof(PoseReadback)) == cudaSuccess &&
         cudaEventCreateWithFlags(&evPose, cudaEventDisableTiming) == cudaSuccess;
  // the per-model stream of multi-model frames exists from the start (its creation inside a frame was measured at
  // up to tens of milliseconds the first time)
  good = good && cudaStreamCreateWithFlags(&mstream, cudaStreamNonBlocking) == cudaSuccess &&
         cudaEventCreateWithFlags(&evJoin, cudaEventDisableTiming) == cudaSuccess;
  ok_ = good;
  if (good) ok_ = uploadPose() == cudaSuccess;
}

Model::~Model() {
  cudaFree(predVertex);
  cudaFree(predNormal);
  cudaFree(predImage);
  cudaFree(icpError);
  cudaFree(buf[0]);
  cudaFree(buf[1]);
  cudaFree(unstable);
  cudaFree(candStaging);
  cudaFree(candBest);
  cudaFree(winner);
  cudaFree(keys);
  cudaFree(indexMaps.index);
  cudaFree(indexMaps.vertConf);
  cudaFree(indexMaps.colorTime);
  cudaFree(indexMaps.normRad);
  cudaFree(splat.image);
  cudaFree(splat.vertexConf);
  cudaFree(splat.normalRad);
  cudaFree(splat.time);
  cudaFree(fill.image);
  cudaFree(fill.vertex);
  cudaFree(fill.normal);
  cudaFree(scan.flags);
  cudaFree(scan.ranks);
  cudaFree(scan.blockSums);
  delete scan.host;
  cudaFree(counters);
  cudaFreeHost(h_counters);
  if (mstream) cudaStreamDestroy(mstream);
  if (evJoin) cudaEventDestroy(evJoin);
  cudaFree(dpose);
  cudaFree(poseLogDev);
  cudaFreeHost(h_readback);
  if (evPose) cudaEventDestroy(evPose);
}

// A pooled model starts over under a new id (CoFusion::spawnObjectModel constructs one, CoFusion.cpp:588-597: some
// 75 allocations and 400 MB of page mapping per spawn): every buffer the constructor zeroed is zeroed again on the
// stream, the host state goes back to its initial values.  Requires the model's earlier work to be enqueued on
// (or joined into) the context's stream.
cudaError_t Model::recycle(unsigned id_, float conf) {
  id = id_;
  confidenceThreshold = conf;
  maxDepth = FLT_MAX;
  usePrediction = false;
  target = 0;
  renderSource = 1;
  count_ub = 0;
  cleanTick = 0;
  poseStale = false;
  for (int i = 0; i < 16; ++i) pose[i] = lastPose[i] = (i % 5 == 0) ? 1.f : 0.f;
  poseLogTs.clear();
  poseLogFrame.clear();
  poseLogHost.clear();
  RET_IF(cudaStreamSynchronize(work));  // h_counters / h_readback are targets of asynchronous copies
  memset(h_counters, 0, sizeof(MapCounters));
  for (auto& z : zeroed_) RET_IF(cudaMemsetAsync(z.first, 0, z.second, work));
  *scan.host = ScanHostState();  // the ticket counter and the status words are zero again
  RET_IF(cudaMemsetAsync(keys, 0xFF, (size_t)ctx->W * ctx->H * sizeof(unsigned long long), work));
  RET_IF(cudaMemsetAsync(winner, 0xFF, (size_t)capacity * sizeof(uint32_t), work));
  RET_IF(odom.recycle(work));
  return uploadPose();
}

// What CoFusion::inactivateModel keeps of a lost model (CoFusion.cpp:620-626): its map (the live surfels, in a
// buffer of their size), pose, threshold and pose log.  Synchronises.
cudaError_t Model::archive(ArchivedModel* out) {
  RET_IF(syncPose());
  RET_IF(fetchPoseLog());
  out->id = id;
  out->confidenceThreshold = confidenceThreshold;
  memcpy(out->pose, pose, sizeof(pose));
  out->poseLogTs = poseLogTs;
  out->poseLogFrame = poseLogFrame;
  out->poseLogHost = poseLogHost;
  unsigned n = 0;
  RET_IF(lastCount(&n));
  out->count = n;
  if (n) {
    RET_IF(cudaMalloc((void**)&out->surfels, (size_t)n * sizeof(Surfel)));
    RET_IF(cudaMemcpyAsync(out->surfels, buf[target], (size_t)n * sizeof(Surfel), cudaMemcpyDeviceToDevice, work));
    RET_IF(cudaStreamSynchronize(work));
  }
  return cudaSuccess;
}
ArchivedModel::~ArchivedModel() { cudaFree(surfels); }

// Run this model's next calls on its own stream, ordered after `after` (an event on the context's stream); join()
// makes the context's stream wait for them.  The per-model stages of a frame are independent
// (`for (auto model : models)`, CoFusion.cpp:465-488, :536-542): with several models their small kernels overlap.
cudaError_t Model::fork(cudaEvent_t after) {
  if (!mstream) {
    RET_IF(cudaStreamCreateWithFlags(&mstream, cudaStreamNonBlocking));
    RET_IF(cudaEventCreateWithFlags(&evJoin, cudaEventDisableTiming));
  }
  RET_IF(cudaStreamWaitEvent(mstream, after, 0));
  work = mstream;
  return cudaSuccess;
}
cudaError_t Model::join() {
  if (work == ctx->stream) return cudaSuccess;
  RET_IF(cudaEventRecord(evJoin, mstream));
  RET_IF(cudaStreamWaitEvent(ctx->stream, evJoin, 0));
  work = ctx->stream;
  return cudaSuccess;
}

cudaError_t Model::appendPoseLog(int64_t ts, int frame) {
  const int n = (int)poseLogTs.size(), fetched = (int)(poseLogHost.size() / 12);
  if (n - fetched >= poseLogCap) {
    if (poseLogCap) RET_IF(fetchPoseLog());  // device chunk full: move it to the host, reuse it
    if (!poseLogDev) {
      poseLogCap = 4096;
      RET_IF(cudaMalloc((void**)&poseLogDev, (size_t)poseLogCap * 12 * sizeof(float)));
    }
  }
  const int slot = (int)poseLogTs.size() - (int)(poseLogHost.size() / 12);
  RET_IF(cudaMemcpyAsync(poseLogDev + (size_t)slot * 12, dpose->pose.m, 12 * sizeof(float), cudaMemcpyDeviceToDevice, work));
  poseLogTs.push_back(ts);
  poseLogFrame.push_back(frame);
  return cudaSuccess;
}

cudaError_t Model::fetchPoseLog() {
  const int fetched = (int)(poseLogHost.size() / 12), pending = (int)poseLogTs.size() - fetched;
  if (pending <= 0) return cudaSuccess;
  poseLogHost.resize((size_t)(fetched + pending) * 12);
  RET_IF(cudaMemcpyAsync(poseLogHost.data() + (size_t)fetched * 12, poseLogDev, (size_t)pending * 12 * sizeof(float),
                         cudaMemcpyDeviceToHost, work));
  return cudaStreamSynchronize(work);
}

cudaError_t Model::uploadPose() {
  PoseDev b;
  memset(&b, 0, sizeof(b));
  float inv[16];
  pose_inverse16(pose, inv);
  for (int i = 0; i < 12; ++i) {
    b.pose.m[i] = pose[i];
    b.inv.m[i] = inv[i];
    b.last.m[i] = lastPose[i];
  }
  for (int r = 0; r < 3; ++r) {
    b.tr[r] = pose[r * 4 + 3];
    for (int c = 0; c < 3; ++c) b.tr[3 + r * 3 + c] = pose[r * 4 + c];
  }
  b.weightBase = fusion_weight_base(pose, lastPose);
  poseStale = false;
  // pageable source: the runtime stages the 224 bytes before returning, `b` may go out of scope
  return cudaMemcpyAsync(dpose, &b, sizeof(b), cudaMemcpyHostToDevice, work);
}

cudaError_t Model::enqueuePoseReadback() {
  RET_IF(cudaMemcpyAsync(h_readback, dpose, sizeof(PoseReadback), cudaMemcpyDeviceToHost, work));  // block + statistics
  RET_IF(cudaEventRecord(evPose, work));
  poseStale = true;
  return cudaSuccess;
}

cudaError_t Model::syncPose() {
  if (!poseStale) return cudaSuccess;
  RET_IF(cudaEventSynchronize(evPose));
  const PoseDev& b = h_readback->block;
  for (int i = 0; i < 12; ++i) {
    pose[i] = b.pose.m[i];
    lastPose[i] = b.last.m[i];
  }
  odom.setStats(h_readback->stats);
  poseStale = false;
  return cudaSuccess;
}

namespace {
__global__ void rgb_to_rgba_kernel(const uint8_t* __restrict__ src, uchar4* __restrict__ dst, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = make_uchar4(src[3 * i], src[3 * i + 1], src[3 * i + 2], 255);
}
}  // namespace

cudaError_t Model::setPrediction(const float* v4, const float* n4, const uint8_t* img, int channels, bool dev) {
  const size_t n = (size_t)ctx->W * ctx->H;
  cudaMemcpyKind k = dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  RET_IF(cudaMemcpyAsync(predVertex, v4, n * 16, k, work));
  RET_IF(cudaMemcpyAsync(predNormal, n4, n * 16, k, work));
  if (channels == 4) {
    RET_IF(cudaMemcpyAsync(predImage, img, n * 4, k, work));
  } else {
    // RGB8 -> RGBA8: staged through the ICP error buffer (4n bytes, rewritten by the next track)
    RET_IF(cudaMemcpyAsync(icpError, img, n * 3, k, work));
    rgb_to_rgba_kernel<<<(unsigned)((n + 255) / 256), 256, 0, work>>>((const uint8_t*)icpError,
                                                                            (uchar4*)predImage, (int)n);
    RET_IF(cudaGetLastError());
  }
  return cudaSuccess;
}

cudaError_t Model::initFirstRGB() {
  return odom.initFirstRGB(ctx->rgb, (size_t)ctx->W * 3, 3, work);
}

namespace {
void pose_inverse(const float* T, float* Ti) {  // Eigen inverse of a rigid 4x4, f32
  memset(Ti, 0, 64);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) Ti[r * 4 + c] = T[c * 4 + r];
  for (int r = 0; r < 3; ++r) Ti[r * 4 + 3] = -(Ti[r * 4 + 0] * T[3] + Ti[r * 4 + 1] * T[7] + Ti[r * 4 + 2] * T[11]);
  Ti[15] = 1;
}
Pose34 to34(const float* T) {
  Pose34 p;
  memcpy(p.m, T, sizeof(p.m));
  return p;
}
}  // namespace

float Model::computeFusionWeight(float weightMultiplier) const {
  // Model.cpp:391-406; the arithmetic lives in pose_math.cuh so that the tracker's epilogue computes the
  // same bits on the device
  return fusion_weight_base(pose, lastPose) * weightMultiplier;
}

cudaError_t Model::initialise(int time, float maxDepthProcessed) {
  RET_IF(launch_surfel_initialise(geom(), ctx->rgb, ctx->depthRaw, ctx->depthFiltered, time, maxDepthProcessed,
                                  buf[target], capacity, candStaging, unstable, scan, counters, work));
  const unsigned n = (unsigned)ctx->W * ctx->H;
  count_ub = n < capacity ? n : capacity;
  ctx->launches += 8;
  return cudaSuccess;
}

// Tighten the host-side upper bound of the live surfel count from the counters that the last completed
// clean() copied back (asynchronously, into pinned memory): count(now) <= count(tick k) + (cleans enqueued
// since k) x candidates per frame.  One extra frame of margin covers a torn read of the 20-byte record.
void Model::refreshCountBound(int /*time*/) {
  const volatile MapCounters* hc = h_counters;
  const unsigned k = hc->cleanTick, c = hc->count;
  if (!k || (int)k > cleanTick) return;
  const unsigned cand_ub = (unsigned)((ctx->W + 1) / 2) * ((ctx->H + 1) / 2);
  const unsigned long long b = (unsigned long long)c + (unsigned long long)(cleanTick - (int)k + 1) * cand_ub;
  if (b < count_ub) count_ub = (unsigned)b;
}

cudaError_t Model::predictIndices(int time, float depthCutoff, int timeDelta) {
  refreshCountBound(time);
  ctx->launches += 2;
  return launch_predict_indices(geom(), buf[target], count_ub, counters, invRef(), time, depthCutoff, timeDelta, keys,
                                indexMaps, work);
}

cudaError_t Model::fuse(int time, float depthCutoff, float weightMultiplier) {
  const float md = depthCutoff < maxDepth ? depthCutoff : maxDepth;  // Model.cpp:443
  ctx->launches += 3;  // associate, scan, apply
  return launch_fuse(geom(), buf[target], count_ub, counters, poseRef(), time, ctx->rgb, ctx->mask, ctx->depthRaw,
                     ctx->depthFiltered, md, WeightRef(&dpose->weightBase, weightMultiplier), id, indexMaps, winner,
                     candStaging, candBest, unstable, scan, work);
}

cudaError_t Model::clean(int time, int timeDelta, float /*depthCutoff*/, float outlierCoefficient) {
  const unsigned cand_ub = (unsigned)((ctx->W + 1) / 2) * ((ctx->H + 1) / 2);
  RET_IF(launch_clean(geom(), buf[target], unstable, buf[renderSource], count_ub, cand_ub, capacity, counters,
                      invRef(), time, confidenceThreshold, timeDelta, ctx->depthFiltered, ctx->mask, id,
                      outlierCoefficient, indexMaps, scan, work));
  cleanTick = time;
  int t = target;
  target = renderSource;
  renderSource = t;
  unsigned ub = count_ub + cand_ub;
  count_ub = ub < capacity ? ub : capacity;
  ctx->launches += 3;  // evaluate, scan, scatter (+ closing of the pass)
  // refresh the host-side bound with the exact count whenever the stream is next synchronised
  RET_IF(cudaMemcpyAsync(h_counters, counters, sizeof(MapCounters), cudaMemcpyDeviceToHost, work));
  return cudaSuccess;
}

cudaError_t Model::combinedPredict(float depthCutoff, int time, int maxTime, int timeDelta) {
  usePrediction = true;
  ctx->launches += 2;
  return launch_combined_predict(geom(), buf[target], count_ub, counters, invRef(), depthCutoff, confidenceThreshold, time,
                                 maxTime, timeDelta, keys, splat, work);
}

cudaError_t Model::performFillIn(bool frameToFrameRGB, bool lost) {
  if (!allowsFillIn) return cudaSuccess;
  ctx->launches += 1;
  return launch_fill_in(geom(), splat, ctx->rgb, ctx->depthFiltered, lost ? 1 : 0, (lost || frameToFrameRGB) ? 1 : 0,
                        fill, counters, 0.75f, work);
}

cudaError_t Model::downloadMap(float* dst, size_t cap, unsigned* count_out) {
  RET_IF(cudaMemcpyAsync(h_counters, counters, sizeof(MapCounters), cudaMemcpyDeviceToHost, work));
  RET_IF(cudaStreamSynchronize(work));
  unsigned n = h_counters->count;
  count_ub = n;
  if (count_out) *count_out = n;
  if (dst && n) {
    if (n > cap) n = (unsigned)cap;
    RET_IF(cudaMemcpyAsync(dst, buf[target], (size_t)n * sizeof(Surfel), cudaMemcpyDeviceToHost, work));
    RET_IF(cudaStreamSynchronize(work));
  }
  return cudaSuccess;
}

cudaError_t Model::uploadMap(const float* src, unsigned count) {
  if (count > capacity) count = capacity;
  RET_IF(cudaMemcpyAsync(buf[target], src, (size_t)count * sizeof(Surfel), cudaMemcpyHostToDevice, work));
  MapCounters c = {count, 0, 0, 0, 0, 0, 0, 0};
  *h_counters = c;
  RET_IF(cudaMemcpyAsync(counters, h_counters, sizeof(MapCounters), cudaMemcpyHostToDevice, work));
  RET_IF(cudaStreamSynchronize(work));
  count_ub = count;
  return cudaSuccess;
}

cudaError_t Model::lastCount(unsigned* out) { return downloadMap(nullptr, 0, out); }

cudaError_t Model::prepareTracking(const TrackParams& tp, bool devicePose) {
  if (!devicePose) {  // host-driven step: the host copies must be current, lastPose <- pose
    RET_IF(syncPose());
    memcpy(lastPose, pose, sizeof(pose));
  }  // otherwise the tracker's epilogue moves pose -> last inside the device block
  cudaStream_t s = work;
  const float *pv = predVertex, *pn = predNormal;
  const uint8_t* pi = predImage;
  PredAlt alt;
  if (usePrediction) {
    // Model::initICP (Model.cpp:350-367): the tracker reads the splat prediction as it is (no copy) -- or, for the
    // camera model, the fill-in images when CoFusion::requiresFillIn says so: the pyramid builders read through that
    // choice (a device flag, no host wait, no selection pass)
    pv = (const float*)splat.vertexConf;
    pn = (const float*)splat.normalRad;
    pi = (const uint8_t*)splat.image;
    if (allowsFillIn) {
      alt.v4 = (const float*)fill.vertex;
      alt.n4 = (const float*)fill.normal;
      alt.img = (const unsigned char*)fill.image;
      alt.sel = &counters->fillInRequired;
      alt.img_always = tp.frameToFrameRGB ? 1 : 0;
    }
  }
  // Model::initICP (Model.cpp:350-367): model pyramids first, then the frame's (fused launches)
  const float* pyr[3] = {ctx->depthPyr[0], ctx->depthPyr[1], ctx->depthPyr[2]};
  RET_IF(odom.initAll(pv, pn, pi, 4, pyr, ctx->rgb, 3, tp.maxDepthProcessed, pose, s, devicePose ? dpose->pose.m : nullptr,
                      alt.sel ? &alt : nullptr));
  ctx->launches += 3;  // model pyramid, frame maps + grey images, all pyramids (lastDepth + both grey images, both levels)
  return cudaSuccess;
}

void Model::finishTracking(const float trans[3], const float rot[9]) {
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) pose[r * 4 + c] = rot[r * 3 + c];
    pose[r * 4 + 3] = trans[r];
  }
  // the tracker synchronised the stream: the counters copied after the last clean are exact now
  if (h_counters->count && h_counters->count < count_ub) count_ub = h_counters->count;
  uploadPose();  // the kernels of this frame read the pose from the device block
}

cudaError_t Model::performTracking(const TrackParams& tp) {
  RET_IF(prepareTracking(tp, false));
  float trans[3] = {pose[3], pose[7], pose[11]};
  float rot[9] = {pose[0], pose[1], pose[2], pose[4], pose[5], pose[6], pose[8], pose[9], pose[10]};
  RET_IF(odom.getIncrementalTransformation(trans, rot, tp.rgbOnly != 0, tp.icpWeight, tp.pyramid != 0,
                                           tp.fastOdom != 0, tp.so3 != 0, icpError, (size_t)ctx->W * 4,
                                           tp.force_host_loop != 0, work));
  finishTracking(trans, rot);
  ctx->launches += 2;  // prepare, persistent GN
  return cudaSuccess;
}

cudaError_t trackModels(Context* ctx, Model* const* models, int n, const TrackParams& tp, bool async) {
  const bool icp = !tp.rgbOnly && tp.icpWeight > 0, rgb = tp.rgbOnly || tp.icpWeight < 100;
  int i = 0;
  while (i < n) {
    int nb = n - i;
    if (nb > RGBDOdometry::kMaxBatch) nb = RGBDOdometry::kMaxBatch;
    const bool batch = icp && rgb && !tp.force_host_loop && models[i]->odom.canBatch(nb);
    if (!batch) {
      RET_IF(models[i]->performTracking(tp));
      i += 1;
      continue;
    }
    if (!ctx->batchScratch) {
      RET_IF(cudaMalloc(&ctx->batchScratch, RGBDOdometry::tiledScratchBytes()));
      RET_IF(cudaMemsetAsync(ctx->batchScratch, 0, RGBDOdometry::tiledScratchBytes(), ctx->stream));
    }
    RGBDOdometry* od[RGBDOdometry::kMaxBatch];
    PoseDev* pd[RGBDOdometry::kMaxBatch];
    float trans[RGBDOdometry::kMaxBatch][3], rot[RGBDOdometry::kMaxBatch][9];
    float* err[RGBDOdometry::kMaxBatch];
    // several models: each one's model pyramids and Sobel / candidate pass run on its own stream
    const bool spread = async && nb > 1;
    if (spread) {
      RET_IF(cudaEventRecord(ctx->evFork, ctx->stream));
      for (int k = 0; k < nb; ++k) RET_IF(models[i + k]->fork(ctx->evFork));
    }
    for (int k = 0; k < nb; ++k) {
      Model* m = models[i + k];
      RET_IF(m->prepareTracking(tp, async));
      if (spread) RET_IF(m->odom.enqueuePrepare(m->work, k == 0 ? ctx->batchScratch : nullptr, nb, k > 0));
      od[k] = &m->odom;
      pd[k] = m->dpose;
      err[k] = m->icpError;
      const float* P = m->pose;
      const float t[3] = {P[3], P[7], P[11]};
      const float r[9] = {P[0], P[1], P[2], P[4], P[5], P[6], P[8], P[9], P[10]};
      memcpy(trans[k], t, sizeof(t));
      memcpy(rot[k], r, sizeof(r));
    }
    if (spread)
      for (int k = 0; k < nb; ++k) RET_IF(models[i + k]->join());
    RET_IF(RGBDOdometry::trackTiled(od, nb, trans, rot, tp.icpWeight, tp.pyramid != 0, tp.fastOdom != 0, tp.so3 != 0, err,
                                    (size_t)ctx->W * 4, ctx->batchScratch, ctx->stream, async ? pd : nullptr, async, spread));
    for (int k = 0; k < nb; ++k) {
      if (async)
        RET_IF(models[i + k]->enqueuePoseReadback());  // pose + stats reach the host when somebody asks (syncPose)
      else
        models[i + k]->finishTracking(trans[k], rot[k]);
    }
    ctx->launches += nb + 1;  // one prepare per model + ONE persistent GN launch
    i += nb;
  }
  return cudaSuccess;
}

}  // namespace cfb
