// gn_device.cu -- device-resident Gauss-Newton loop of the tracker (RGBDOdometry::deviceLoop).
//
// The reference crosses the host<->device boundary >= 60 times per model per frame: each of the 19
// GN iterations is 3 x (kernel, reduceSum<<<1,1024>>>, cudaDeviceSynchronize, 116-byte D2H) plus a
// host 6x6 LDLT (RGBDOdometry.cpp:347-461, reduce.cu:474-482).  At 640x480 the per-iteration data
// (<= 34 MB, L2 resident on B200) moves in a few microseconds, so those crossings ARE the frame
// time.  Here the whole sequence is enqueued once:
//     gn_init -> so3_iter x10 -> gn_begin -> [pass1, pass2] x (4+5+10)
//   pass1 = RGB residual + ICP reduction fused (both only depend on the current pose),
//   pass2 = RGB Jacobian reduction; its finalising block (the last one to retire) runs the FP64
//           combine / LDLT / SE(3) update / next-warp computation of gn_math.h in one thread,
// and the host reads back 12 pose floats + stats once per frame.  SO(3) convergence tests
// (RGBDOdometry.cpp:285-292) become a device flag that turns the remaining so3 launches into no-ops.
#include <float.h>

#include "gn_serial.cuh"
#include "image_kernels.cuh"

namespace cfb {
namespace {
using namespace dev;
constexpr int kThreads = 256;

#define RET_IF(e)                       \
  do {                                  \
    cudaError_t e__ = (e);              \
    if (e__ != cudaSuccess) return e__; \
  } while (0)

// ---- gn_init / so3 / begin as tiny kernels (graph path) ------------------------------------------
__global__ void gn_init_kernel(GNState* g, StepScratch* sc, const float* __restrict__ pose_in, LevelK k_so3) {
  if (threadIdx.x == 0) gn_init_serial(g, sc, pose_in, k_so3);
}

__global__ void __launch_bounds__(kThreads)
so3_iter_kernel(const unsigned char* __restrict__ lastImage, const unsigned char* __restrict__ nextImage,
                size_t img_pitch, int cols, int rows, LevelK k, GNState* g, StepScratch* sc) {
  if (g->so3_done) return;  // uniform: set only by a previous launch
  __shared__ Mat33 M[3];
  __shared__ float red[(kThreads / 32) * 32];
  __shared__ float out32[32];
  for (int i = threadIdx.x; i < 27; i += blockDim.x) ((float*)M)[i] = ((const float*)&g->so3_imageBasis)[i];
  __syncthreads();
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
  const int N = cols * rows;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < N; p += blockDim.x * gridDim.x) {
    int y = p / cols, x = p - y * cols;
    so3_pixel(lastImage, nextImage, img_pitch, cols, rows, M[0], M[2], M[1], x, y, acc);
  }
  float bt = block_reduce32(acc, red);
  if (!grid_finalize32(bt, sc->partials, &sc->ticket, red, out32)) return;
  if (threadIdx.x == 0) so3_update_serial(g, out32, k);
}

__global__ void gn_begin_kernel(GNState* g, int use_so3, LevelK k_first) {
  if (threadIdx.x == 0) gn_begin_serial(g, use_so3, k_first);
}

__global__ void rgb_prepare_kernel(const unsigned char* __restrict__ img, int W, int H,
                                   const float* __restrict__ nextDepth, float minScale, short* __restrict__ dx,
                                   short* __restrict__ dy, unsigned char* __restrict__ cand) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  rgb_prepare_pixel(img, W, H, nextDepth, minScale, dx, dy, cand, x, y);
}

// ---- pass 1: RGB residual + ICP reduction, fused ------------------------------------------------
__global__ void __launch_bounds__(kThreads)
gn_pass1_kernel(const IcpArgs ia, const RgbResidualArgs ra, const unsigned char* __restrict__ cand, GNState* g,
                StepScratch* sc) {
  __shared__ IcpPose P;
  __shared__ RgbWarp Wp;
  __shared__ float red[(kThreads / 32) * 32];
  __shared__ float out32[32];
  __shared__ int scnt[kThreads / 32], ssig[kThreads / 32];
  for (int i = threadIdx.x; i < (int)(sizeof(IcpPose) / 4); i += blockDim.x) ((float*)&P)[i] = ((const float*)&g->pose)[i];
  for (int i = threadIdx.x; i < (int)(sizeof(RgbWarp) / 4); i += blockDim.x) ((float*)&Wp)[i] = ((const float*)&g->warp)[i];
  __syncthreads();
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
  int cnt = 0, sig = 0;
  const int N = ia.cols * ia.rows;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < N; p += blockDim.x * gridDim.x) {
    int y = p / ia.cols, x = p - y * ia.cols;
    icp_pixel(ia, P, x, y, acc);
    DataTerm c;
    c.valid = false;
    c.zero = make_short2(0, 0);
    c.one = make_short2(0, 0);
    c.diff = 0.f;
    int sq;
    if (__ldg(cand + p) && rgb_residual_cand(ra, Wp, x, y, c, sq)) {
      cnt += 1;
      sig += sq;
    }
    int4 raw;
    raw.x = (int)((unsigned short)c.zero.x | ((unsigned)(unsigned short)c.zero.y << 16));
    raw.y = (int)((unsigned short)c.one.x | ((unsigned)(unsigned short)c.one.y << 16));
    raw.z = __float_as_int(c.diff);
    raw.w = c.valid ? 1 : 0;
    reinterpret_cast<int4*>(ra.corres)[p] = raw;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    sig += __shfl_xor_sync(0xffffffffu, sig, o);
  }
  if ((threadIdx.x & 31) == 0) {
    scnt[threadIdx.x >> 5] = cnt;
    ssig[threadIdx.x >> 5] = sig;
  }
  float bt = block_reduce32(acc, red);  // contains a __syncthreads
  if (threadIdx.x == 0) {
    int c = 0, s = 0;
    for (int w = 0; w < kThreads / 32; ++w) {
      c += scnt[w];
      s += ssig[w];
    }
    atomicAdd(&sc->rgb_count, c);
    atomicAdd(&sc->rgb_sigma, s);
  }
  if (grid_finalize32(bt, sc->partials, &sc->ticket, red, out32)) {
    if (threadIdx.x < 32) g->icp_result[threadIdx.x] = out32[threadIdx.x];
  }
}

// ---- pass 2: RGB Jacobian reduction + FP64 GN step in the finalising block -----------------------
__global__ void __launch_bounds__(kThreads)
gn_pass2_kernel(const RgbStepArgs a, float icpWeight, LevelK k_next, int is_last, GNState* g, StepScratch* sc) {
  __shared__ float red[(kThreads / 32) * 32];
  __shared__ float out32[32];
  const int cnt = sc->rgb_count, sg = sc->rgb_sigma;
  float tmpError;
  const float sigma = rgb_sigma_from_counts(cnt, sg, &tmpError);
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
  const int N = a.cols * a.rows;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < N; p += blockDim.x * gridDim.x) {
    int4 raw = __ldcg(reinterpret_cast<const int4*>(a.corres) + p);
    DataTerm c;
    c.zero = make_short2((short)(raw.x & 0xffff), (short)((unsigned)raw.x >> 16));
    c.one = make_short2((short)(raw.y & 0xffff), (short)((unsigned)raw.y >> 16));
    c.diff = __int_as_float(raw.z);
    c.valid = (raw.w & 0xff) != 0;
    rgb_step_pixel(a, sigma, c, acc);
  }
  float bt = block_reduce32(acc, red);
  if (!grid_finalize32(bt, sc->partials, &sc->ticket, red, out32)) return;
  if (threadIdx.x == 0) gn_solve_serial(g, sc, out32, icpWeight, k_next, is_last, tmpError, cnt);
}

int grid_for(int N, int per_thread) {
  int want = (N + kThreads * per_thread - 1) / (kThreads * per_thread);
  int cap = num_sms() * 2;
  if (cap > kMaxBlocks) cap = kMaxBlocks;
  if (want < 1) want = 1;
  return want < cap ? want : cap;
}

}  // namespace

// Enqueue the whole tracker sequence on `s` (capturable: no host synchronisation inside).
cudaError_t RGBDOdometry::enqueueDeviceLoop(float icpWeight, bool pyramid, bool fastOdom, bool so3, float* err,
                                            size_t err_pitch, cudaStream_t s) {
  struct Out {
    float trans[3];
    float rot[9];
    TrackStats st;
  };
  static_assert(sizeof(Out) <= 2048, "staging");
  float* h_in = (float*)((char*)h_pinned + 1536);
  Out* ho = (Out*)((char*)h_pinned + 2048);
  RET_IF(cudaMemcpyAsync(d_pose_in, h_in, 12 * sizeof(float), cudaMemcpyHostToDevice, s));
  const dim3 b2(32, 8);
  for (int i = 0; i < NUM_PYRS; i++) {
    int w = width >> i, h = height >> i;
    float minScale = (float)(pow(minimumGradientMagnitudes[i], 2.0) / pow(sobelScale, 2.0));
    rgb_prepare_kernel<<<dim3((w + 31) / 32, (h + 7) / 8), b2, 0, s>>>(nextImage[i], w, h, (next_is_last_ ? lastDepth[i] : nextDepth[i]), minScale,
                                                                         nextdIdx[i], nextdIdy[i], rgbCand[i]);
    RET_IF(launch_project_to_point_cloud(lastDepth[i], (size_t)w * 4, w, h, intr.level(i), pointClouds[i],
                                         (size_t)w * 12, s));
  }
  auto LK = [&](int l) {
    Intr k = intr.level(l);
    return LevelK{k.fx, k.fy, k.cx, k.cy};
  };
  gn_init_kernel<<<1, 32, 0, s>>>(gn, scratch, d_pose_in, LK(2));
  if (so3) {
    const int L = 2, w = width >> L, h = height >> L;
    for (int it = 0; it < 10; ++it)
      so3_iter_kernel<<<grid_for(w * h, 1), kThreads, 0, s>>>(lastNextImage[L], nextImage[L], (size_t)w, w, h,
                                                              LK(L), gn, scratch);
  }
  int iterations[NUM_PYRS] = {fastOdom ? 3 : 10, pyramid ? 5 : 0, pyramid ? 4 : 0};
  int sched[32], n = 0;  // level of each GN iteration, coarse to fine
  for (int i = NUM_PYRS - 1; i >= 0; --i)
    for (int j = 0; j < iterations[i]; ++j) sched[n++] = i;
  gn_begin_kernel<<<1, 32, 0, s>>>(gn, so3 ? 1 : 0, LK(n ? sched[0] : 0));
  for (int q = 0; q < n; ++q) {
    const int i = sched[q];
    const int w = width >> i, h = height >> i;
    const Intr k = intr.level(i);
    const size_t p = (size_t)w * 4;
    IcpArgs ia;
    ia.vmap_curr = {vmaps_curr_[i], p};
    ia.nmap_curr = {nmaps_curr_[i], p};
    ia.vmap_g_prev = {vmaps_g_prev_[i], p};
    ia.nmap_g_prev = {nmaps_g_prev_[i], p};
    ia.intr = k;
    ia.distThres = distThres_;
    ia.angleThres = angleThres_;
    ia.cols = w;
    ia.rows = h;
    const bool last_of_l0 = (i == 0 && (q + 1 == n || sched[q + 1] != 0));
    ia.error_map = last_of_l0 ? err : nullptr;
    ia.error_pitch = err_pitch;
    RgbResidualArgs ra;
    ra.minScale = 0.f;  // folded into rgbCand
    ra.maxDepthDelta = maxDepthDeltaRGB;
    ra.dIdx = nextdIdx[i];
    ra.dIdy = nextdIdy[i];
    ra.grad_pitch = (size_t)w * 2;
    ra.lastDepth = lastDepth[i];
    ra.nextDepth = (next_is_last_ ? lastDepth[i] : nextDepth[i]);
    ra.depth_pitch = p;
    ra.lastImage = lastImage[i];
    ra.nextImage = nextImage[i];
    ra.img_pitch = (size_t)w;
    ra.corres = corresImg[i];
    ra.cols = w;
    ra.rows = h;
    RgbStepArgs sa;
    sa.corres = corresImg[i];
    sa.cloud = pointClouds[i];
    sa.cloud_pitch = (size_t)w * 12;
    sa.dIdx = nextdIdx[i];
    sa.dIdy = nextdIdy[i];
    sa.grad_pitch = (size_t)w * 2;
    sa.fx = k.fx;
    sa.fy = k.fy;
    sa.sobelScale = sobelScale;
    sa.cols = w;
    sa.rows = h;
    const int g1 = grid_for(w * h, 2);
    gn_pass1_kernel<<<g1, kThreads, 0, s>>>(ia, ra, rgbCand[i], gn, scratch);
    const bool is_last = (q + 1 == n);
    gn_pass2_kernel<<<g1, kThreads, 0, s>>>(sa, icpWeight, LK(is_last ? i : sched[q + 1]), is_last ? 1 : 0, gn,
                                            scratch);
  }
  RET_IF(cudaGetLastError());
  // one small D2H per frame: pose + stats
  RET_IF(cudaMemcpyAsync(ho->trans, gn->out_trans, 12 * sizeof(float), cudaMemcpyDeviceToHost, s));
  RET_IF(cudaMemcpyAsync(&ho->st, &gn->stats, sizeof(TrackStats), cudaMemcpyDeviceToHost, s));
  return cudaSuccess;
}

cudaError_t RGBDOdometry::deviceLoop(float trans[3], float rot[9], float icpWeight, bool pyramid, bool fastOdom,
                                     bool so3, float* err, size_t err_pitch, cudaStream_t s) {
  struct Out {
    float trans[3];
    float rot[9];
    TrackStats st;
  };
  float* h_in = (float*)((char*)h_pinned + 1536);
  Out* ho = (Out*)((char*)h_pinned + 2048);
  memcpy(h_in, trans, 3 * sizeof(float));
  memcpy(h_in + 3, rot, 9 * sizeof(float));

  // CUDA graph of the whole sequence (the ~60 launches are otherwise launch-latency bound).  The
  // legacy default stream cannot be captured: direct launches there.
  bool launched = false;
  if (mode_ == 0) {
    if (!tiled_scratch_) {
      RET_IF(cudaMalloc(&tiled_scratch_, tiledScratchBytes()));
      RET_IF(cudaMemsetAsync(tiled_scratch_, 0, tiledScratchBytes(), s));
    }
    RGBDOdometry* od[1] = {this};
    float* errs[1] = {err};
    return trackTiled(od, 1, (float(*)[3])trans, (float(*)[9])rot, icpWeight, pyramid, fastOdom, so3, errs, err_pitch,
                      tiled_scratch_, s);
  } else if (use_graphs_ && s != 0 && s != cudaStreamLegacy) {
    GraphKey key{parity_, err, err_pitch, icpWeight, pyramid, fastOdom, so3};
    cudaGraphExec_t exec = nullptr;
    for (auto& e : graphs_)
      if (e.key.parity == key.parity && e.key.err == key.err && e.key.err_pitch == key.err_pitch &&
          e.key.icpWeight == key.icpWeight && e.key.pyramid == key.pyramid && e.key.fastOdom == key.fastOdom &&
          e.key.so3 == key.so3)
        exec = e.exec;
    if (!exec) {
      cudaGraph_t graph = nullptr;
      RET_IF(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
      cudaError_t ce = enqueueDeviceLoop(icpWeight, pyramid, fastOdom, so3, err, err_pitch, s);
      cudaError_t ee = cudaStreamEndCapture(s, &graph);
      if (ce != cudaSuccess) return ce;
      RET_IF(ee);
      RET_IF(cudaGraphInstantiate(&exec, graph, 0));
      cudaGraphDestroy(graph);
      graphs_.push_back({key, exec});
    }
    RET_IF(cudaGraphLaunch(exec, s));
    launched = true;
  }
  if (!launched) RET_IF(enqueueDeviceLoop(icpWeight, pyramid, fastOdom, so3, err, err_pitch, s));
  RET_IF(cudaStreamSynchronize(s));
  if (time_kernel_) kernelTiming(nullptr, nullptr, false);
  memcpy(trans, ho->trans, sizeof(float) * 3);
  memcpy(rot, ho->rot, sizeof(float) * 9);
  stats_ = ho->st;
  if (so3) {
    for (int i = 0; i < NUM_PYRS; i++) {
      unsigned char* t = lastNextImage[i];
      lastNextImage[i] = nextImage[i];
      nextImage[i] = t;
    }
    parity_ ^= 1;
  }
  return cudaSuccess;
}

}  // namespace cfb
