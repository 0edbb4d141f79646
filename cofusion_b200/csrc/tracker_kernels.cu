// tracker_kernels.cu -- the four per-pixel residual/Jacobian reductions of the projective ICP+RGB
// tracker, hand-written for sm_100a.
//
// What they compute (arithmetic spec; reference implementation in
// Core/Cuda/reduce.cu: ICPReduction :257-423, RGBResidual :748-891, RGBReduction :503-633,
// SO3Reduction :973-1116).  What is different by design:
//   * ONE launch per step.  The reference needs kernel + reduceSum<<<1,1024>>> + device sync + D2H;
//     here the last block to retire folds the per-block partials in a fixed order
//     (grid_finalize32) and leaves the packed sums in device memory for the device-resident
//     Gauss-Newton step.
//   * warp transpose reduction with real SHFL (the reference's "__shfl_down" is a shared-memory +
//     2 barrier emulation for __CUDA_ARCH__ > 700, reduce.cu:56-80): 31 SHFL per warp for all 29
//     sums instead of 290 block barriers.
//   * grid sized from the SM count (persistent, grid-stride), 4-pixel ILP per thread at 640x480.
#include "tracker_kernels.cuh"

#include "tracker_device.cuh"

namespace cfb {

namespace {
using namespace dev;

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
icp_step_kernel(const IcpArgs a, const IcpPose* __restrict__ d_pose, StepScratch* __restrict__ sc) {
  __shared__ IcpPose P;
  __shared__ float red[(kThreads / 32) * 32];
  __shared__ float out32[32];
  for (int i = threadIdx.x; i < (int)(sizeof(IcpPose) / 4); i += blockDim.x)
    ((float*)&P)[i] = ((const float*)d_pose)[i];
  __syncthreads();

  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;

  const int N = a.cols * a.rows;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += blockDim.x * gridDim.x) {
    int y = i / a.cols;
    int x = i - y * a.cols;
    icp_pixel(a, P, x, y, acc);
  }
  float bt = block_reduce32(acc, red);
  if (grid_finalize32(bt, sc->partials, &sc->ticket, red, out32)) {
    if (threadIdx.x < 32) sc->result[threadIdx.x] = out32[threadIdx.x];
  }
}

__global__ void __launch_bounds__(kThreads)
rgb_residual_kernel(const RgbResidualArgs a, const RgbWarp* __restrict__ d_warp,
                    StepScratch* __restrict__ sc) {
  __shared__ RgbWarp Wp;
  for (int i = threadIdx.x; i < (int)(sizeof(RgbWarp) / 4); i += blockDim.x)
    ((float*)&Wp)[i] = ((const float*)d_warp)[i];
  __syncthreads();
  int cnt = 0, sig = 0;
  const int N = a.cols * a.rows;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += blockDim.x * gridDim.x) {
    int i = k / a.cols;
    int j0 = k - i * a.cols;
    DataTerm c;
    int sq;
    if (rgb_residual_pixel(a, Wp, j0, i, c, sq)) {
      cnt += 1;
      sig += sq;  // wrapping int32, as the reference
    }
    // one 16-byte store
    int4 raw;
    raw.x = (int)((unsigned short)c.zero.x | ((unsigned)(unsigned short)c.zero.y << 16));
    raw.y = (int)((unsigned short)c.one.x | ((unsigned)(unsigned short)c.one.y << 16));
    raw.z = __float_as_int(c.diff);
    raw.w = c.valid ? 1 : 0;
    reinterpret_cast<int4*>(a.corres)[k] = raw;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    sig += __shfl_xor_sync(0xffffffffu, sig, o);
  }
  __shared__ int scnt[kThreads / 32], ssig[kThreads / 32];
  if ((threadIdx.x & 31) == 0) {
    scnt[threadIdx.x >> 5] = cnt;
    ssig[threadIdx.x >> 5] = sig;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int c = 0, s = 0;
    for (int w = 0; w < kThreads / 32; ++w) {
      c += scnt[w];
      s += ssig[w];
    }
    // integer adds commute exactly: atomics keep the result deterministic
    atomicAdd(&sc->rgb_count, c);
    atomicAdd(&sc->rgb_sigma, s);
  }
}

__global__ void __launch_bounds__(kThreads)
rgb_step_kernel(const RgbStepArgs a, float sigma_arg, StepScratch* __restrict__ sc) {
  __shared__ float red[(kThreads / 32) * 32];
  __shared__ float out32[32];
  float sigma = sigma_arg;
  if (sigma_arg < -1.5f) {
    // RGBDOdometry.cpp:373-374: tmpError = sqrt(sigma)/count ; sigmaVal = (tmpError==0) ? 1 : count
    int cnt = sc->rgb_count, sg = sc->rgb_sigma;
    float tmpError = (float)(sqrt((double)sg) / (double)cnt);
    sigma = (tmpError == 0.f) ? 1.f : (float)cnt;
  }
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
  const int N = a.cols * a.rows;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += blockDim.x * gridDim.x) {
    int4 raw = __ldg(reinterpret_cast<const int4*>(a.corres) + k);
    DataTerm c;
    c.zero = make_short2((short)(raw.x & 0xffff), (short)((unsigned)raw.x >> 16));
    c.one = make_short2((short)(raw.y & 0xffff), (short)((unsigned)raw.y >> 16));
    c.diff = __int_as_float(raw.z);
    c.valid = (raw.w & 0xff) != 0;
    rgb_step_pixel(a, sigma, c, acc);
  }
  float bt = block_reduce32(acc, red);
  if (grid_finalize32(bt, sc->partials, &sc->ticket, red, out32)) {
    if (threadIdx.x < 32) sc->result[threadIdx.x] = out32[threadIdx.x];
  }
}

__global__ void __launch_bounds__(kThreads)
so3_step_kernel(const So3Args a, StepScratch* __restrict__ sc) {
  __shared__ float red[(kThreads / 32) * 32];
  __shared__ float out32[32];
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
  const int N = a.cols * a.rows, cols = a.cols, rows = a.rows;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += blockDim.x * gridDim.x) {
    int y = k / cols;
    int x = k - y * cols;
    so3_pixel(a.lastImage, a.nextImage, a.img_pitch, cols, rows, a.imageBasis, a.kinv, a.krlr, x, y, acc);
  }
  float bt = block_reduce32(acc, red);
  if (grid_finalize32(bt, sc->partials, &sc->ticket, red, out32)) {
    if (threadIdx.x < 32) sc->result[threadIdx.x] = out32[threadIdx.x];
  }
}

int grid_for(int N, int per_thread) {
  int want = (N + kThreads * per_thread - 1) / (kThreads * per_thread);
  int cap = num_sms() * 4;
  if (cap > kMaxBlocks) cap = kMaxBlocks;
  if (want < 1) want = 1;
  return want < cap ? want : cap;
}

}  // namespace

int num_sms() {  // of the CURRENT device (the ABI entries run under the device of their handle)
  static int cache[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!cache[dev]) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = n > 0 ? n : 148;
  }
  return cache[dev];
}

cudaError_t launch_icp_step(const IcpArgs& a, const IcpPose* d_pose, StepScratch* sc, cudaStream_t s) {
  icp_step_kernel<<<grid_for(a.cols * a.rows, 2), kThreads, 0, s>>>(a, d_pose, sc);
  return cudaGetLastError();
}
cudaError_t launch_rgb_residual(const RgbResidualArgs& a, const RgbWarp* d_warp, StepScratch* sc,
                                cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(&sc->rgb_count, 0, 2 * sizeof(int), s);
  if (e != cudaSuccess) return e;
  rgb_residual_kernel<<<grid_for(a.cols * a.rows, 2), kThreads, 0, s>>>(a, d_warp, sc);
  return cudaGetLastError();
}
cudaError_t launch_rgb_step(const RgbStepArgs& a, float sigma, StepScratch* sc, cudaStream_t s) {
  rgb_step_kernel<<<grid_for(a.cols * a.rows, 2), kThreads, 0, s>>>(a, sigma, sc);
  return cudaGetLastError();
}
cudaError_t launch_so3_step(const So3Args& a, StepScratch* sc, cudaStream_t s) {
  so3_step_kernel<<<grid_for(a.cols * a.rows, 2), kThreads, 0, s>>>(a, sc);
  return cudaGetLastError();
}

}  // namespace cfb
