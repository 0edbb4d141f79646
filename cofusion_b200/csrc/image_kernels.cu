// image_kernels.cu -- per-pixel image preparation for the tracker and the surfel stage (sm_100a).
// Compiled with -fmad=false: outputs that feed integer decisions downstream (bilateral depth ->
// surfel association / cleaning, grey -> photometric gates) are bit-identical to the CPU oracle.
//
// One wrapper per reference free function (Core/Cuda/cudafuncs.cuh:108-191) so each can be diffed
// 1:1 against the reference kernel; the arithmetic spec is SURVEY.md Appendix A6.  Differences by
// design: no per-call cudaMalloc/cudaFree of the 25 Gaussian taps (cudafuncs.cu:523-531, :581-587;
// taps live in __constant__), no device-wide sync after a launch (:434, :714, :750), every launch
// on the caller's stream, invalid vertices get NaN in all three planes (the reference writes the x
// plane only, :131).
#include "image_kernels.cuh"

#include "detmath.cuh"

namespace cfb {
namespace {

__constant__ float c_gauss25[25] = {1, 4, 6, 4, 1, 4, 16, 24, 16, 4, 6, 24, 36,
                                    24, 6, 4, 16, 24, 16, 4, 1, 4, 6, 4, 1};
// Core/Cuda/cudafuncs.cu:691-697
__constant__ float c_sobel_x[9] = {0.52201f, 0.00000f, -0.52201f, 0.79451f, -0.00000f,
                                   -0.79451f, 0.52201f, 0.00000f, -0.52201f};
__constant__ float c_sobel_y[9] = {0.52201f, 0.79451f, 0.52201f, 0.00000f, 0.00000f,
                                   0.00000f, -0.52201f, -0.79451f, -0.52201f};

inline dim3 grid2d(int w, int h, dim3 b) { return dim3((w + b.x - 1) / b.x, (h + b.y - 1) / b.y); }
const dim3 kBlock(32, 8);

// ---- a1: 13x13 bilateral on metric depth (depth_bilateral_metric.frag:30-76) -------------------
// 32x8 tile + 6-pixel halo staged in shared memory: each input texel is read from L2/HBM once per
// tile instead of 169 times.
constexpr int BR = 6;
__global__ void bilateral_kernel(const float* __restrict__ depth, size_t dpitch, int W, int H, float maxD,
                                 float* __restrict__ out, size_t opitch) {
  __shared__ float tile[8 + 2 * BR][32 + 2 * BR + 1];
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 8;
  for (int ty = threadIdx.y; ty < 8 + 2 * BR; ty += 8)
    for (int tx = threadIdx.x; tx < 32 + 2 * BR; tx += 32) {
      int gx = x0 + tx - BR, gy = y0 + ty - BR;
      float v = 0.f;
      if (gx >= 0 && gx < W && gy >= 0 && gy < H) v = __ldg(row_ptr(depth, dpitch, gy) + gx);
      tile[ty][tx] = v;
    }
  __syncthreads();
  const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
  if (x >= W || y >= H) return;
  const float value = tile[threadIdx.y + BR][threadIdx.x + BR];
  float res = 0.f;
  if (!(value > maxD || value < 0.3f)) {
    const float sigma_space2_inv_half = 0.024691358f;
    const float sigma_color2_inv_half = 555.556f;
    const int D = 2 * BR + 1;
    int tx = min(x - D / 2 + D, W), ty = min(y - D / 2 + D, H);
    float sum1 = 0.f, sum2 = 0.f;
    for (int cy = max(y - D / 2, 0); cy < ty; ++cy)
      for (int cx = max(x - D / 2, 0); cx < tx; ++cx) {
        float tmp = tile[cy - y0 + BR][cx - x0 + BR];
        float dx = (float)x - (float)cx, dy = (float)y - (float)cy;
        float space2 = dx * dx + dy * dy;
        float color2 = (value - tmp) * (value - tmp);
        float weight = det_expf(-(space2 * sigma_space2_inv_half + color2 * sigma_color2_inv_half));
        sum1 += tmp * weight;
        sum2 += weight;
      }
    res = sum1 / sum2;
  }
  row_ptr(out, opitch, y)[x] = res;
}

// ---- a2: 5x5 Gaussian 2x downsample of f32 depth skipping NaN (cudafuncs.cu:333-364) -----------
__global__ void pyr_down_gauss_f_kernel(const float* __restrict__ src, size_t spitch, int sw, int sh,
                                        float* __restrict__ dst, size_t dpitch, int dw, int dh) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= dw || y >= dh) return;
  const int D = 5;
  int tx = min(2 * x - D / 2 + D, sw - 1), ty = min(2 * y - D / 2 + D, sh - 1);
  float sum = 0.f;
  int count = 0;
  for (int cy = max(0, 2 * y - D / 2); cy < ty; ++cy) {
    const float* r = row_ptr(src, spitch, cy);
    for (int cx = max(0, 2 * x - D / 2); cx < tx; ++cx) {
      float s = __ldg(r + cx);
      if (!isnan(s)) {
        float w = c_gauss25[(ty - cy - 1) * 5 + (tx - cx - 1)];
        sum += s * w;
        count = (int)((float)count + w);  // `int += float`, :359
      }
    }
  }
  row_ptr(dst, dpitch, y)[x] = sum / (float)count;
}

// ---- a2/a5: same for u8, skipping zeros (cudafuncs.cu:534-564) --------------------------------
__global__ void pyr_down_uchar_kernel(const unsigned char* __restrict__ src, size_t spitch, int sw, int sh,
                                      unsigned char* __restrict__ dst, size_t dpitch, int dw, int dh) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= dw || y >= dh) return;
  const int D = 5;
  int tx = min(2 * x - D / 2 + D, sw - 1), ty = min(2 * y - D / 2 + D, sh - 1);
  float sum = 0.f;
  int count = 0;
  for (int cy = max(0, 2 * y - D / 2); cy < ty; ++cy) {
    const unsigned char* r = row_ptr(src, spitch, cy);
    for (int cx = max(0, 2 * x - D / 2); cx < tx; ++cx) {
      unsigned char s = __ldg(r + cx);
      if (s > 0) {
        float w = c_gauss25[(ty - cy - 1) * 5 + (tx - cx - 1)];
        sum += (float)s * w;
        count = (int)((float)count + w);
      }
    }
  }
  float q = sum / (float)count;
  int v = isnan(q) ? 0 : (int)q;
  row_ptr(dst, dpitch, y)[x] = (unsigned char)min(max(v, 0), 255);
}

// ---- a4: depth -> planar vertex map (cudafuncs.cu:109-134) -------------------------------------
__global__ void create_vmap_kernel(const float* __restrict__ depth, size_t dpitch, int W, int H,
                                   float fx_inv, float fy_inv, float cx, float cy, float cutoff,
                                   float* __restrict__ vmap, size_t vpitch) {
  int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y * blockDim.y + threadIdx.y;
  if (u >= W || v >= H) return;
  float z = __ldg(row_ptr(depth, dpitch, v) + u);
  float vx = qnan(), vy = qnan(), vz = qnan();
  if (z != 0 && z < cutoff) {
    vx = z * ((float)u - cx) * fx_inv;
    vy = z * ((float)v - cy) * fy_inv;
    vz = z;
  }
  row_ptr(vmap, vpitch, v)[u] = vx;
  row_ptr(vmap, vpitch, v + H)[u] = vy;
  row_ptr(vmap, vpitch, v + 2 * H)[u] = vz;
}

// ---- a4: forward-difference normals (cudafuncs.cu:152-189) -------------------------------------
__global__ void create_nmap_kernel(int rows, int cols, const float* __restrict__ vmap, size_t vpitch,
                                   float* __restrict__ nmap, size_t npitch) {
  int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y * blockDim.y + threadIdx.y;
  if (u >= cols || v >= rows) return;
  float3 r = make_float3(qnan(), qnan(), qnan());
  if (!(u == cols - 1 || v == rows - 1)) {
    float3 v00, v01, v10;
    v00.x = __ldg(row_ptr(vmap, vpitch, v) + u);
    v01.x = __ldg(row_ptr(vmap, vpitch, v) + u + 1);
    v10.x = __ldg(row_ptr(vmap, vpitch, v + 1) + u);
    if (!isnan(v00.x) && !isnan(v01.x) && !isnan(v10.x)) {
      v00.y = __ldg(row_ptr(vmap, vpitch, v + rows) + u);
      v01.y = __ldg(row_ptr(vmap, vpitch, v + rows) + u + 1);
      v10.y = __ldg(row_ptr(vmap, vpitch, v + 1 + rows) + u);
      v00.z = __ldg(row_ptr(vmap, vpitch, v + 2 * rows) + u);
      v01.z = __ldg(row_ptr(vmap, vpitch, v + 2 * rows) + u + 1);
      v10.z = __ldg(row_ptr(vmap, vpitch, v + 1 + 2 * rows) + u);
      r = normalized(cross(v01 - v00, v10 - v00));
    }
  }
  row_ptr(nmap, npitch, v)[u] = r.x;
  row_ptr(nmap, npitch, v + rows)[u] = r.y;
  row_ptr(nmap, npitch, v + 2 * rows)[u] = r.z;
}

// ---- a3: AoS float4 prediction -> planar maps, z==0 -> NaN (cudafuncs.cu:271-311) --------------
__global__ void copy_maps_kernel(int rows, int cols, const float4* __restrict__ v4, const float4* __restrict__ n4,
                                 float* __restrict__ vmap, size_t vpitch, float* __restrict__ nmap, size_t npitch) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= cols || y >= rows) return;
  float4 vs = __ldg(&v4[y * cols + x]), ns = __ldg(&n4[y * cols + x]);
  bool ok = !(vs.z == 0);
  float q = qnan();
  row_ptr(vmap, vpitch, y)[x] = ok ? vs.x : q;
  row_ptr(vmap, vpitch, y + rows)[x] = ok ? vs.y : q;
  row_ptr(vmap, vpitch, y + 2 * rows)[x] = ok ? vs.z : q;
  row_ptr(nmap, npitch, y)[x] = ok ? ns.x : q;
  row_ptr(nmap, npitch, y + rows)[x] = ok ? ns.y : q;
  row_ptr(nmap, npitch, y + 2 * rows)[x] = ok ? ns.z : q;
}

// ---- a3: 2x2 mean with NaN poisoning (cudafuncs.cu:366-417) ------------------------------------
template <bool NORMALIZE>
__global__ void resize_map_kernel(int drows, int dcols, int srows, const float* __restrict__ in, size_t ipitch,
                                  float* __restrict__ out, size_t opitch) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= dcols || y >= drows) return;
  int xs = 2 * x, ys = 2 * y;
  const float2 a0 = __ldg(reinterpret_cast<const float2*>(row_ptr(in, ipitch, ys) + xs));
  const float2 a1 = __ldg(reinterpret_cast<const float2*>(row_ptr(in, ipitch, ys + 1) + xs));
  float3 n = make_float3(qnan(), qnan(), qnan());
  if (!(isnan(a0.x) || isnan(a0.y) || isnan(a1.x) || isnan(a1.y))) {
    n.x = (a0.x + a0.y + a1.x + a1.y) / 4;
    const float2 b0 = __ldg(reinterpret_cast<const float2*>(row_ptr(in, ipitch, ys + srows) + xs));
    const float2 b1 = __ldg(reinterpret_cast<const float2*>(row_ptr(in, ipitch, ys + srows + 1) + xs));
    n.y = (b0.x + b0.y + b1.x + b1.y) / 4;
    const float2 c0 = __ldg(reinterpret_cast<const float2*>(row_ptr(in, ipitch, ys + 2 * srows) + xs));
    const float2 c1 = __ldg(reinterpret_cast<const float2*>(row_ptr(in, ipitch, ys + 2 * srows + 1) + xs));
    n.z = (c0.x + c0.y + c1.x + c1.y) / 4;
    if (NORMALIZE) n = normalized(n);
  }
  row_ptr(out, opitch, y)[x] = n.x;
  row_ptr(out, opitch, y + drows)[x] = n.y;
  row_ptr(out, opitch, y + 2 * drows)[x] = n.z;
}

// ---- a3: rigid transform of a map pair, in place allowed (cudafuncs.cu:207-249) ----------------
__global__ void transform_maps_kernel(int rows, int cols, const float* vsrc, size_t vspitch, const float* nsrc,
                                      size_t nspitch, Mat33 R, float3 t, float* vdst, size_t vdpitch,
                                      float* ndst, size_t ndpitch) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= cols || y >= rows) return;
  float3 vd = make_float3(qnan(), qnan(), qnan()), nd = vd;
  float vx = row_ptr(vsrc, vspitch, y)[x];
  if (!isnan(vx)) {
    float3 vs = make_float3(vx, row_ptr(vsrc, vspitch, y + rows)[x], row_ptr(vsrc, vspitch, y + 2 * rows)[x]);
    vd = mul(R, vs) + t;
  }
  float nx = row_ptr(nsrc, nspitch, y)[x];
  if (!isnan(nx)) {
    float3 ns = make_float3(nx, row_ptr(nsrc, nspitch, y + rows)[x], row_ptr(nsrc, nspitch, y + 2 * rows)[x]);
    nd = mul(R, ns);
  }
  row_ptr(vdst, vdpitch, y)[x] = vd.x;
  row_ptr(vdst, vdpitch, y + rows)[x] = vd.y;
  row_ptr(vdst, vdpitch, y + 2 * rows)[x] = vd.z;
  row_ptr(ndst, ndpitch, y)[x] = nd.x;
  row_ptr(ndst, ndpitch, y + rows)[x] = nd.y;
  row_ptr(ndst, ndpitch, y + 2 * rows)[x] = nd.z;
}

// ---- a5: z of AoS float4 vertices -> depth, out of range -> NaN (cudafuncs.cu:602-613) ---------
__global__ void vertices_to_depth_kernel(const float4* __restrict__ v4, int W, int H, float cutoff,
                                         float* __restrict__ dst, size_t dpitch) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  float z = __ldg(&v4[y * W + x]).z;
  row_ptr(dst, dpitch, y)[x] = (z > cutoff || z <= 0) ? qnan() : z;
}

// ---- a5: grey = (int)(0.114*c0 + 0.299*c1 + 0.587*c2) (cudafuncs.cu:634-638) ---------------------
// evaluated as fma(c2,.587, fma(c1,.299, c0*.114)): the contraction nvcc's default -fmad=true gives
// the reference expression (frozen choice shared with the oracle).
__global__ void rgb_to_intensity_kernel(const unsigned char* __restrict__ rgb, size_t pitch, int channels,
                                        int W, int H, unsigned char* __restrict__ dst, size_t dpitch) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  const unsigned char* p = row_ptr(rgb, pitch, y) + x * channels;
  float s = __fmaf_rn((float)p[2], 0.587f, __fmaf_rn((float)p[1], 0.299f, __fmul_rn((float)p[0], 0.114f)));
  row_ptr(dst, dpitch, y)[x] = (unsigned char)(int)s;
}

// ---- a6: 3x3 gradient taps over the border-clamped window (cudafuncs.cu:658-683) ---------------
__global__ void derivative_kernel(const unsigned char* __restrict__ src, size_t spitch, int W, int H,
                                  short* __restrict__ dx, short* __restrict__ dy, size_t gpitch) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  float dxVal = 0.f, dyVal = 0.f;
  int k = 8;
  for (int j = max(y - 1, 0); j <= min(y + 1, H - 1); j++) {
    const unsigned char* r = row_ptr(src, spitch, j);
    for (int i = max(x - 1, 0); i <= min(x + 1, W - 1); i++) {
      float p = (float)__ldg(r + i);
      dxVal += p * c_sobel_x[k];
      dyVal += p * c_sobel_y[k];
      --k;
    }
  }
  row_ptr(dx, gpitch, y)[x] = (short)dxVal;
  row_ptr(dy, gpitch, y)[x] = (short)dyVal;
}

// ---- a6: depth -> AoS float3 cloud (cudafuncs.cu:718-736) --------------------------------------
__global__ void project_points_kernel(const float* __restrict__ depth, size_t dpitch, int W, int H,
                                      float invFx, float invFy, float cx, float cy, float* __restrict__ cloud,
                                      size_t cpitch) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  float z = __ldg(row_ptr(depth, dpitch, y) + x);
  float* c = (float*)((char*)cloud + (size_t)y * cpitch) + 3 * x;
  c[0] = ((float)x - cx) * z * invFx;
  c[1] = ((float)y - cy) * z * invFy;
  c[2] = z;
}

}  // namespace

cudaError_t launch_bilateral(const float* depth, size_t dpitch, int W, int H, float maxD, float* out,
                             size_t opitch, cudaStream_t s) {
  bilateral_kernel<<<grid2d(W, H, kBlock), kBlock, 0, s>>>(depth, dpitch, W, H, maxD, out, opitch);
  return cudaGetLastError();
}
cudaError_t launch_pyr_down_gauss_f(const float* src, size_t spitch, int sw, int sh, float* dst,
                                    size_t dpitch, cudaStream_t s) {
  pyr_down_gauss_f_kernel<<<grid2d(sw / 2, sh / 2, kBlock), kBlock, 0, s>>>(src, spitch, sw, sh, dst, dpitch,
                                                                            sw / 2, sh / 2);
  return cudaGetLastError();
}
cudaError_t launch_pyr_down_uchar(const unsigned char* src, size_t spitch, int sw, int sh,
                                  unsigned char* dst, size_t dpitch, cudaStream_t s) {
  pyr_down_uchar_kernel<<<grid2d(sw / 2, sh / 2, kBlock), kBlock, 0, s>>>(src, spitch, sw, sh, dst, dpitch,
                                                                          sw / 2, sh / 2);
  return cudaGetLastError();
}
cudaError_t launch_create_vmap(const float* depth, size_t dpitch, int W, int H, Intr k, float cutoff,
                               float* vmap, size_t vpitch, cudaStream_t s) {
  create_vmap_kernel<<<grid2d(W, H, kBlock), kBlock, 0, s>>>(depth, dpitch, W, H, 1.f / k.fx, 1.f / k.fy, k.cx,
                                                             k.cy, cutoff, vmap, vpitch);
  return cudaGetLastError();
}
cudaError_t launch_create_nmap(const float* vmap, size_t vpitch, int W, int H, float* nmap, size_t npitch,
                               cudaStream_t s) {
  create_nmap_kernel<<<grid2d(W, H, kBlock), kBlock, 0, s>>>(H, W, vmap, vpitch, nmap, npitch);
  return cudaGetLastError();
}
cudaError_t launch_copy_maps(const float* v4, const float* n4, int W, int H, float* vmap, size_t vpitch,
                             float* nmap, size_t npitch, cudaStream_t s) {
  copy_maps_kernel<<<grid2d(W, H, kBlock), kBlock, 0, s>>>(H, W, (const float4*)v4, (const float4*)n4, vmap,
                                                           vpitch, nmap, npitch);
  return cudaGetLastError();
}
cudaError_t launch_resize_map(const float* in, size_t ipitch, int sw, int sh, bool normalize, float* out,
                              size_t opitch, cudaStream_t s) {
  dim3 g = grid2d(sw / 2, sh / 2, kBlock);
  if (normalize)
    resize_map_kernel<true><<<g, kBlock, 0, s>>>(sh / 2, sw / 2, sh, in, ipitch, out, opitch);
  else
    resize_map_kernel<false><<<g, kBlock, 0, s>>>(sh / 2, sw / 2, sh, in, ipitch, out, opitch);
  return cudaGetLastError();
}
cudaError_t launch_transform_maps(const float* vsrc, size_t vspitch, const float* nsrc, size_t nspitch, int W,
                                  int H, const Mat33& R, const float t[3], float* vdst, size_t vdpitch,
                                  float* ndst, size_t ndpitch, cudaStream_t s) {
  transform_maps_kernel<<<grid2d(W, H, kBlock), kBlock, 0, s>>>(H, W, vsrc, vspitch, nsrc, nspitch, R,
                                                                make_float3(t[0], t[1], t[2]), vdst, vdpitch,
                                                                ndst, ndpitch);
  return cudaGetLastError();
}
cudaError_t launch_vertices_to_depth(const float* v4, int W, int H, float cutoff, float* dst, size_t dpitch,
                                     cudaStream_t s) {
  vertices_to_depth_kernel<<<grid2d(W, H, kBlock), kBlock, 0, s>>>((const float4*)v4, W, H, cutoff, dst, dpitch);
  return cudaGetLastError();
}
cudaError_t launch_rgb_to_intensity(const unsigned char* rgb, size_t pitch, int channels, int W, int H,
                                    unsigned char* dst, size_t dpitch, cudaStream_t s) {
  rgb_to_intensity_kernel<<<grid2d(W, H, kBlock), kBlock, 0, s>>>(rgb, pitch, channels, W, H, dst, dpitch);
  return cudaGetLastError();
}
cudaError_t launch_derivative_images(const unsigned char* src, size_t spitch, int W, int H, short* dx, short* dy,
                                     size_t gpitch, cudaStream_t s) {
  derivative_kernel<<<grid2d(W, H, kBlock), kBlock, 0, s>>>(src, spitch, W, H, dx, dy, gpitch);
  return cudaGetLastError();
}
cudaError_t launch_project_to_point_cloud(const float* depth, size_t dpitch, int W, int H, Intr k, float* cloud,
                                          size_t cpitch, cudaStream_t s) {
  project_points_kernel<<<grid2d(W, H, kBlock), kBlock, 0, s>>>(depth, dpitch, W, H, 1.0f / k.fx, 1.0f / k.fy, k.cx,
                                                                k.cy, cloud, cpitch);
  return cudaGetLastError();
}

}  // namespace cfb
