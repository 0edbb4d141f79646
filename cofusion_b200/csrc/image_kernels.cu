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

#include <stdlib.h>
#include <string.h>

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
// Rows per thread.  Measured on B200 (VGA): 1 row/thread 78 us (1200 CTAs on 1184 resident slots: the
// last 16 run as a short second wave), 3 rows/thread 160 us -- the kernel lives on occupancy (64
// warps/SM hide the exp() dependency chains), so the single-row shape stays.
// A 640x480 frame is 1200 tiles of 32x8 on 1184 resident CTA slots (148 SMs x 8): the last 16 tiles run as a second,
// nearly empty wave.  9-row tiles (1080 CTAs, one wave, warp 0 takes a second row) were measured: no faster
// (0.620 vs 0.618 ms per frame) -- the kernels of the other stream fill that tail anyway.
constexpr int TR = 8;
__global__ void __launch_bounds__(256) bilateral_kernel(const float* __restrict__ depth, size_t dpitch, int W, int H,
                                                        float maxD, float* __restrict__ out, size_t opitch) {
  pdl_prologue();
  __shared__ float tile[TR + 2 * BR][32 + 2 * BR + 1];
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * TR;
  for (int ty = threadIdx.y; ty < TR + 2 * BR; ty += 8)
    for (int tx = threadIdx.x; tx < 32 + 2 * BR; tx += 32) {
      int gx = x0 + tx - BR, gy = y0 + ty - BR;
      float v = 0.f;
      if (gx >= 0 && gx < W && gy >= 0 && gy < H) v = __ldg(row_ptr(depth, dpitch, gy) + gx);
      tile[ty][tx] = v;
    }
  __syncthreads();
  const int x = x0 + threadIdx.x;
  if (x >= W) return;
#pragma unroll 1
  for (int ly = threadIdx.y; ly < TR; ly += 8) {
    const int y = y0 + ly;
    if (y >= H) return;
    const float value = tile[ly + BR][threadIdx.x + BR];
    float res = 0.f;
    if (!(value > maxD || value < 0.3f)) {
      const float sigma_space2_inv_half = 0.024691358f;
      const float sigma_color2_inv_half = 555.556f;
      const int D = 2 * BR + 1;
      float sum1 = 0.f, sum2 = 0.f;
      if (x >= BR && x + BR < W && y >= BR && y + BR < H) {
        // interior (97 % of a VGA frame): the window is the full 13x13, so the loops unroll completely
        // and the spatial term (float)x - (float)cx = -ox (exact) folds into one constant per tap; same
        // taps, same order, same roundings as the clamped loop below
        // two taps per step on the packed f32x2 pipe (det_expf2_nonpos: bit-identical lanes); the running sums
        // stay scalar and in tap order
        const float2 vv = make_float2(value, value);
        const float2 ncol = make_float2(-sigma_color2_inv_half, -sigma_color2_inv_half);
#pragma unroll
        for (int oy = -BR; oy <= BR; ++oy) {
#pragma unroll
          for (int ox = -BR; ox + 1 <= BR; ox += 2) {
            const float2 tmp = make_float2(tile[ly + BR + oy][threadIdx.x + BR + ox], tile[ly + BR + oy][threadIdx.x + BR + ox + 1]);
            const float dy = (float)(-oy), dx0 = (float)(-ox), dx1 = (float)(-ox - 1);
            // -(space2 * ss + color2 * sc) = (-(space2 * ss)) + color2 * (-sc): negation is exact
            const float2 nsp = make_float2(-((dx0 * dx0 + dy * dy) * sigma_space2_inv_half), -((dx1 * dx1 + dy * dy) * sigma_space2_inv_half));
            const float2 d = __fadd2_rn(vv, make_float2(-tmp.x, -tmp.y));
            const float2 color2 = __fmul2_rn(d, d);
            // this product and sum stay scalar: ptxas fuses mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 (seen in the
            // SASS, also with inline PTX and -fmad=false), which would change the rounding of the exponent
            const float2 arg = make_float2(__fadd_rn(nsp.x, __fmul_rn(color2.x, ncol.x)), __fadd_rn(nsp.y, __fmul_rn(color2.y, ncol.y)));
            const float2 weight = det_expf2_nonpos(arg);
            const float2 tw = __fmul2_rn(tmp, weight);
            sum1 += tw.x;
            sum2 += weight.x;
            sum1 += tw.y;
            sum2 += weight.y;
          }
          {  // the 13th tap of the row
            const int ox = BR;
            const float tmp = tile[ly + BR + oy][threadIdx.x + BR + ox];
            const float dx = (float)(-ox), dy = (float)(-oy);
            const float space2 = dx * dx + dy * dy;
            const float color2 = (value - tmp) * (value - tmp);
            const float weight = det_expf(-(space2 * sigma_space2_inv_half + color2 * sigma_color2_inv_half));
            sum1 += tmp * weight;
            sum2 += weight;
          }
        }
      } else {
        int tx = min(x - D / 2 + D, W), ty = min(y - D / 2 + D, H);
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
      }
      res = sum1 / sum2;
    }
    row_ptr(out, opitch, y)[x] = res;
  }
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

// ================================================================================================
// Fused pyramid builders used by Model::performTracking (identical arithmetic and operation order
// as the one-function-per-launch versions above; 28 launches per tracked frame become 10).

// copyMaps + resizeVMap/NMap x2 + tranformMaps x3 + verticesToDepth (RGBDOdometry.cpp:143-175,:179):
// one thread per level-2 pixel = 4x4 level-0 block.
struct ModelPyrOut {
  float *v[3], *n[3];  // planar, unpitched
  float* depth0;       // lastDepth level 0
};
__global__ void model_pyramid_kernel(const float4* __restrict__ v4, const float4* __restrict__ n4, int W, int H,
                                     Mat33 R, float3 t, const float* __restrict__ pose34_dev, float cutoffRGB,
                                     ModelPyrOut o, const float4* __restrict__ v4_alt, const float4* __restrict__ n4_alt,
                                     const unsigned* __restrict__ sel) {
  pdl_prologue();
  if (sel && *sel != 0) {  // fill-in images instead of the splat prediction (see PredAlt)
    v4 = v4_alt;
    n4 = n4_alt;
  }
  const int X = blockIdx.x * blockDim.x + threadIdx.x, Y = blockIdx.y * blockDim.y + threadIdx.y;
  const int W2 = W / 4, H2 = H / 4, W1 = W / 2, H1 = H / 2;
  if (X >= W2 || Y >= H2) return;
  if (pose34_dev) {  // the model pose lives on the device (row-major 3x4): no host round trip
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) R.m[r * 3 + c] = __ldg(pose34_dev + r * 4 + c);
    t = make_float3(__ldg(pose34_dev + 3), __ldg(pose34_dev + 7), __ldg(pose34_dev + 11));
  }
  const float q = qnan();
  float3 v0[4][4], n0[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int x = 4 * X + i, y = 4 * Y + j;
      const float4 vs = __ldg(&v4[y * W + x]), ns = __ldg(&n4[y * W + x]);
      const bool ok = !(vs.z == 0);
      v0[j][i] = ok ? make_float3(vs.x, vs.y, vs.z) : make_float3(q, q, q);
      n0[j][i] = ok ? make_float3(ns.x, ns.y, ns.z) : make_float3(q, q, q);
      o.depth0[y * W + x] = (vs.z > cutoffRGB || vs.z <= 0) ? q : vs.z;
    }
  float3 v1[2][2], n1[2][2];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float3 a = v0[2 * j][2 * i], b = v0[2 * j][2 * i + 1], c = v0[2 * j + 1][2 * i], d = v0[2 * j + 1][2 * i + 1];
      v1[j][i] = (isnan(a.x) || isnan(b.x) || isnan(c.x) || isnan(d.x))
                     ? make_float3(q, q, q)
                     : make_float3((a.x + b.x + c.x + d.x) / 4, (a.y + b.y + c.y + d.y) / 4, (a.z + b.z + c.z + d.z) / 4);
      const float3 e = n0[2 * j][2 * i], f = n0[2 * j][2 * i + 1], g = n0[2 * j + 1][2 * i], h = n0[2 * j + 1][2 * i + 1];
      n1[j][i] = (isnan(e.x) || isnan(f.x) || isnan(g.x) || isnan(h.x))
                     ? make_float3(q, q, q)
                     : normalized(make_float3((e.x + f.x + g.x + h.x) / 4, (e.y + f.y + g.y + h.y) / 4,
                                              (e.z + f.z + g.z + h.z) / 4));
    }
  float3 v2, n2;
  {
    const float3 a = v1[0][0], b = v1[0][1], c = v1[1][0], d = v1[1][1];
    v2 = (isnan(a.x) || isnan(b.x) || isnan(c.x) || isnan(d.x))
             ? make_float3(q, q, q)
             : make_float3((a.x + b.x + c.x + d.x) / 4, (a.y + b.y + c.y + d.y) / 4, (a.z + b.z + c.z + d.z) / 4);
    const float3 e = n1[0][0], f = n1[0][1], g = n1[1][0], h = n1[1][1];
    n2 = (isnan(e.x) || isnan(f.x) || isnan(g.x) || isnan(h.x))
             ? make_float3(q, q, q)
             : normalized(make_float3((e.x + f.x + g.x + h.x) / 4, (e.y + f.y + g.y + h.y) / 4, (e.z + f.z + g.z + h.z) / 4));
  }
  auto put = [&](float* vp, float* np, int w, int h, int x, int y, float3 v, float3 n) {
    float3 vd = make_float3(q, q, q), nd = vd;
    if (!isnan(v.x)) vd = mul(R, v) + t;
    if (!isnan(n.x)) nd = mul(R, n);
    vp[y * w + x] = vd.x;
    vp[(y + h) * w + x] = vd.y;
    vp[(y + 2 * h) * w + x] = vd.z;
    np[y * w + x] = nd.x;
    np[(y + h) * w + x] = nd.y;
    np[(y + 2 * h) * w + x] = nd.z;
  };
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) put(o.v[0], o.n[0], W, H, 4 * X + i, 4 * Y + j, v0[j][i], n0[j][i]);
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int i = 0; i < 2; ++i) put(o.v[1], o.n[1], W1, H1, 2 * X + i, 2 * Y + j, v1[j][i], n1[j][i]);
  put(o.v[2], o.n[2], W2, H2, X, Y, v2, n2);
}

// createVMap + createNMap for the three levels in one launch (RGBDOdometry.cpp:110-118)
struct FrameMapsArgs {
  const float* depth[3];
  float *v[3], *n[3];
  int w[3], h[3];
  float fx_inv[3], fy_inv[3], cx[3], cy[3];
  float cutoff;
  // optional: imageBGRToIntensity of up to two images of ni pixels rides along (the tracker's grey images)
  const unsigned char* img[2];
  unsigned char* grey[2];
  int ch[2], ni;
  const unsigned char* imgA_alt;  // image 0 comes from here when *selA != 0 or always (see PredAlt)
  const unsigned* selA;
  int altA_always;
};
__global__ void frame_maps_kernel(const FrameMapsArgs a) {
  pdl_prologue();
  int p = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    const int w = a.w[l], h = a.h[l], n = w * h;
    if (p < n) {
      const int v = p / w, u = p - v * w;
      const float* D = a.depth[l];
      const float fxi = a.fx_inv[l], fyi = a.fy_inv[l], cx = a.cx[l], cy = a.cy[l];
      const float q = qnan();
      auto vert = [&](int uu, int vv, bool& ok) {
        const float z = __ldg(D + vv * w + uu);
        ok = (z != 0 && z < a.cutoff);
        return make_float3(z * ((float)uu - cx) * fxi, z * ((float)vv - cy) * fyi, z);
      };
      bool ok00, ok01 = false, ok10 = false;
      const float3 v00 = vert(u, v, ok00);
      float* V = a.v[l];
      V[v * w + u] = ok00 ? v00.x : q;
      V[(v + h) * w + u] = ok00 ? v00.y : q;
      V[(v + 2 * h) * w + u] = ok00 ? v00.z : q;
      float3 r = make_float3(q, q, q);
      if (!(u == w - 1 || v == h - 1)) {
        const float3 v01 = vert(u + 1, v, ok01), v10 = vert(u, v + 1, ok10);
        if (ok00 && ok01 && ok10) r = normalized(cross(v01 - v00, v10 - v00));
      }
      float* N = a.n[l];
      N[v * w + u] = r.x;
      N[(v + h) * w + u] = r.y;
      N[(v + 2 * h) * w + u] = r.z;
      return;
    }
    p -= n;
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    if (p < a.ni) {
      if (a.img[k]) {
        const unsigned char* base = a.img[k];
        if (k == 0 && a.imgA_alt && (a.altA_always || (a.selA && *a.selA != 0))) base = a.imgA_alt;
        const unsigned char* q = base + (size_t)p * a.ch[k];
        const float s = __fmaf_rn((float)q[2], 0.587f, __fmaf_rn((float)q[1], 0.299f, __fmul_rn((float)q[0], 0.114f)));
        a.grey[k][p] = (unsigned char)(int)s;
      }
      return;
    }
    p -= a.ni;
  }
}

// imageBGRToIntensity for two images (model prediction RGBA8, frame RGB8) in one launch
__global__ void intensity2_kernel(const unsigned char* __restrict__ a, int cha, unsigned char* __restrict__ da,
                                  const unsigned char* __restrict__ b, int chb, unsigned char* __restrict__ db, int n) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned char* p = blockIdx.y ? (b + (size_t)i * chb) : (a + (size_t)i * cha);
  const float s = __fmaf_rn((float)p[2], 0.587f, __fmaf_rn((float)p[1], 0.299f, __fmul_rn((float)p[0], 0.114f)));
  (blockIdx.y ? db : da)[i] = (unsigned char)(int)s;
}
// pyrDownUcharGauss for two images in one launch
__global__ void pyr_down_uchar2_kernel(const unsigned char* __restrict__ sa, unsigned char* __restrict__ da,
                                       const unsigned char* __restrict__ sb, unsigned char* __restrict__ db, int sw,
                                       int sh) {
  const int dw = sw / 2, dh = sh / 2;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= dw || y >= dh) return;
  const unsigned char* src = blockIdx.z ? sb : sa;
  unsigned char* dst = blockIdx.z ? db : da;
  const int D = 5;
  const int tx = min(2 * x - D / 2 + D, sw - 1), ty = min(2 * y - D / 2 + D, sh - 1);
  float sum = 0.f;
  int count = 0;
  for (int cy = max(0, 2 * y - D / 2); cy < ty; ++cy)
    for (int cx = max(0, 2 * x - D / 2); cx < tx; ++cx) {
      const unsigned char s = __ldg(src + cy * sw + cx);
      if (s > 0) {
        const float w = c_gauss25[(ty - cy - 1) * 5 + (tx - cx - 1)];
        sum += (float)s * w;
        count = (int)((float)count + w);
      }
    }
  const float qv = sum / (float)count;
  const int v = isnan(qv) ? 0 : (int)qv;
  dst[y * dw + x] = (unsigned char)min(max(v, 0), 255);
}


// ---- a2/a5 fused: BOTH pyramid levels of up to four images in ONE launch.  The reference builds every level
// with its own launch (cudafuncs.cu:510-532, :566-588; 6 launches per frame here before), each a few microseconds of
// latency for a tiny output.  A CTA owns a 16x8 tile of level 2: it stages the 76x44 level-0 pixels that tile
// depends on in shared memory, produces the 36x20 level-1 pixels under it (writing the 32x16 it owns) and from
// those its level-2 pixels.  Per output pixel the taps, their order and every rounding are those of
// pyr_down_gauss_f_kernel / pyr_down_uchar_kernel above (the clamped window that drops the last row / column and
// indexes the weights from the window's end, the float -> int weight count): results are bit-identical.
struct PyrJob {
  const void* src;  // level 0, unpitched sw x sh
  void *l1, *l2;    // level 1 (sw/2 x sh/2), level 2 (sw/4 x sh/4)
  int is_u8;        // 0: f32, NaN = invalid; 1: u8, 0 = invalid
};
struct PyrJobs {
  PyrJob j[4];
  int sw, sh;
};
constexpr int P2W = 16, P2H = 8;                  // level-2 tile
constexpr int P1W = 2 * P2W + 4, P1H = 2 * P2H + 4;  // level-1 pixels under it: 36 x 20
constexpr int P0W = 2 * P1W + 4, P0H = 2 * P1H + 4;  // level-0 pixels under those: 76 x 44

// one output pixel (x, y) of a (sw x sh) -> (sw/2 x sh/2) reduction; the source is read through `at(cx, cy)`,
// invalid samples are NaN
template <class At>
__device__ __forceinline__ float gauss_down_pixel(int x, int y, int sw, int sh, At at) {
  const int D = 5;
  const int tx = min(2 * x - D / 2 + D, sw - 1), ty = min(2 * y - D / 2 + D, sh - 1);
  float sum = 0.f;
  int count = 0;
  for (int cy = max(0, 2 * y - D / 2); cy < ty; ++cy)
    for (int cx = max(0, 2 * x - D / 2); cx < tx; ++cx) {
      const float s = at(cx, cy);
      if (!isnan(s)) {
        const float w = c_gauss25[(ty - cy - 1) * 5 + (tx - cx - 1)];
        sum += s * w;
        count = (int)((float)count + w);
      }
    }
  return sum / (float)count;
}
__device__ __forceinline__ float u8_quantise(float q) {  // pyr_down_uchar_kernel's store, kept as a float (0 = invalid)
  const int v = isnan(q) ? 0 : (int)q;
  return (float)min(max(v, 0), 255);
}

__global__ void __launch_bounds__(256) pyramid2_kernel(const PyrJobs jobs) {
  pdl_prologue();
  __shared__ float t0[P0H][P0W + 1];
  __shared__ float t1[P1H][P1W + 1];
  const PyrJob job = jobs.j[blockIdx.z];
  const int sw = jobs.sw, sh = jobs.sh, w1 = sw / 2, h1 = sh / 2, w2 = sw / 4, h2 = sh / 4;
  const int X2 = blockIdx.x * P2W, Y2 = blockIdx.y * P2H;  // level-2 tile origin
  const int X1 = 2 * X2 - 2, Y1 = 2 * Y2 - 2;              // level-1 region origin
  const int X0 = 2 * X1 - 2, Y0 = 2 * Y1 - 2;              // level-0 region origin
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const float q = qnan();
  for (int i = tid; i < P0W * P0H; i += 256) {
    const int ly = i / P0W, lx = i - ly * P0W, gx = X0 + lx, gy = Y0 + ly;
    float v = q;
    if (gx >= 0 && gy >= 0 && gx < sw && gy < sh) {
      if (job.is_u8) {
        const unsigned char s = __ldg((const unsigned char*)job.src + (size_t)gy * sw + gx);
        v = s > 0 ? (float)s : q;
      } else {
        v = __ldg((const float*)job.src + (size_t)gy * sw + gx);
      }
    }
    t0[ly][lx] = v;
  }
  __syncthreads();
  for (int i = tid; i < P1W * P1H; i += 256) {
    const int ly = i / P1W, lx = i - ly * P1W, gx = X1 + lx, gy = Y1 + ly;
    float v = q;
    if (gx >= 0 && gy >= 0 && gx < w1 && gy < h1) {
      v = gauss_down_pixel(gx, gy, sw, sh, [&](int cx, int cy) { return t0[cy - Y0][cx - X0]; });
      const bool own = lx >= 2 && lx < P1W - 2 && ly >= 2 && ly < P1H - 2;
      if (job.is_u8) {
        v = u8_quantise(v);
        if (own) ((unsigned char*)job.l1)[(size_t)gy * w1 + gx] = (unsigned char)v;
        v = v > 0.f ? v : q;
      } else if (own) {
        ((float*)job.l1)[(size_t)gy * w1 + gx] = v;
      }
    }
    t1[ly][lx] = v;
  }
  __syncthreads();
  if (tid < P2W * P2H) {
    const int ly = tid / P2W, lx = tid - ly * P2W, gx = X2 + lx, gy = Y2 + ly;
    if (gx < w2 && gy < h2) {
      const float v = gauss_down_pixel(gx, gy, w1, h1, [&](int cx, int cy) { return t1[cy - Y1][cx - X1]; });
      if (job.is_u8)
        ((unsigned char*)job.l2)[(size_t)gy * w2 + gx] = (unsigned char)u8_quantise(v);
      else
        ((float*)job.l2)[(size_t)gy * w2 + gx] = v;
    }
  }
}

}  // namespace

// ---- f1: frame ingest on the device (GUI/Tools/KlgLogReader.cpp:53-84: raw u16 depth x 0.001 as cv::Mat::convertTo
// does it in f32; Core/FrameData.h:38-41: flipColors swaps the first and third channel)
__global__ void ingest_kernel(const uint8_t* __restrict__ img, const uint16_t* __restrict__ d16, float scale, int flip,
                              uint8_t* __restrict__ rgb, float* __restrict__ depth, int n) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (d16) depth[i] = __fmul_rn((float)__ldg(d16 + i), scale);
  if (img) {
    const uint8_t a = __ldg(img + 3 * i), b = __ldg(img + 3 * i + 1), c = __ldg(img + 3 * i + 2);
    rgb[3 * i] = flip ? c : a;
    rgb[3 * i + 1] = b;
    rgb[3 * i + 2] = flip ? a : c;
  }
}
cudaError_t launch_ingest(const uint8_t* img, const uint16_t* d16, float scale, int flip, uint8_t* rgb, float* depth, int n,
                          cudaStream_t s) {
  CFB_PDL(launch_pdl(ingest_kernel, (n + 255) / 256, 256, 0, s, img, d16, scale, flip, rgb, depth, n));
  return cudaGetLastError();
}

cudaError_t launch_pyramid2(int njobs, const void* const* src, void* const* l1, void* const* l2, const int* is_u8, int sw, int sh,
                            cudaStream_t s) {
  if (njobs < 1 || njobs > 4 || (sw % 4) || (sh % 4)) return cudaErrorInvalidValue;
  PyrJobs jobs;
  memset(&jobs, 0, sizeof(jobs));
  for (int k = 0; k < njobs; ++k) jobs.j[k] = PyrJob{src[k], l1[k], l2[k], is_u8[k]};
  jobs.sw = sw;
  jobs.sh = sh;
  const dim3 g((sw / 4 + P2W - 1) / P2W, (sh / 4 + P2H - 1) / P2H, njobs);
  CFB_PDL(launch_pdl(pyramid2_kernel, g, dim3(32, 8), 0, s, jobs));
  return cudaGetLastError();
}

cudaError_t launch_bilateral(const float* depth, size_t dpitch, int W, int H, float maxD, float* out,
                             size_t opitch, cudaStream_t s) {
  CFB_PDL(launch_pdl(bilateral_kernel, dim3((W + 31) / 32, (H + TR - 1) / TR), kBlock, 0, s, depth, dpitch, W, H, maxD, out, opitch));
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

cudaError_t launch_model_pyramid(const float* v4, const float* n4, int W, int H, const Mat33& R, const float t[3],
                                 float cutoffRGB, float* const v[3], float* const n[3], float* depth0,
                                 cudaStream_t s, const float* pose34_dev, const PredAlt* alt) {
  ModelPyrOut o;
  for (int i = 0; i < 3; ++i) {
    o.v[i] = v[i];
    o.n[i] = n[i];
  }
  o.depth0 = depth0;
  const dim3 b(32, 4);
  CFB_PDL(launch_pdl(model_pyramid_kernel, grid2d(W / 4, H / 4, b), b, 0, s, (const float4*)v4, (const float4*)n4, W, H, R,
                                                            make_float3(t[0], t[1], t[2]), pose34_dev, cutoffRGB, o,
                     (const float4*)(alt ? alt->v4 : nullptr), (const float4*)(alt ? alt->n4 : nullptr), alt ? alt->sel : nullptr));
  return cudaGetLastError();
}
cudaError_t launch_frame_maps(const float* const depth[3], int W, int H, Intr K, float cutoff, float* const v[3],
                              float* const n[3], cudaStream_t s, const unsigned char* imgA, int chA, unsigned char* greyA,
                              const unsigned char* imgB, int chB, unsigned char* greyB, const PredAlt* altA) {
  FrameMapsArgs a;
  a.imgA_alt = altA ? altA->img : nullptr;
  a.selA = altA ? altA->sel : nullptr;
  a.altA_always = altA ? altA->img_always : 0;
  int total = 0;
  a.img[0] = imgA;
  a.img[1] = imgB;
  a.grey[0] = greyA;
  a.grey[1] = greyB;
  a.ch[0] = chA;
  a.ch[1] = chB;
  a.ni = (imgA || imgB) ? W * H : 0;
  for (int l = 0; l < 3; ++l) {
    const Intr k = K.level(l);
    a.depth[l] = depth[l];
    a.v[l] = v[l];
    a.n[l] = n[l];
    a.w[l] = W >> l;
    a.h[l] = H >> l;
    a.fx_inv[l] = 1.f / k.fx;
    a.fy_inv[l] = 1.f / k.fy;
    a.cx[l] = k.cx;
    a.cy[l] = k.cy;
    total += a.w[l] * a.h[l];
  }
  a.cutoff = cutoff;
  total += 2 * a.ni;
  CFB_PDL(launch_pdl(frame_maps_kernel, (total + 255) / 256, 256, 0, s, a));
  return cudaGetLastError();
}
cudaError_t launch_intensity2(const unsigned char* a, int cha, unsigned char* da, const unsigned char* b, int chb,
                              unsigned char* db, int n, cudaStream_t s) {
  CFB_PDL(launch_pdl(intensity2_kernel, dim3((n + 255) / 256, 2), 256, 0, s, a, cha, da, b, chb, db, n));
  return cudaGetLastError();
}
cudaError_t launch_pyr_down_uchar2(const unsigned char* sa, unsigned char* da, const unsigned char* sb,
                                   unsigned char* db, int sw, int sh, cudaStream_t s) {
  dim3 g = grid2d(sw / 2, sh / 2, kBlock);
  g.z = 2;
  pyr_down_uchar2_kernel<<<g, kBlock, 0, s>>>(sa, da, sb, db, sw, sh);
  return cudaGetLastError();
}

}  // namespace cfb
