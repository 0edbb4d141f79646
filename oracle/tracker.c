/*
 * tracker.c -- CPU oracle (TEST INFRASTRUCTURE, see cf_oracle.h) for the projective ICP+RGB
 * tracker: image preparation kernels, the four reduction steps and the host Gauss-Newton loop.
 *
 * Every function cites the reference lines (relative to /root/reference) it restates.
 * Compile with -ffp-contract=off: the restatement is plain IEEE f32/f64, no FMA contraction.
 */
#include "cf_oracle.h"
#include "detmath.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NUM_PYRS 3

static inline float QNAN_BITS_F(void) {
  union {
    uint32_t u;
    float f;
  } c;
  c.u = 0x7fffffffu; /* Core/Cuda/cudafuncs.cu:131 */
  return c.f;
}
#define QNAN (QNAN_BITS_F())

/* __float2int_rn: round-to-nearest-even, NaN -> 0, saturating (CUDA semantics). */
static inline int f2i_rn(float v) {
  if (v != v) return 0;
  if (v >= 2147483648.0f) return 2147483647;
  if (v <= -2147483648.0f) return (-2147483647 - 1);
  return (int)lrintf(v);
}

typedef struct {
  float x, y, z;
} f3;
static inline f3 mk3(float x, float y, float z) {
  f3 r = {x, y, z};
  return r;
}
static inline f3 sub3(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline f3 add3(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline f3 cross3(f3 a, f3 b) {
  return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float norm3(f3 a) { return sqrtf(dot3(a, a)); }
/* Core/Cuda/operators.cuh:80-84 (rsqrtf) */
static inline f3 normalized3(f3 a) {
  float rn = 1.0f / sqrtf(dot3(a, a));
  return mk3(a.x * rn, a.y * rn, a.z * rn);
}
/* Core/Cuda/operators.cuh:86-89, row-major 3x3 */
static inline f3 mul33(const float* m, f3 a) {
  return mk3(dot3(mk3(m[0], m[1], m[2]), a), dot3(mk3(m[3], m[4], m[5]), a),
             dot3(mk3(m[6], m[7], m[8]), a));
}

/* ===================================================================== image preparation */

/* Core/Shaders/depth_bilateral_metric.frag:30-76 (via CoFusion::filterDepth, CoFusion.cpp:567-574).
 * Frozen GL semantics: nearest sampling, tap (cx,cy) reads texel (cx,cy); exp() = orc_expf. */
ORC_FMA_CLONES void orc_bilateral_filter(const float* depth, int W, int H, float maxD, float* out) {
  const float sigma_space2_inv_half = 0.024691358f;
  const float sigma_color2_inv_half = 555.556f;
  const int R = 6, D = R * 2 + 1;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      float value = depth[y * W + x];
      if (value > maxD || value < 0.3f) {
        out[y * W + x] = 0;
        continue;
      }
      int tx = x - D / 2 + D < W ? x - D / 2 + D : W;
      int ty = y - D / 2 + D < H ? y - D / 2 + D : H;
      float sum1 = 0, sum2 = 0;
      for (int cy = (y - D / 2 > 0 ? y - D / 2 : 0); cy < ty; ++cy)
        for (int cx = (x - D / 2 > 0 ? x - D / 2 : 0); cx < tx; ++cx) {
          float tmp = depth[cy * W + cx];
          float space2 = ((float)x - (float)cx) * ((float)x - (float)cx) +
                         ((float)y - (float)cy) * ((float)y - (float)cy);
          float color2 = (value - tmp) * (value - tmp);
          float weight = orc_expf(-(space2 * sigma_space2_inv_half + color2 * sigma_color2_inv_half));
          sum1 += tmp * weight;
          sum2 += weight;
        }
      out[y * W + x] = sum1 / sum2;
    }
}

static const float kGauss25[25] = {1, 4, 6, 4, 1, 4, 16, 24, 16, 4, 6, 24, 36,
                                   24, 6, 4, 16, 24, 16, 4, 1, 4, 6, 4, 1};

/* Core/Cuda/cudafuncs.cu:333-364 (pyrDownKernelGaussF), host :510-532. dst is (sh/2)x(sw/2). */
void orc_pyr_down_gauss_f(const float* src, int sw, int sh, float* dst) {
  int dw = sw / 2, dh = sh / 2;
  const int D = 5;
  for (int y = 0; y < dh; ++y)
    for (int x = 0; x < dw; ++x) {
      int tx = 2 * x - D / 2 + D < sw - 1 ? 2 * x - D / 2 + D : sw - 1;
      int ty = 2 * y - D / 2 + D < sh - 1 ? 2 * y - D / 2 + D : sh - 1;
      int cy = 2 * y - D / 2 > 0 ? 2 * y - D / 2 : 0;
      float sum = 0;
      int count = 0;
      for (; cy < ty; ++cy)
        for (int cx = (2 * x - D / 2 > 0 ? 2 * x - D / 2 : 0); cx < tx; ++cx) {
          float s = src[cy * sw + cx];
          if (!(s != s)) {
            float w = kGauss25[(ty - cy - 1) * 5 + (tx - cx - 1)];
            sum += s * w;
            count = (int)((float)count + w); /* `count += float` on an int, :359 */
          }
        }
      dst[y * dw + x] = (float)(sum / (float)count);
    }
}

/* Core/Cuda/cudafuncs.cu:534-564 (pyrDownKernelIntensityGauss). */
void orc_pyr_down_uchar_gauss(const uint8_t* src, int sw, int sh, uint8_t* dst) {
  int dw = sw / 2, dh = sh / 2;
  const int D = 5;
  for (int y = 0; y < dh; ++y)
    for (int x = 0; x < dw; ++x) {
      int tx = 2 * x - D / 2 + D < sw - 1 ? 2 * x - D / 2 + D : sw - 1;
      int ty = 2 * y - D / 2 + D < sh - 1 ? 2 * y - D / 2 + D : sh - 1;
      int cy = 2 * y - D / 2 > 0 ? 2 * y - D / 2 : 0;
      float sum = 0;
      int count = 0;
      for (; cy < ty; ++cy)
        for (int cx = (2 * x - D / 2 > 0 ? 2 * x - D / 2 : 0); cx < tx; ++cx) {
          uint8_t s = src[cy * sw + cx];
          if (s > 0) {
            float w = kGauss25[(ty - cy - 1) * 5 + (tx - cx - 1)];
            sum += (float)s * w;
            count = (int)((float)count + w);
          }
        }
      /* float -> uchar: CUDA cvt.rzi saturating, NaN (0/0) -> 0 */
      float q = sum / (float)count;
      int v = (q != q) ? 0 : (int)q;
      dst[y * dw + x] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
}

/* Core/Cuda/cudafuncs.cu:109-134 (computeVmapKernel); fx_inv = 1.f/fx computed on host :148.
 * Invalid pixels: the reference writes NaN to the x plane only; the oracle writes NaN to all three
 * planes (consumers only ever test the x plane). */
void orc_create_vmap(const float* depth, int W, int H, float fx, float fy, float cx, float cy,
                     float cutoff, float* vmap) {
  float fx_inv = 1.f / fx, fy_inv = 1.f / fy;
  for (int v = 0; v < H; ++v)
    for (int u = 0; u < W; ++u) {
      float z = depth[v * W + u];
      if (z != 0 && z < cutoff) {
        vmap[v * W + u] = z * ((float)u - cx) * fx_inv;
        vmap[(v + H) * W + u] = z * ((float)v - cy) * fy_inv;
        vmap[(v + 2 * H) * W + u] = z;
      } else {
        vmap[v * W + u] = QNAN;
        vmap[(v + H) * W + u] = QNAN;
        vmap[(v + 2 * H) * W + u] = QNAN;
      }
    }
}

/* Core/Cuda/cudafuncs.cu:152-189 (computeNmapKernel). */
void orc_create_nmap(const float* vmap, int W, int H, float* nmap) {
  for (int v = 0; v < H; ++v)
    for (int u = 0; u < W; ++u) {
      float* nx = &nmap[v * W + u];
      float* ny = &nmap[(v + H) * W + u];
      float* nz = &nmap[(v + 2 * H) * W + u];
      *nx = *ny = *nz = QNAN;
      if (u == W - 1 || v == H - 1) continue;
      f3 v00, v01, v10;
      v00.x = vmap[v * W + u];
      v01.x = vmap[v * W + u + 1];
      v10.x = vmap[(v + 1) * W + u];
      if (v00.x != v00.x || v01.x != v01.x || v10.x != v10.x) continue;
      v00.y = vmap[(v + H) * W + u];
      v01.y = vmap[(v + H) * W + u + 1];
      v10.y = vmap[(v + 1 + H) * W + u];
      v00.z = vmap[(v + 2 * H) * W + u];
      v01.z = vmap[(v + 2 * H) * W + u + 1];
      v10.z = vmap[(v + 1 + 2 * H) * W + u];
      f3 r = normalized3(cross3(sub3(v01, v00), sub3(v10, v00)));
      *nx = r.x;
      *ny = r.y;
      *nz = r.z;
    }
}

/* Core/Cuda/cudafuncs.cu:271-311 (copyMapsKernel). */
void orc_copy_maps(const float* v4, const float* n4, int W, int H, float* vmap, float* nmap) {
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const float* vs = &v4[(y * W + x) * 4];
      const float* ns = &n4[(y * W + x) * 4];
      int ok = !(vs[2] == 0);
      for (int k = 0; k < 3; ++k) {
        vmap[(y + k * H) * W + x] = ok ? vs[k] : QNAN;
        nmap[(y + k * H) * W + x] = ok ? ns[k] : QNAN;
      }
    }
}

/* Core/Cuda/cudafuncs.cu:366-417 (resizeMapKernel<normalize>). out is (sh/2)x(sw/2) planar. */
void orc_resize_map(const float* in, int sw, int sh, int normalize, float* out) {
  int dw = sw / 2, dh = sh / 2;
  for (int y = 0; y < dh; ++y)
    for (int x = 0; x < dw; ++x) {
      int xs = x * 2, ys = y * 2;
      float x00 = in[ys * sw + xs], x01 = in[ys * sw + xs + 1];
      float x10 = in[(ys + 1) * sw + xs], x11 = in[(ys + 1) * sw + xs + 1];
      if (x00 != x00 || x01 != x01 || x10 != x10 || x11 != x11) {
        out[y * dw + x] = QNAN;
        out[(y + dh) * dw + x] = QNAN;
        out[(y + 2 * dh) * dw + x] = QNAN;
        continue;
      }
      f3 n;
      n.x = (x00 + x01 + x10 + x11) / 4;
      const float* py = in + sh * sw;
      n.y = (py[ys * sw + xs] + py[ys * sw + xs + 1] + py[(ys + 1) * sw + xs] +
             py[(ys + 1) * sw + xs + 1]) / 4;
      const float* pz = in + 2 * sh * sw;
      n.z = (pz[ys * sw + xs] + pz[ys * sw + xs + 1] + pz[(ys + 1) * sw + xs] +
             pz[(ys + 1) * sw + xs + 1]) / 4;
      if (normalize) n = normalized3(n);
      out[y * dw + x] = n.x;
      out[(y + dh) * dw + x] = n.y;
      out[(y + 2 * dh) * dw + x] = n.z;
    }
}

/* Core/Cuda/cudafuncs.cu:207-249 (tranformMapsKernel). src may alias dst (the reference calls it
 * in place, RGBDOdometry.cpp:171). */
void orc_transform_maps(const float* vsrc, const float* nsrc, int W, int H, const float R[9],
                        const float t[3], float* vdst, float* ndst) {
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      f3 vd = mk3(QNAN, QNAN, QNAN), nd = mk3(QNAN, QNAN, QNAN);
      f3 vs;
      vs.x = vsrc[y * W + x];
      if (!(vs.x != vs.x)) {
        vs.y = vsrc[(y + H) * W + x];
        vs.z = vsrc[(y + 2 * H) * W + x];
        vd = add3(mul33(R, vs), mk3(t[0], t[1], t[2]));
      }
      f3 ns;
      ns.x = nsrc[y * W + x];
      if (!(ns.x != ns.x)) {
        ns.y = nsrc[(y + H) * W + x];
        ns.z = nsrc[(y + 2 * H) * W + x];
        nd = mul33(R, ns);
      }
      vdst[y * W + x] = vd.x;
      vdst[(y + H) * W + x] = vd.y;
      vdst[(y + 2 * H) * W + x] = vd.z;
      ndst[y * W + x] = nd.x;
      ndst[(y + H) * W + x] = nd.y;
      ndst[(y + 2 * H) * W + x] = nd.z;
    }
}

/* Core/Cuda/cudafuncs.cu:602-613 (verticesToDepthKernel). */
void orc_vertices_to_depth(const float* v4, int W, int H, float cutoff, float* depth) {
  for (int i = 0; i < W * H; ++i) {
    float z = v4[i * 4 + 2];
    depth[i] = (z > cutoff || z <= 0) ? QNAN : z;
  }
}

/* Core/Cuda/cudafuncs.cu:626-639 (bgr2IntensityKernel). The legacy texture-reference host wrapper
 * (:641-653) does not compile with CUDA 12; the arithmetic is restated from :634-638: the weights
 * are applied to bytes 0,1,2 of the uploaded RGB(A) texel in that order and the float sum is
 * truncated to int.  The sum is evaluated as fma(c2,.587, fma(c1,.299, c0*.114)) -- the contraction
 * nvcc's default -fmad=true produces for `a*A + b*B + c*C` (frozen choice, documented in
 * DESIGN.md). */
void orc_rgb_to_intensity(const uint8_t* rgb, int channels, int W, int H, uint8_t* grey) {
  for (int i = 0; i < W * H; ++i) {
    const uint8_t* p = &rgb[i * channels];
    float s = fmaf((float)p[2], 0.587f, fmaf((float)p[1], 0.299f, (float)p[0] * 0.114f));
    int value = (int)s;
    grey[i] = (uint8_t)value;
  }
}

/* Core/Cuda/cudafuncs.cu:655-715 (applyKernel + constant taps).  The tap index counts down from 8
 * over the border-CLAMPED window, so border pixels use a shifted subset of taps (quirk kept). */
void orc_derivative_images(const uint8_t* img, int W, int H, int16_t* dx, int16_t* dy) {
  static const float gsx[9] = {0.52201f, 0.00000f, -0.52201f, 0.79451f, -0.00000f,
                               -0.79451f, 0.52201f, 0.00000f, -0.52201f};
  static const float gsy[9] = {0.52201f, 0.79451f, 0.52201f, 0.00000f, 0.00000f,
                               0.00000f, -0.52201f, -0.79451f, -0.52201f};
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      float dxVal = 0, dyVal = 0;
      int k = 8;
      int j1 = y + 1 < H - 1 ? y + 1 : H - 1, i1 = x + 1 < W - 1 ? x + 1 : W - 1;
      for (int j = (y - 1 > 0 ? y - 1 : 0); j <= j1; ++j)
        for (int i = (x - 1 > 0 ? x - 1 : 0); i <= i1; ++i) {
          dxVal += (float)img[j * W + i] * gsx[k];
          dyVal += (float)img[j * W + i] * gsy[k];
          --k;
        }
      dx[y * W + x] = (int16_t)dxVal; /* float -> short, truncation */
      dy[y * W + x] = (int16_t)dyVal;
    }
}

/* Core/Cuda/cudafuncs.cu:718-736 (projectPointsKernel); invFx = 1.0f/fx on host :748. */
void orc_project_to_point_cloud(const float* depth, int W, int H, float fx, float fy, float cx,
                                float cy, float* cloud3) {
  float invFx = 1.0f / fx, invFy = 1.0f / fy;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      float z = depth[y * W + x];
      cloud3[(y * W + x) * 3 + 0] = (float)(((float)x - cx) * z * invFx);
      cloud3[(y * W + x) * 3 + 1] = (float)(((float)y - cy) * z * invFy);
      cloud3[(y * W + x) * 3 + 2] = z;
    }
}

/* ===================================================================== reduction steps */

static void unpack_se3(const double acc[29], float A[36], float b[6], float residual[2]) {
  /* Core/Cuda/reduce.cu:484-498 */
  int shift = 0;
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 7; ++j) {
      float value = (float)acc[shift++];
      if (j == 6)
        b[i] = value;
      else
        A[j * 6 + i] = A[i * 6 + j] = value;
    }
  if (residual) {
    residual[0] = (float)acc[27];
    residual[1] = (float)acc[28];
  }
}

static void accumulate_row7(double acc[29], const float row[7], int found) {
  /* Core/Cuda/reduce.cu:357-391: 27 upper-triangular products, row6^2, inlier flag.  Products are
   * f32 (as on the device); the oracle sums them in f64 so it is an order-independent witness of
   * the f32 tree reductions on either GPU implementation. */
  int k = 0;
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 7; ++j) acc[k++] += (double)(row[i] * row[j]);
  acc[27] += (double)(row[6] * row[6]);
  acc[28] += (double)(float)found;
}

/* Core/Cuda/reduce.cu:257-499 (ICPReduction::search / getProducts, icpStep host unpack). */
void orc_icp_step(const float Rcurr[9], const float tcurr[3], const float* vmap_curr,
                  const float* nmap_curr, const float Rprev_inv[9], const float tprev[3], float fx,
                  float fy, float cx, float cy, const float* vmap_g_prev, const float* nmap_g_prev,
                  float distThres, float angleThres, int W, int H, float A[36], float b[6],
                  float residual[2], float* error_map) {
  double acc[29];
  memset(acc, 0, sizeof(acc));
  const f3 tc = mk3(tcurr[0], tcurr[1], tcurr[2]), tp = mk3(tprev[0], tprev[1], tprev[2]);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      float row[7] = {0, 0, 0, 0, 0, 0, 0};
      int found = 0;
      f3 vcurr = mk3(vmap_curr[y * W + x], vmap_curr[(y + H) * W + x],
                     vmap_curr[(y + 2 * H) * W + x]);
      f3 vcurr_g = add3(mul33(Rcurr, vcurr), tc);
      f3 vcurr_cp = mul33(Rprev_inv, sub3(vcurr_g, tp));
      int ux = f2i_rn(vcurr_cp.x * fx / vcurr_cp.z + cx);
      int uy = f2i_rn(vcurr_cp.y * fy / vcurr_cp.z + cy);
      if (ux < 0 || uy < 0 || ux >= W || uy >= H || vcurr_cp.z < 0) {
        if (error_map) error_map[y * W + x] = 0.0f;
      } else {
        f3 vprev_g = mk3(vmap_g_prev[uy * W + ux], vmap_g_prev[(uy + H) * W + ux],
                         vmap_g_prev[(uy + 2 * H) * W + ux]);
        f3 ncurr = mk3(nmap_curr[y * W + x], nmap_curr[(y + H) * W + x],
                       nmap_curr[(y + 2 * H) * W + x]);
        f3 ncurr_g = mul33(Rcurr, ncurr);
        f3 nprev_g = mk3(nmap_g_prev[uy * W + ux], nmap_g_prev[(uy + H) * W + ux],
                         nmap_g_prev[(uy + 2 * H) * W + ux]);
        float dist = norm3(sub3(vprev_g, vcurr_g));
        float sine = norm3(cross3(ncurr_g, nprev_g));
        if (error_map) error_map[y * W + x] = isfinite(dist) ? dist : 0.0f;
        found = (sine < angleThres && dist <= distThres && !(ncurr.x != ncurr.x) &&
                 !(nprev_g.x != nprev_g.x));
        if (found) {
          f3 s_cp = mul33(Rprev_inv, sub3(vcurr_g, tp));
          f3 d_cp = mul33(Rprev_inv, sub3(vprev_g, tp));
          f3 n_cp = mul33(Rprev_inv, nprev_g);
          f3 c = cross3(s_cp, n_cp);
          row[0] = n_cp.x;
          row[1] = n_cp.y;
          row[2] = n_cp.z;
          row[3] = c.x;
          row[4] = c.y;
          row[5] = c.z;
          row[6] = dot3(n_cp, sub3(s_cp, d_cp));
        }
      }
      accumulate_row7(acc, row, found);
    }
  unpack_se3(acc, A, b, residual);
}

/* Core/Cuda/reduce.cu:748-971 (RGBResidual::getProducts, computeRgbResidual).  sigmaSum is the
 * wrapping int32 sum of (int)(diff*diff) (:851, may overflow -- kept). */
void orc_rgb_residual(float minScale, const int16_t* dIdx, const int16_t* dIdy,
                      const float* lastDepth, const float* nextDepth, const uint8_t* lastImage,
                      const uint8_t* nextImage, OrcDataTerm* corres, float maxDepthDelta,
                      const float kt[3], const float krkinv[9], int W, int H, int* sigmaSum,
                      int* count) {
  uint32_t sig = 0, cnt = 0;
  for (int i = 0; i < H; ++i)
    for (int j0 = 0; j0 < W; ++j0) {
      OrcDataTerm c;
      memset(&c, 0, sizeof(c));
      if (j0 < W - 5 && i < H - 1) {
        int valid = 1;
        int u1 = i + 2 < H ? i + 2 : H, v1 = j0 + 2 < W ? j0 + 2 : W;
        for (int u = (i - 2 > 0 ? i - 2 : 0); u < u1; u++)
          for (int v = (j0 - 2 > 0 ? j0 - 2 : 0); v < v1; v++)
            valid = valid && (nextImage[u * W + v] > 0);
        if (valid) {
          short valx = dIdx[i * W + j0], valy = dIdy[i * W + j0];
          float mTwo = (float)((valx * valx) + (valy * valy));
          if (mTwo >= minScale) {
            int y = i, x = j0;
            float d1 = nextDepth[y * W + x];
            if (!(d1 != d1)) {
              float fx_ = (float)x, fy_ = (float)y;
              float transformed_d1 =
                  (float)(d1 * (krkinv[6] * fx_ + krkinv[7] * fy_ + krkinv[8]) + kt[2]);
              int u0 = f2i_rn((d1 * (krkinv[0] * fx_ + krkinv[1] * fy_ + krkinv[2]) + kt[0]) /
                              transformed_d1);
              int v0 = f2i_rn((d1 * (krkinv[3] * fx_ + krkinv[4] * fy_ + krkinv[5]) + kt[1]) /
                              transformed_d1);
              if (u0 >= 0 && v0 >= 0 && u0 < W && v0 < H) {
                float d0 = lastDepth[v0 * W + u0];
                if (d0 > 0 && fabsf(transformed_d1 - d0) <= maxDepthDelta &&
                    lastImage[v0 * W + u0] != 0) {
                  c.zero_x = (int16_t)u0;
                  c.zero_y = (int16_t)v0;
                  c.one_x = (int16_t)x;
                  c.one_y = (int16_t)y;
                  c.diff = (float)nextImage[y * W + x] - (float)lastImage[v0 * W + u0];
                  c.valid = 1;
                  cnt += 1;
                  sig += (uint32_t)(int)(c.diff * c.diff);
                }
              }
            }
          }
        }
      }
      corres[i * W + j0] = c;
    }
  *count = (int)cnt;
  *sigmaSum = (int)sig;
}

/* Core/Cuda/reduce.cu:503-687 (RGBReduction::getProducts, rgbStep). */
void orc_rgb_step(const OrcDataTerm* corres, float sigma, const float* cloud3, float fx, float fy,
                  const int16_t* dIdx, const int16_t* dIdy, float sobelScale, int W, int H,
                  float A[36], float b[6]) {
  double acc[29];
  memset(acc, 0, sizeof(acc));
  for (int i = 0; i < W * H; ++i) {
    const OrcDataTerm* c = &corres[i];
    float row[7] = {0, 0, 0, 0, 0, 0, 0};
    int found = c->valid;
    if (found) {
      float w = sigma + fabsf(c->diff);
      w = w > 1.19209290E-07F ? 1.0f / w : 1.0f;
      if (sigma == -1) w = 1;
      row[6] = -w * c->diff;
      const float* P = &cloud3[(c->zero_y * W + c->zero_x) * 3];
      float invz = (float)(1.0 / (double)P[2]);
      float dI_dx_val = w * sobelScale * (float)dIdx[c->one_y * W + c->one_x];
      float dI_dy_val = w * sobelScale * (float)dIdy[c->one_y * W + c->one_x];
      float v0 = dI_dx_val * fx * invz;
      float v1 = dI_dy_val * fy * invz;
      float v2 = -(v0 * P[0] + v1 * P[1]) * invz;
      row[0] = v0;
      row[1] = v1;
      row[2] = v2;
      row[3] = -P[2] * v1 + P[1] * v2;
      row[4] = P[2] * v0 - P[0] * v2;
      row[5] = -P[1] * v0 + P[0] * v1;
    }
    accumulate_row7(acc, row, found);
  }
  unpack_se3(acc, A, b, NULL);
}

/* Core/Cuda/reduce.cu:973-1176 (SO3Reduction, so3Step). */
static void so3_gradient(const uint8_t* img, int W, int x, int y, float* gx, float* gy) {
  float actu = (float)img[y * W + x];
  float back = (float)img[y * W + x - 1], fore = (float)img[y * W + x + 1];
  *gx = ((back + actu) / 2.0f) - ((fore + actu) / 2.0f);
  back = (float)img[(y - 1) * W + x];
  fore = (float)img[(y + 1) * W + x];
  *gy = ((back + actu) / 2.0f) - ((fore + actu) / 2.0f);
}

void orc_so3_step(const uint8_t* lastImage, const uint8_t* nextImage, const float imageBasis[9],
                  const float kinv[9], const float krlr[9], int W, int H, float A[9], float bvec[3],
                  float residual[2]) {
  double acc[11];
  memset(acc, 0, sizeof(acc));
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      f3 unwarped = mk3((float)x, (float)y, 1.0f);
      f3 warped = mul33(imageBasis, unwarped);
      int wx = f2i_rn(warped.x / warped.z), wy = f2i_rn(warped.y / warped.z);
      int found = (wx >= 1 && wx < W - 1 && wy >= 1 && wy < H - 1 && x >= 1 && x < W - 1 &&
                   y >= 1 && y < H - 1);
      float row[4] = {0, 0, 0, 0};
      if (found) {
        float gnx, gny, glx, gly;
        so3_gradient(nextImage, W, wx, wy, &gnx, &gny);
        so3_gradient(lastImage, W, x, y, &glx, &gly);
        float gx = (gnx + glx) / 2.0f, gy = (gny + gly) / 2.0f;
        f3 point = mul33(kinv, unwarped);
        float z2 = point.z * point.z;
        float a = krlr[0], b = krlr[1], c = krlr[2], d = krlr[3], e = krlr[4], f = krlr[5],
              g = krlr[6], h = krlr[7], i = krlr[8];
        float fx_ = (float)x, fy_ = (float)y;
        f3 left = mk3(((point.z * (d * gy + a * gx)) - (gy * g * fy_) - (gx * g * fx_)) / z2,
                      ((point.z * (e * gy + b * gx)) - (gy * h * fy_) - (gx * h * fx_)) / z2,
                      ((point.z * (f * gy + c * gx)) - (gy * i * fy_) - (gx * i * fx_)) / z2);
        f3 jac = cross3(left, point);
        row[0] = jac.x;
        row[1] = jac.y;
        row[2] = jac.z;
        row[3] = -((float)nextImage[wy * W + wx] - (float)lastImage[y * W + x]);
      }
      int k = 0;
      for (int p = 0; p < 3; ++p)
        for (int q = p; q < 4; ++q) acc[k++] += (double)(row[p] * row[q]);
      acc[9] += (double)(row[3] * row[3]);
      acc[10] += (double)(float)found;
    }
  int shift = 0;
  for (int i = 0; i < 3; ++i)
    for (int j = i; j < 4; ++j) {
      float value = (float)acc[shift++];
      if (j == 3)
        bvec[i] = value;
      else
        A[j * 3 + i] = A[i * 3 + j] = value;
    }
  residual[0] = (float)acc[9];
  residual[1] = (float)acc[10];
}

/* ===================================================================== small dense linear algebra
 * (stand-ins for the Eigen calls of RGBDOdometry.cpp; Eigen is not in this image) */

/* x = A^-1 b via LDL^T with diagonal pivoting (Eigen::LDLT picks the largest remaining |diagonal|
 * each step; RGBDOdometry.cpp:298, :435). A is row-major n x n symmetric, n <= 6. */
static void ldlt_solve_d(const double* Ain, const double* bin, int n, double* x) {
  double A[36], b[6];
  int perm[6];
  memcpy(A, Ain, sizeof(double) * n * n);
  for (int i = 0; i < n; ++i) {
    perm[i] = i;
    b[i] = bin[i];
  }
  for (int k = 0; k < n; ++k) {
    int p = k;
    double best = fabs(A[k * n + k]);
    for (int i = k + 1; i < n; ++i)
      if (fabs(A[i * n + i]) > best) {
        best = fabs(A[i * n + i]);
        p = i;
      }
    if (p != k) { /* symmetric row/column swap */
      for (int j = 0; j < n; ++j) {
        double t = A[k * n + j];
        A[k * n + j] = A[p * n + j];
        A[p * n + j] = t;
      }
      for (int i = 0; i < n; ++i) {
        double t = A[i * n + k];
        A[i * n + k] = A[i * n + p];
        A[i * n + p] = t;
      }
      int ti = perm[k];
      perm[k] = perm[p];
      perm[p] = ti;
      double tb = b[k];
      b[k] = b[p];
      b[p] = tb;
    }
    double d = A[k * n + k];
    if (d == 0) continue;
    for (int i = k + 1; i < n; ++i) {
      double l = A[i * n + k] / d;
      for (int j = k + 1; j < n; ++j) A[i * n + j] -= l * A[k * n + j];
      A[i * n + k] = l;
    }
  }
  /* forward L y = b */
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j) b[i] -= A[i * n + j] * b[j];
  for (int i = 0; i < n; ++i) b[i] = A[i * n + i] != 0 ? b[i] / A[i * n + i] : 0;
  for (int i = n - 1; i >= 0; --i)
    for (int j = i + 1; j < n; ++j) b[i] -= A[j * n + i] * b[j];
  for (int i = 0; i < n; ++i) x[perm[i]] = b[i];
}

/* Core/Utils/OdometryProvider.h:32-67 */
static void rodrigues_d(const double src[3], double R[9]) {
  static const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  memcpy(R, I, sizeof(I));
  double rx = src[0], ry = src[1], rz = src[2];
  double theta = sqrt(rx * rx + ry * ry + rz * rz);
  if (theta >= DBL_EPSILON) {
    double c = cos(theta), s = sin(theta), c1 = 1. - c, itheta = theta ? 1. / theta : 0.;
    rx *= itheta;
    ry *= itheta;
    rz *= itheta;
    double rrt[9] = {rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz};
    double r_x[9] = {0, -rz, ry, rz, 0, -rx, -ry, rx, 0};
    for (int k = 0; k < 9; k++) R[k] = c * I[k] + c1 * rrt[k] + s * r_x[k];
  }
}

static void mat3_mul_d(const double* a, const double* b, double* c) {
  double r[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += a[i * 3 + k] * b[k * 3 + j];
      r[i * 3 + j] = s;
    }
  memcpy(c, r, sizeof(r));
}
static void mat4_mul_d(const double* a, const double* b, double* c) {
  double r[16];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[k * 4 + j];
      r[i * 4 + j] = s;
    }
  memcpy(c, r, sizeof(r));
}
/* inverse of a rigid 4x4 (the reference calls the general Eigen inverse, RGBDOdometry.cpp:348) */
static void rigid_inverse_d(const double* T, double* Ti) {
  double r[16] = {0};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r[i * 4 + j] = T[j * 4 + i];
  for (int i = 0; i < 3; ++i)
    r[i * 4 + 3] = -(r[i * 4 + 0] * T[3] + r[i * 4 + 1] * T[7] + r[i * 4 + 2] * T[11]);
  r[15] = 1;
  memcpy(Ti, r, sizeof(r));
}
/* f32 3x3 inverse by cofactors (Eigen fixed-size inverse, RGBDOdometry.cpp:316) */
static void mat3_inverse_f(const float* m, float* inv) {
  float c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8],
        c02 = m[3] * m[7] - m[4] * m[6];
  float det = m[0] * c00 + m[1] * c01 + m[2] * c02;
  float id = 1.0f / det;
  inv[0] = c00 * id;
  inv[1] = (m[2] * m[7] - m[1] * m[8]) * id;
  inv[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  inv[3] = c01 * id;
  inv[4] = (m[0] * m[8] - m[2] * m[6]) * id;
  inv[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  inv[6] = c02 * id;
  inv[7] = (m[1] * m[6] - m[0] * m[7]) * id;
  inv[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

/* ===================================================================== RGBDOdometry restatement */

struct OrcOdometry {
  int W, H;
  float cx, cy, fx, fy, distThres, angleThres;
  float sobelScale, maxDepthDeltaRGB, maxDepthRGB;
  float minGrad[NUM_PYRS];
  float *vmaps_g_prev[NUM_PYRS], *nmaps_g_prev[NUM_PYRS], *vmaps_curr[NUM_PYRS],
      *nmaps_curr[NUM_PYRS];
  float *lastDepth[NUM_PYRS], *nextDepth[NUM_PYRS], *depthPyr[NUM_PYRS], *cloud[NUM_PYRS];
  uint8_t *lastImage[NUM_PYRS], *nextImage[NUM_PYRS], *lastNextImage[NUM_PYRS];
  int16_t *dIdx[NUM_PYRS], *dIdy[NUM_PYRS];
  OrcDataTerm* corres[NUM_PYRS];
  float* vmaps_tmp; /* AoS float4, RGBDOdometry.cpp:99 */
};

OrcOdometry* orc_odom_create(int W, int H, float cx, float cy, float fx, float fy, float distThresh,
                             float angleThresh) {
  OrcOdometry* o = (OrcOdometry*)calloc(1, sizeof(OrcOdometry));
  o->W = W;
  o->H = H;
  o->cx = cx;
  o->cy = cy;
  o->fx = fx;
  o->fy = fy;
  o->distThres = distThresh;
  o->angleThres = angleThresh;
  /* RGBDOdometry.cpp:31-34, :103-105 */
  o->sobelScale = (float)(1.0 / pow(2.0, 3));
  o->maxDepthDeltaRGB = 0.07f;
  o->maxDepthRGB = 6.0f;
  o->minGrad[0] = 5;
  o->minGrad[1] = 3;
  o->minGrad[2] = 1;
  for (int i = 0; i < NUM_PYRS; ++i) {
    size_t n = (size_t)(W >> i) * (H >> i);
    o->vmaps_g_prev[i] = (float*)calloc(n * 3, 4);
    o->nmaps_g_prev[i] = (float*)calloc(n * 3, 4);
    o->vmaps_curr[i] = (float*)calloc(n * 3, 4);
    o->nmaps_curr[i] = (float*)calloc(n * 3, 4);
    o->lastDepth[i] = (float*)calloc(n, 4);
    o->nextDepth[i] = (float*)calloc(n, 4);
    o->depthPyr[i] = (float*)calloc(n, 4);
    o->cloud[i] = (float*)calloc(n * 3, 4);
    o->lastImage[i] = (uint8_t*)calloc(n, 1);
    o->nextImage[i] = (uint8_t*)calloc(n, 1);
    o->lastNextImage[i] = (uint8_t*)calloc(n, 1);
    o->dIdx[i] = (int16_t*)calloc(n, 2);
    o->dIdy[i] = (int16_t*)calloc(n, 2);
    o->corres[i] = (OrcDataTerm*)calloc(n, sizeof(OrcDataTerm));
  }
  o->vmaps_tmp = (float*)calloc((size_t)W * H * 4, 4);
  return o;
}

void orc_odom_destroy(OrcOdometry* o) {
  if (!o) return;
  for (int i = 0; i < NUM_PYRS; ++i) {
    free(o->vmaps_g_prev[i]);
    free(o->nmaps_g_prev[i]);
    free(o->vmaps_curr[i]);
    free(o->nmaps_curr[i]);
    free(o->lastDepth[i]);
    free(o->nextDepth[i]);
    free(o->depthPyr[i]);
    free(o->cloud[i]);
    free(o->lastImage[i]);
    free(o->nextImage[i]);
    free(o->lastNextImage[i]);
    free(o->dIdx[i]);
    free(o->dIdy[i]);
    free(o->corres[i]);
  }
  free(o->vmaps_tmp);
  free(o);
}

/* RGBDOdometry.cpp:177-194 (populateRGBDData): depth-from-vmaps_tmp + intensity pyramids. */
static void populate_rgbd(OrcOdometry* o, const uint8_t* img, int ch, float** destDepths,
                          uint8_t** destImages) {
  orc_vertices_to_depth(o->vmaps_tmp, o->W, o->H, o->maxDepthRGB, destDepths[0]);
  for (int i = 0; i + 1 < NUM_PYRS; i++)
    orc_pyr_down_gauss_f(destDepths[i], o->W >> i, o->H >> i, destDepths[i + 1]);
  orc_rgb_to_intensity(img, ch, o->W, o->H, destImages[0]);
  for (int i = 0; i + 1 < NUM_PYRS; i++)
    orc_pyr_down_uchar_gauss(destImages[i], o->W >> i, o->H >> i, destImages[i + 1]);
}

/* RGBDOdometry.cpp:143-175 (initICPModel) + :196-199 (initRGBModel). */
void orc_odom_init_model(OrcOdometry* o, const float* v4, const float* n4, const uint8_t* img,
                         int img_channels, const float pose[16]) {
  memcpy(o->vmaps_tmp, v4, (size_t)o->W * o->H * 16);
  orc_copy_maps(v4, n4, o->W, o->H, o->vmaps_g_prev[0], o->nmaps_g_prev[0]);
  for (int i = 1; i < NUM_PYRS; ++i) {
    orc_resize_map(o->vmaps_g_prev[i - 1], o->W >> (i - 1), o->H >> (i - 1), 0, o->vmaps_g_prev[i]);
    orc_resize_map(o->nmaps_g_prev[i - 1], o->W >> (i - 1), o->H >> (i - 1), 1, o->nmaps_g_prev[i]);
  }
  float R[9] = {pose[0], pose[1], pose[2], pose[4], pose[5], pose[6], pose[8], pose[9], pose[10]};
  float t[3] = {pose[3], pose[7], pose[11]};
  for (int i = 0; i < NUM_PYRS; ++i)
    orc_transform_maps(o->vmaps_g_prev[i], o->nmaps_g_prev[i], o->W >> i, o->H >> i, R, t,
                       o->vmaps_g_prev[i], o->nmaps_g_prev[i]);
  populate_rgbd(o, img, img_channels, o->lastDepth, o->lastImage);
}

/* Model::generateCUDATextures (Model.cpp:319-348) + RGBDOdometry::initICP(pyramid) (:110-118)
 * + initRGB (:201-204).  Quirk kept: initRGB reads vmaps_tmp, which still holds the MODEL
 * prediction, so nextDepth == lastDepth (SURVEY.md section 7, quirk list). */
void orc_odom_init_frame(OrcOdometry* o, const float* depth_filtered, const uint8_t* rgb,
                         int img_channels, float depthCutoff) {
  memcpy(o->depthPyr[0], depth_filtered, (size_t)o->W * o->H * 4);
  for (int i = 1; i < NUM_PYRS; ++i)
    orc_pyr_down_gauss_f(o->depthPyr[i - 1], o->W >> (i - 1), o->H >> (i - 1), o->depthPyr[i]);
  for (int i = 0; i < NUM_PYRS; ++i) {
    int div = 1 << i;
    orc_create_vmap(o->depthPyr[i], o->W >> i, o->H >> i, o->fx / div, o->fy / div, o->cx / div,
                    o->cy / div, depthCutoff, o->vmaps_curr[i]);
    orc_create_nmap(o->vmaps_curr[i], o->W >> i, o->H >> i, o->nmaps_curr[i]);
  }
  populate_rgbd(o, rgb, img_channels, o->nextDepth, o->nextImage);
}

/* RGBDOdometry.cpp:206-215 */
void orc_odom_init_first_rgb(OrcOdometry* o, const uint8_t* rgb, int img_channels) {
  orc_rgb_to_intensity(rgb, img_channels, o->W, o->H, o->lastNextImage[0]);
  for (int i = 0; i + 1 < NUM_PYRS; i++)
    orc_pyr_down_uchar_gauss(o->lastNextImage[i], o->W >> i, o->H >> i, o->lastNextImage[i + 1]);
}

const void* orc_odom_view(OrcOdometry* o, int which, int level) {
  switch (which) {
    case 0: return o->vmaps_curr[level];
    case 1: return o->nmaps_curr[level];
    case 2: return o->vmaps_g_prev[level];
    case 3: return o->nmaps_g_prev[level];
    case 4: return o->lastDepth[level];
    case 5: return o->nextDepth[level];
    case 6: return o->lastImage[level];
    case 7: return o->nextImage[level];
    case 8: return o->dIdx[level];
    case 9: return o->dIdy[level];
    case 10: return o->lastNextImage[level];
    case 11: return o->cloud[level];
  }
  return 0;
}

static void K_of(const OrcOdometry* o, int level, double K[9], double Kinv[9]) {
  int div = 1 << level;
  /* CameraModel::operator()(level) divides in f32 (types.cuh:94-98) */
  float fx = o->fx / div, fy = o->fy / div, cx = o->cx / div, cy = o->cy / div;
  double k[9] = {fx, 0, cx, 0, fy, cy, 0, 0, 1};
  memcpy(K, k, sizeof(k));
  double ki[9] = {1.0 / fx, 0, -(double)cx / fx, 0, 1.0 / fy, -(double)cy / fy, 0, 0, 1};
  memcpy(Kinv, ki, sizeof(ki));
}

/* ---- default (CPU) step backend ---- */
static void cpu_so3(void* u, OrcOdometry* o, int L, const float* ib, const float* kinv,
                    const float* krlr, float* A, float* b, float* res) {
  (void)u;
  orc_so3_step(o->lastNextImage[L], o->nextImage[L], ib, kinv, krlr, o->W >> L, o->H >> L, A, b,
               res);
}
static void cpu_rgb_residual(void* u, OrcOdometry* o, int i, float minScale, const float* kt,
                             const float* krkinv, int* sigma, int* count) {
  (void)u;
  orc_rgb_residual(minScale, o->dIdx[i], o->dIdy[i], o->lastDepth[i], o->nextDepth[i],
                   o->lastImage[i], o->nextImage[i], o->corres[i], o->maxDepthDeltaRGB, kt, krkinv,
                   o->W >> i, o->H >> i, sigma, count);
}
static void cpu_icp(void* u, OrcOdometry* o, int i, const float* Rcurr, const float* tcurr,
                    const float* Rprev_inv, const float* tprev, float* A, float* b, float* res,
                    float* error_map) {
  (void)u;
  int div = 1 << i;
  orc_icp_step(Rcurr, tcurr, o->vmaps_curr[i], o->nmaps_curr[i], Rprev_inv, tprev, o->fx / div,
               o->fy / div, o->cx / div, o->cy / div, o->vmaps_g_prev[i], o->nmaps_g_prev[i],
               o->distThres, o->angleThres, o->W >> i, o->H >> i, A, b, res, error_map);
}
static void cpu_rgb_step(void* u, OrcOdometry* o, int i, float sigma, float* A, float* b) {
  (void)u;
  int div = 1 << i;
  orc_rgb_step(o->corres[i], sigma, o->cloud[i], o->fx / div, o->fy / div, o->dIdx[i], o->dIdy[i],
               o->sobelScale, o->W >> i, o->H >> i, A, b);
}
static const OrcStepBackend kCpuBackend = {0, 0, cpu_so3, cpu_rgb_residual, cpu_icp, cpu_rgb_step, 0};

void orc_odom_track(OrcOdometry* o, float trans[3], float rot[9], int rgbOnly, float icpWeight,
                    int pyramid, int fastOdom, int so3, float* icp_error_map, OrcTrackStats* st) {
  orc_odom_track_ex(o, trans, rot, rgbOnly, icpWeight, pyramid, fastOdom, so3, icp_error_map, st,
                    &kCpuBackend);
}

void orc_odom_dims(const OrcOdometry* o, int* W, int* H, float intr[4]) {
  *W = o->W;
  *H = o->H;
  intr[0] = o->fx;
  intr[1] = o->fy;
  intr[2] = o->cx;
  intr[3] = o->cy;
}

/* RGBDOdometry.cpp:217-477 (getIncrementalTransformation). */
void orc_odom_track_ex(OrcOdometry* o, float trans[3], float rot[9], int rgbOnly, float icpWeight,
                       int pyramid, int fastOdom, int so3, float* icp_error_map, OrcTrackStats* st,
                       const OrcStepBackend* be) {
  int icp = !rgbOnly && icpWeight > 0;
  int rgb = rgbOnly || icpWeight < 100;
  float Rprev[9], tprev[3], Rcurr[9], tcurr[3];
  memcpy(Rprev, rot, sizeof(Rprev));
  memcpy(tprev, trans, sizeof(tprev));
  memcpy(Rcurr, rot, sizeof(Rcurr));
  memcpy(tcurr, trans, sizeof(tcurr));
  OrcTrackStats s;
  memset(&s, 0, sizeof(s));

  if (rgb)
    for (int i = 0; i < NUM_PYRS; i++)
      orc_derivative_images(o->nextImage[i], o->W >> i, o->H >> i, o->dIdx[i], o->dIdy[i]);

  /* projectToPointCloud is issued per level inside the loop in the reference (:333); its inputs do
   * not change, so the oracle computes all levels up front (lets a device backend upload once). */
  if (rgb)
    for (int i = 0; i < NUM_PYRS; i++) {
      int div = 1 << i;
      orc_project_to_point_cloud(o->lastDepth[i], o->W >> i, o->H >> i, o->fx / div, o->fy / div,
                                 o->cx / div, o->cy / div, o->cloud[i]);
    }
  if (be->begin) be->begin(be->user, o);

  double resultR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};

  if (so3) { /* :239-310 */
    int L = 2;
    float R_lr[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double K[9], Kinv[9];
    K_of(o, L, K, Kinv);
    float lastError = FLT_MAX / 2, lastCount = FLT_MAX / 2;
    double lastResultR[9];
    memcpy(lastResultR, resultR, sizeof(resultR));
    for (int i = 0; i < 10; i++) {
      double tmp[9], homography[9], K_R_lr[9];
      mat3_mul_d(K, resultR, tmp);
      mat3_mul_d(tmp, Kinv, homography);
      mat3_mul_d(K, resultR, K_R_lr);
      float imageBasis[9], kinv[9], krlr[9], jtj[9], jtr[3], residual[2];
      for (int k = 0; k < 9; ++k) {
        imageBasis[k] = (float)homography[k];
        kinv[k] = (float)Kinv[k];
        krlr[k] = (float)K_R_lr[k];
      }
      be->so3_step(be->user, o, L, imageBasis, kinv, krlr, jtj, jtr, residual);
      s.so3_iterations++;
      s.lastSO3Error = sqrtf(residual[0]) / residual[1];
      s.lastSO3Count = residual[1];
      if (s.lastSO3Error < lastError && fabsf(lastError - s.lastSO3Count) < 0.001f) {
        break;
      } else if (s.lastSO3Error > lastError + 0.001f) {
        s.lastSO3Error = lastError;
        s.lastSO3Count = lastCount;
        memcpy(resultR, lastResultR, sizeof(resultR));
        break;
      }
      lastError = s.lastSO3Error;
      lastCount = s.lastSO3Count;
      memcpy(lastResultR, resultR, sizeof(resultR));
      /* Vector3f delta = jtj.ldlt().solve(jtr) -- f32 in Eigen; solved here in f64 from the f32
       * inputs and rounded to f32 */
      double Ad[9], bd[3], xd[3];
      for (int k = 0; k < 9; ++k) Ad[k] = jtj[k];
      for (int k = 0; k < 3; ++k) bd[k] = jtr[k];
      ldlt_solve_d(Ad, bd, 3, xd);
      double delta[3] = {(double)(float)xd[0], (double)(float)xd[1], (double)(float)xd[2]};
      double rotUpdate[9];
      rodrigues_d(delta, rotUpdate);
      float ru[9], nr[9];
      for (int k = 0; k < 9; ++k) ru[k] = (float)rotUpdate[k];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
          nr[r * 3 + c] = ru[r * 3 + 0] * R_lr[0 * 3 + c] + ru[r * 3 + 1] * R_lr[1 * 3 + c] +
                          ru[r * 3 + 2] * R_lr[2 * 3 + c];
      memcpy(R_lr, nr, sizeof(nr));
      for (int k = 0; k < 9; ++k) resultR[k] = R_lr[k];
    }
  }

  int iterations[NUM_PYRS];
  iterations[0] = fastOdom ? 3 : 10;
  iterations[1] = pyramid ? 5 : 0;
  iterations[2] = pyramid ? 4 : 0;

  float Rprev_inv[9];
  mat3_inverse_f(Rprev, Rprev_inv);

  double resultRt[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  if (so3)
    for (int x = 0; x < 3; x++)
      for (int y = 0; y < 3; y++) resultRt[x * 4 + y] = resultR[x * 3 + y];

  float lastRGBError = 0;
  for (int i = NUM_PYRS - 1; i >= 0; i--) {
    double K[9], Kinv[9];
    K_of(o, i, K, Kinv);
    lastRGBError = FLT_MAX;
    for (int j = 0; j < iterations[i]; j++) {
      double Rt[16];
      rigid_inverse_d(resultRt, Rt);
      double R[9] = {Rt[0], Rt[1], Rt[2], Rt[4], Rt[5], Rt[6], Rt[8], Rt[9], Rt[10]};
      double tmp[9], KRK_inv[9];
      mat3_mul_d(K, R, tmp);
      mat3_mul_d(tmp, Kinv, KRK_inv);
      float krkInv[9];
      for (int k = 0; k < 9; ++k) krkInv[k] = (float)KRK_inv[k];
      double Kt[3];
      for (int r = 0; r < 3; ++r)
        Kt[r] = K[r * 3 + 0] * Rt[3] + K[r * 3 + 1] * Rt[7] + K[r * 3 + 2] * Rt[11];
      float kt[3] = {(float)Kt[0], (float)Kt[1], (float)Kt[2]};

      int sigma = 0, rgbSize = 0;
      if (rgb) {
        float minScale = (float)(pow(o->minGrad[i], 2.0) / pow(o->sobelScale, 2.0));
        be->rgb_residual(be->user, o, i, minScale, kt, krkInv, &sigma, &rgbSize);
      }
      float tmpError = (float)(sqrt((double)sigma) / (double)rgbSize);
      float sigmaVal = (tmpError == 0) ? 1 : (float)rgbSize;
      if (rgbOnly && tmpError > lastRGBError) break;
      lastRGBError = tmpError;
      s.lastRGBError = tmpError;
      s.lastRGBCount = (float)rgbSize;
      if (rgbOnly) sigmaVal = -1;

      float A_icp[36], b_icp[6], residual[2] = {0, 0};
      memset(A_icp, 0, sizeof(A_icp));
      memset(b_icp, 0, sizeof(b_icp));
      if (icp) {
        be->icp_step(be->user, o, i, Rcurr, tcurr, Rprev_inv, tprev, A_icp, b_icp, residual,
                     (i == 0 && j == iterations[i] - 1) ? icp_error_map : NULL);
      }
      s.lastICPError = sqrtf(residual[0]) / residual[1];
      s.lastICPCount = residual[1];

      float A_rgbd[36], b_rgbd[6];
      memset(A_rgbd, 0, sizeof(A_rgbd));
      memset(b_rgbd, 0, sizeof(b_rgbd));
      if (rgb)
        be->rgb_step(be->user, o, i, sigmaVal, A_rgbd, b_rgbd);

      double result[6];
      if (icp && rgb) {
        double wgt = icpWeight;
        for (int k = 0; k < 36; ++k) s.lastA[k] = (double)A_rgbd[k] + wgt * wgt * (double)A_icp[k];
        for (int k = 0; k < 6; ++k) s.lastb[k] = (double)b_rgbd[k] + wgt * (double)b_icp[k];
      } else if (icp) {
        for (int k = 0; k < 36; ++k) s.lastA[k] = A_icp[k];
        for (int k = 0; k < 6; ++k) s.lastb[k] = b_icp[k];
      } else {
        for (int k = 0; k < 36; ++k) s.lastA[k] = A_rgbd[k];
        for (int k = 0; k < 6; ++k) s.lastb[k] = b_rgbd[k];
      }
      ldlt_solve_d(s.lastA, s.lastb, 6, result);

      /* OdometryProvider::computeUpdateSE3 (OdometryProvider.h:69-89) */
      double rvec[3] = {result[3], result[4], result[5]}, Rup[9];
      rodrigues_d(rvec, Rup);
      double Up[16] = {Rup[0], Rup[1], Rup[2], result[0], Rup[3], Rup[4], Rup[5], result[1],
                       Rup[6], Rup[7], Rup[8], result[2], 0,      0,      0,      1};
      mat4_mul_d(Up, resultRt, resultRt);
      float Ro[9], to[3];
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) Ro[r * 3 + c] = (float)resultRt[r * 4 + c];
        to[r] = (float)resultRt[r * 4 + 3];
      }
      /* currentT = [Rprev|tprev] * rgbOdom.inverse() (:452-460), Isometry inverse = transpose */
      float ti[3];
      for (int r = 0; r < 3; ++r)
        ti[r] = -(Ro[0 * 3 + r] * to[0] + Ro[1 * 3 + r] * to[1] + Ro[2 * 3 + r] * to[2]);
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c)
          Rcurr[r * 3 + c] = Rprev[r * 3 + 0] * Ro[c * 3 + 0] + Rprev[r * 3 + 1] * Ro[c * 3 + 1] +
                             Rprev[r * 3 + 2] * Ro[c * 3 + 2];
        tcurr[r] = Rprev[r * 3 + 0] * ti[0] + Rprev[r * 3 + 1] * ti[1] + Rprev[r * 3 + 2] * ti[2] +
                   tprev[r];
      }
    }
  }

  if (rgb) {
    float d[3] = {tcurr[0] - tprev[0], tcurr[1] - tprev[1], tcurr[2] - tprev[2]};
    if (sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) > 0.3f) {
      memcpy(Rcurr, Rprev, sizeof(Rcurr));
      memcpy(tcurr, tprev, sizeof(tcurr));
    }
  }
  if (be->end) be->end(be->user, o);
  if (so3)
    for (int i = 0; i < NUM_PYRS; i++) {
      uint8_t* t = o->lastNextImage[i];
      o->lastNextImage[i] = o->nextImage[i];
      o->nextImage[i] = t;
    }
  memcpy(trans, tcurr, sizeof(tcurr));
  memcpy(rot, Rcurr, sizeof(Rcurr));
  if (st) *st = s;
}
