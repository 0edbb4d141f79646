// gn_serial.cuh -- the serial (one-thread) pieces of the device-resident Gauss-Newton loop and the
// per-pixel photometric helpers, shared by the multi-kernel graph path (gn_device.cu) and the
// persistent cooperative kernel (gn_persistent.cu).  Reference: Core/Utils/RGBDOdometry.cpp:217-477.
#pragma once
#include <float.h>

#include "gn_math.h"
#include "odometry.cuh"
#include "tracker_device.cuh"

namespace cfb {
namespace dev {

struct LevelK {  // f32 level intrinsics (CameraModel::operator())
  float fx, fy, cx, cy;
};

__device__ __forceinline__ void so3_matrices(GNState* g, LevelK k) {
  double K[9], Kinv[9], KR[9], H[9];
  gn::make_K(k.fx, k.fy, k.cx, k.cy, K, Kinv);
  gn::mul3(K, g->resultR, KR);
  gn::mul3(KR, Kinv, H);
  for (int q = 0; q < 9; ++q) {
    g->so3_imageBasis.m[q] = (float)H[q];
    g->so3_kinv.m[q] = (float)Kinv[q];
    g->so3_krlr.m[q] = (float)KR[q];
  }
}


// reset state for a new frame (RGBDOdometry.cpp:224-255, :316-318)
__device__ __forceinline__ void gn_init_serial(GNState* g, StepScratch* sc, const float* __restrict__ pose_in,
                                               LevelK k_so3) {
  for (int q = 0; q < 9; ++q) {
    g->Rprev[q] = pose_in[3 + q];
    g->pose.Rcurr.m[q] = pose_in[3 + q];
    g->resultR[q] = g->lastResultR[q] = (q % 4 == 0) ? 1.0 : 0.0;
    g->R_lr[q] = (q % 4 == 0) ? 1.f : 0.f;
  }
  for (int q = 0; q < 3; ++q) g->pose.tprev[q] = g->pose.tcurr[q] = pose_in[q];
  for (int q = 0; q < 12; ++q) g->out_trans[q] = pose_in[q];  // out_trans[3] + out_rot[9] are contiguous
  gn::inverse3f(g->Rprev, g->pose.Rprev_inv.m);
  g->so3_lastError = FLT_MAX / 2;
  g->so3_lastCount = FLT_MAX / 2;
  g->so3_done = 0;
  TrackStats z = {};
  g->stats = z;
  so3_matrices(g, k_so3);
  if (sc) {
    sc->rgb_count = 0;
    sc->rgb_sigma = 0;
  }
}

// host logic of one SO(3) iteration after the reduction (RGBDOdometry.cpp:281-308); sets so3_done
__device__ __forceinline__ void so3_update_serial(GNState* g, const float* out32, LevelK k) {
  float jtj[9], jtr[3];
  gn::unpack_so3(out32, jtj, jtr);
  TrackStats& st = g->stats;
  st.so3_iterations++;
  st.lastSO3Error = sqrtf(out32[9]) / out32[10];
  st.lastSO3Count = out32[10];
  if (st.lastSO3Error < g->so3_lastError && fabsf(g->so3_lastError - st.lastSO3Count) < 0.001f) {
    g->so3_done = 1;
    return;
  } else if (st.lastSO3Error > g->so3_lastError + 0.001f) {
    st.lastSO3Error = g->so3_lastError;
    st.lastSO3Count = g->so3_lastCount;
    for (int q = 0; q < 9; ++q) g->resultR[q] = g->lastResultR[q];
    g->so3_done = 1;
    return;
  }
  g->so3_lastError = st.lastSO3Error;
  g->so3_lastCount = st.lastSO3Count;
  for (int q = 0; q < 9; ++q) g->lastResultR[q] = g->resultR[q];
  double Ad[9], bd[3], xd[3];
  for (int q = 0; q < 9; ++q) Ad[q] = jtj[q];
  for (int q = 0; q < 3; ++q) bd[q] = jtr[q];
  gn::ldlt_solve_unrolled<3>(Ad, bd, xd);
  double delta[3] = {(double)(float)xd[0], (double)(float)xd[1], (double)(float)xd[2]};
  double rotUpdate[9];
  gn::rodrigues(delta, rotUpdate);
  float ru[9], nr[9];
  for (int q = 0; q < 9; ++q) ru[q] = (float)rotUpdate[q];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      nr[r * 3 + c] = ru[r * 3] * g->R_lr[c] + ru[r * 3 + 1] * g->R_lr[3 + c] + ru[r * 3 + 2] * g->R_lr[6 + c];
  for (int q = 0; q < 9; ++q) {
    g->R_lr[q] = nr[q];
    g->resultR[q] = nr[q];
  }
  so3_matrices(g, k);
}

// seed resultRt with the SO(3) result, first warp (RGBDOdometry.cpp:320-328)
__device__ __forceinline__ void gn_begin_serial(GNState* g, int use_so3, LevelK k_first) {
  for (int q = 0; q < 16; ++q) g->resultRt[q] = (q % 5 == 0) ? 1.0 : 0.0;
  if (use_so3)
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) g->resultRt[r * 4 + c] = g->resultR[r * 3 + c];
  double K[9], Kinv[9];
  gn::make_K(k_first.fx, k_first.fy, k_first.cx, k_first.cy, K, Kinv);
  gn::pose_to_warp(g->resultRt, K, Kinv, g->warp.krkinv.m, g->warp.kt);
}

// RGBDOdometry.cpp:373-374
__device__ __forceinline__ float rgb_sigma_from_counts(int cnt, int sg, float* tmpErrorOut) {
  const float tmpError = (float)(sqrt((double)sg) / (double)cnt);
  if (tmpErrorOut) *tmpErrorOut = tmpError;
  return (tmpError == 0.f) ? 1.f : (float)cnt;
}

// ---- frame-side photometric preparation, once per level per frame: gradient images
// (cudafuncs.cu:658-683) + every iteration-invariant gate of RGBResidual::getProducts folded into one
// byte per pixel: j0 < W-5, i < H-1 (reduce.cu:799), 4x4 window of nextImage > 0 (:803-814),
// gradient magnitude gate (:823-825), nextDepth not NaN (:832).
__constant__ float c_sx[9] = {0.52201f, 0.00000f, -0.52201f, 0.79451f, -0.00000f, -0.79451f, 0.52201f, 0.00000f, -0.52201f};
__constant__ float c_sy[9] = {0.52201f, 0.79451f, 0.52201f, 0.00000f, 0.00000f, 0.00000f, -0.52201f, -0.79451f, -0.52201f};
__device__ __forceinline__ void rgb_prepare_pixel(const unsigned char* __restrict__ img, int W, int H,
                                                  const float* __restrict__ nextDepth, float minScale,
                                                  short* __restrict__ dx, short* __restrict__ dy,
                                                  unsigned char* __restrict__ cand, int x, int y) {
  float dxVal = 0.f, dyVal = 0.f;
  int k = 8;
  for (int j = max(y - 1, 0); j <= min(y + 1, H - 1); j++)
    for (int i = max(x - 1, 0); i <= min(x + 1, W - 1); i++) {
      float p = (float)__ldg(img + j * W + i);
      dxVal = __fadd_rn(dxVal, __fmul_rn(p, c_sx[k]));  // no FMA contraction: bit-identical to
      dyVal = __fadd_rn(dyVal, __fmul_rn(p, c_sy[k]));  // computeDerivativeImages in image_kernels.cu
      --k;
    }
  const short sx = (short)dxVal, sy = (short)dyVal;
  dx[y * W + x] = sx;
  dy[y * W + x] = sy;
  unsigned ok = (x < W - 5 && y < H - 1) ? 1u : 0u;
  for (int u = max(y - 2, 0); u < min(y + 2, H); u++)
    for (int v = max(x - 2, 0); v < min(x + 2, W); v++) ok &= (unsigned)(__ldg(img + u * W + v) > 0);
  const float mTwo = (float)((sx * sx) + (sy * sy));
  ok &= (unsigned)(mTwo >= minScale);
  ok &= (unsigned)(!isnan(__ldg(nextDepth + y * W + x)));
  cand[y * W + x] = (unsigned char)ok;
}

// RGBResidual::getProducts for a pixel that already passed the invariant gates (reduce.cu:827-853)
__device__ __forceinline__ bool rgb_residual_cand(const RgbResidualArgs& a, const RgbWarp& Wp, int x, int y,
                                                  DataTerm& corres, int& sq) {
  float d1 = __ldg(row_ptr(a.nextDepth, a.depth_pitch, y) + x);
  const float* k = Wp.krkinv.m;
  float transformed_d1 = d1 * (k[6] * x + k[7] * y + k[8]) + Wp.kt[2];
  int u0 = __float2int_rn((d1 * (k[0] * x + k[1] * y + k[2]) + Wp.kt[0]) / transformed_d1);
  int v0 = __float2int_rn((d1 * (k[3] * x + k[4] * y + k[5]) + Wp.kt[1]) / transformed_d1);
  if (!(u0 >= 0 && v0 >= 0 && u0 < a.cols && v0 < a.rows)) return false;
  float d0 = __ldg(row_ptr(a.lastDepth, a.depth_pitch, v0) + u0);
  unsigned char li = __ldg(row_ptr(a.lastImage, a.img_pitch, v0) + u0);
  if (!(d0 > 0 && fabsf(transformed_d1 - d0) <= a.maxDepthDelta && li != 0)) return false;
  corres.zero = make_short2((short)u0, (short)v0);
  corres.one = make_short2((short)x, (short)y);
  corres.diff = (float)__ldg(row_ptr(a.nextImage, a.img_pitch, y) + x) - (float)li;
  corres.valid = true;
  sq = (int)(corres.diff * corres.diff);
  return true;
}

// RGBDOdometry.cpp:412-460 in one thread, FP64: combine ICP + RGB normal equations, LDLT, SE(3)
// update, new pose, next warp (or the final pose with the 0.3 m sanity reset, :464-467).
__device__ __forceinline__ void gn_solve_serial(GNState* g, StepScratch* sc, const float* out32, float icpWeight,
                                                LevelK k_next, int is_last, float tmpError, int cnt) {
  TrackStats& st = g->stats;
  st.lastRGBError = tmpError;
  st.lastRGBCount = (float)cnt;
  const float* icp = g->icp_result;
  st.lastICPError = sqrtf(icp[27]) / icp[28];
  st.lastICPCount = icp[28];
  double A_icp[36], b_icp[6], A_rgb[36], b_rgb[6];
  gn::unpack_se3(icp, A_icp, b_icp);
  gn::unpack_se3(out32, A_rgb, b_rgb);
  const double w = icpWeight;
  double A[36], bb[6], x[6];
#pragma unroll
  for (int q = 0; q < 36; ++q) A[q] = A_rgb[q] + w * w * A_icp[q];
#pragma unroll
  for (int q = 0; q < 6; ++q) bb[q] = b_rgb[q] + w * b_icp[q];
  if (is_last || sc) {  // lastA / lastb are only reported for the final iteration (RGBDOdometry.h:62-70);
#pragma unroll         // the multi-kernel path (sc != nullptr) keeps them current for the per-step tests
    for (int q = 0; q < 36; ++q) st.lastA[q] = A[q];
#pragma unroll
    for (int q = 0; q < 6; ++q) st.lastb[q] = bb[q];
  }
  gn::ldlt_solve_unrolled<6>(A, bb, x);
  double Rt[16];
  float Rprev[9], tprev[3], Rcurr[9], tcurr[3];
#pragma unroll
  for (int q = 0; q < 16; ++q) Rt[q] = g->resultRt[q];
#pragma unroll
  for (int q = 0; q < 9; ++q) Rprev[q] = g->Rprev[q];
#pragma unroll
  for (int q = 0; q < 3; ++q) tprev[q] = g->pose.tprev[q];
  gn::update_se3(Rt, x);
  gn::compose_pose(Rprev, tprev, Rt, Rcurr, tcurr);
#pragma unroll
  for (int q = 0; q < 16; ++q) g->resultRt[q] = Rt[q];
#pragma unroll
  for (int q = 0; q < 9; ++q) g->pose.Rcurr.m[q] = Rcurr[q];
#pragma unroll
  for (int q = 0; q < 3; ++q) g->pose.tcurr[q] = tcurr[q];
  if (sc) {
    sc->rgb_count = 0;
    sc->rgb_sigma = 0;
  }
  if (!is_last) {
    double K[9], Kinv[9];
    float krk[9], kt[3];
    gn::make_K(k_next.fx, k_next.fy, k_next.cx, k_next.cy, K, Kinv);
    gn::pose_to_warp(Rt, K, Kinv, krk, kt);
#pragma unroll
    for (int q = 0; q < 9; ++q) g->warp.krkinv.m[q] = krk[q];
#pragma unroll
    for (int q = 0; q < 3; ++q) g->warp.kt[q] = kt[q];
  } else {
    // RGBDOdometry.cpp:464-467: photometric sanity reset
    float d0 = tcurr[0] - tprev[0], d1 = tcurr[1] - tprev[1], d2 = tcurr[2] - tprev[2];
    bool reset = sqrtf(d0 * d0 + d1 * d1 + d2 * d2) > 0.3f;
#pragma unroll
    for (int q = 0; q < 9; ++q) g->out_rot[q] = reset ? Rprev[q] : Rcurr[q];
#pragma unroll
    for (int q = 0; q < 3; ++q) g->out_trans[q] = reset ? tprev[q] : tcurr[q];
  }
}

}  // namespace dev
}  // namespace cfb
