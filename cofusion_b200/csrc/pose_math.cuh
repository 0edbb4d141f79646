// pose_math.cuh -- the small f32 pose algebra of Model (inverse, fusion weight) written once for the host
// and the device.  Every float operation is an explicit round-to-nearest intrinsic on the device (never
// contracted into an FMA, whatever the flags of the including translation unit) and a plain operation in
// the host build (no FMA instructions there), so both sides -- and the CPU oracle, which states the same
// expressions -- produce the same bits.  This is what lets a frame run without a host round trip: the
// tracker kernel derives everything the fuse / clean / predict kernels need from the new pose on the device.
#pragma once
#include <math.h>

#include "detmath.cuh"
#include "surfel_kernels.cuh"

namespace cfb {

#ifdef __CUDA_ARCH__
#define CFB_FMUL(a, b) __fmul_rn((a), (b))
#define CFB_FADD(a, b) __fadd_rn((a), (b))
#define CFB_FDIV(a, b) __fdiv_rn((a), (b))
#define CFB_FSQRT(a) __fsqrt_rn((a))
#else
#define CFB_FMUL(a, b) ((a) * (b))
#define CFB_FADD(a, b) ((a) + (b))
#define CFB_FDIV(a, b) ((a) / (b))
#define CFB_FSQRT(a) sqrtf((a))
#endif

// Eigen inverse of a rigid 4x4 (row-major, 16 floats): R^T, -R^T t
__host__ __device__ inline void pose_inverse16(const float* T, float* Ti) {
  for (int i = 0; i < 16; ++i) Ti[i] = 0.f;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) Ti[r * 4 + c] = T[c * 4 + r];
  for (int r = 0; r < 3; ++r)
    Ti[r * 4 + 3] = -CFB_FADD(CFB_FADD(CFB_FMUL(Ti[r * 4 + 0], T[3]), CFB_FMUL(Ti[r * 4 + 1], T[7])), CFB_FMUL(Ti[r * 4 + 2], T[11]));
  Ti[15] = 1.f;
}

// Model::computeFusionWeight (Model.cpp:391-406) with multiplier 1: weight from max(|t|, |log R|) of
// diff = pose^-1 * lastPose (rodrigues2, Model.cpp:816-857, without the SVD re-orthogonalisation)
__host__ __device__ inline float fusion_weight_base(const float* pose, const float* lastPose) {
  float pinv[16], d[16];
  pose_inverse16(pose, pinv);
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      float s = 0;
      for (int k = 0; k < 4; ++k) s = CFB_FADD(s, CFB_FMUL(pinv[r * 4 + k], lastPose[k * 4 + c]));
      d[r * 4 + c] = s;
    }
  const float tn = CFB_FSQRT(CFB_FADD(CFB_FADD(CFB_FMUL(d[3], d[3]), CFB_FMUL(d[7], d[7])), CFB_FMUL(d[11], d[11])));
  double rx = (double)CFB_FADD(d[9], -d[6]), ry = (double)CFB_FADD(d[2], -d[8]), rz = (double)CFB_FADD(d[4], -d[1]);
  double s = sqrt(CFB_DMUL(CFB_DADD(CFB_DADD(CFB_DMUL(rx, rx), CFB_DMUL(ry, ry)), CFB_DMUL(rz, rz)), 0.25));
  double c = CFB_DMUL(CFB_DADD(CFB_DADD(CFB_DADD((double)d[0], (double)d[5]), (double)d[10]), -1.0), 0.5);
  c = c > 1. ? 1. : c < -1. ? -1. : c;
  double theta = det_acos(c);
  if (s < 1e-5) {
    if (c > 0) {
      rx = ry = rz = 0;
    } else {
      double t = CFB_DMUL((double)CFB_FADD(d[0], 1.f), 0.5);
      rx = sqrt(t > 0 ? t : 0);
      t = CFB_DMUL((double)CFB_FADD(d[5], 1.f), 0.5);
      ry = CFB_DMUL(sqrt(t > 0 ? t : 0), (d[1] < 0 ? -1.0 : 1.0));
      t = CFB_DMUL((double)CFB_FADD(d[10], 1.f), 0.5);
      rz = CFB_DMUL(sqrt(t > 0 ? t : 0), (d[2] < 0 ? -1.0 : 1.0));
      if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && (d[6] > 0) != (CFB_DMUL(ry, rz) > 0)) rz = -rz;
      theta = theta / sqrt(CFB_DADD(CFB_DADD(CFB_DMUL(rx, rx), CFB_DMUL(ry, ry)), CFB_DMUL(rz, rz)));
      rx = CFB_DMUL(rx, theta);
      ry = CFB_DMUL(ry, theta);
      rz = CFB_DMUL(rz, theta);
    }
  } else {
    const double vth = CFB_DMUL(1.0 / CFB_DMUL(2.0, s), theta);
    rx = CFB_DMUL(rx, vth);
    ry = CFB_DMUL(ry, vth);
    rz = CFB_DMUL(rz, vth);
  }
  const float fx = (float)rx, fy = (float)ry, fz = (float)rz;
  const float rn = CFB_FSQRT(CFB_FADD(CFB_FADD(CFB_FMUL(fx, fx), CFB_FMUL(fy, fy)), CFB_FMUL(fz, fz)));
  float weighting = tn > rn ? tn : rn;
  const float largest = 0.01f, minWeight = 0.5f;
  if (weighting > largest) weighting = largest;
  const float w = CFB_FADD(1.0f, -CFB_FDIV(weighting, largest));
  return w > minWeight ? w : minWeight;
}

// refresh a model's device pose block from the tracker output (t[3], R[9]): last <- pose, pose <- new
__host__ __device__ inline void pose_block_update(PoseDev* pd, const float* trans, const float* rot) {
  float oldp[16], newp[16], inv[16];
  for (int i = 0; i < 12; ++i) oldp[i] = pd->pose.m[i];
  oldp[12] = oldp[13] = oldp[14] = 0.f;
  oldp[15] = 1.f;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) newp[r * 4 + c] = rot[r * 3 + c];
    newp[r * 4 + 3] = trans[r];
  }
  newp[12] = newp[13] = newp[14] = 0.f;
  newp[15] = 1.f;
  pose_inverse16(newp, inv);
  const float wb = fusion_weight_base(newp, oldp);
  for (int i = 0; i < 12; ++i) {
    pd->last.m[i] = oldp[i];
    pd->pose.m[i] = newp[i];
    pd->inv.m[i] = inv[i];
  }
  for (int i = 0; i < 3; ++i) pd->tr[i] = trans[i];
  for (int i = 0; i < 9; ++i) pd->tr[3 + i] = rot[i];
  pd->weightBase = wb;
}

}  // namespace cfb
