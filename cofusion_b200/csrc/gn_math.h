// gn_math.h -- the FP64 Gauss-Newton host step of the tracker, written once for host AND device
// (the device-resident GN loop runs it in one thread of the finalising block, so a frame needs no
// host round trip; the generic host loop in odometry.cu calls the very same functions).
//
// Mirrors, without Eigen (not in this image; the module must not depend on it):
//   Core/Utils/RGBDOdometry.cpp:316        Rprev.inverse()            -> inverse3f (cofactors)
//   Core/Utils/RGBDOdometry.cpp:348-358    Rt^-1, K R K^-1, K t       -> pose_to_warp
//   Core/Utils/RGBDOdometry.cpp:425-435    lastA/lastb, ldlt().solve  -> combine_and_solve / ldlt_solve
//   Core/Utils/OdometryProvider.h:32-89    rodrigues, computeUpdateSE3-> rodrigues / update_se3
//   Core/Utils/RGBDOdometry.cpp:452-460    currentT = Tprev * odom^-1 -> compose_pose
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define CFB_HD __host__ __device__ __forceinline__
#else
#define CFB_HD inline
#endif

namespace cfb {
namespace gn {

// Solve A x = b for symmetric A (row-major n x n, n <= 6) by LDL^T with diagonal pivoting (largest
// remaining |d_ii| first -- the pivot rule of Eigen::LDLT, which the reference calls).
template <int N>
CFB_HD void ldlt_solve(const double* Ain, const double* bin, double* x) {
  double A[N * N], b[N];
  int perm[N];
  for (int i = 0; i < N * N; ++i) A[i] = Ain[i];
  for (int i = 0; i < N; ++i) {
    perm[i] = i;
    b[i] = bin[i];
  }
  for (int k = 0; k < N; ++k) {
    int p = k;
    double best = fabs(A[k * N + k]);
    for (int i = k + 1; i < N; ++i) {
      double v = fabs(A[i * N + i]);
      if (v > best) {
        best = v;
        p = i;
      }
    }
    if (p != k) {
      for (int j = 0; j < N; ++j) {
        double t = A[k * N + j];
        A[k * N + j] = A[p * N + j];
        A[p * N + j] = t;
      }
      for (int i = 0; i < N; ++i) {
        double t = A[i * N + k];
        A[i * N + k] = A[i * N + p];
        A[i * N + p] = t;
      }
      int ti = perm[k];
      perm[k] = perm[p];
      perm[p] = ti;
      double tb = b[k];
      b[k] = b[p];
      b[p] = tb;
    }
    double d = A[k * N + k];
    if (d == 0.0) continue;
    double inv = 1.0 / d;
    for (int i = k + 1; i < N; ++i) {
      double l = A[i * N + k] * inv;
      for (int j = k + 1; j < N; ++j) A[i * N + j] -= l * A[k * N + j];
      A[i * N + k] = l;
    }
  }
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < i; ++j) b[i] -= A[i * N + j] * b[j];
  for (int i = 0; i < N; ++i) b[i] = (A[i * N + i] != 0.0) ? b[i] / A[i * N + i] : 0.0;
  for (int i = N - 1; i >= 0; --i)
    for (int j = i + 1; j < N; ++j) b[i] -= A[j * N + i] * b[j];
  for (int i = 0; i < N; ++i) x[perm[i]] = b[i];
}

// Same solve without pivoting, every loop bound a compile-time constant so that the whole
// factorisation lives in registers when run by a single GPU thread (the pivoted version indexes
// through perm[] and spills to local memory: ~10 us per call on B200, measured).  For the symmetric
// positive (semi-)definite normal equations of the tracker both give the same solution to FP64
// rounding; zero pivots (no inliers) yield zeros like the pivoted version.
template <int N>
CFB_HD void ldlt_solve_unrolled(const double* Ain, const double* bin, double* x) {
  double A[N * N], b[N];
#pragma unroll
  for (int i = 0; i < N * N; ++i) A[i] = Ain[i];
#pragma unroll
  for (int i = 0; i < N; ++i) b[i] = bin[i];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const double d = A[k * N + k];
    const double inv = (d != 0.0) ? 1.0 / d : 0.0;
#pragma unroll
    for (int i = k + 1; i < N; ++i) {
      const double l = A[i * N + k] * inv;
#pragma unroll
      for (int j = k + 1; j < N; ++j) A[i * N + j] -= l * A[k * N + j];
      A[i * N + k] = l;
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < i; ++j) b[i] -= A[i * N + j] * b[j];
#pragma unroll
  for (int i = 0; i < N; ++i) b[i] = (A[i * N + i] != 0.0) ? b[i] / A[i * N + i] : 0.0;
#pragma unroll
  for (int i = N - 1; i >= 0; --i)
#pragma unroll
    for (int j = i + 1; j < N; ++j) b[i] -= A[j * N + i] * b[j];
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = b[i];
}

CFB_HD void rodrigues(const double r[3], double R[9]) {
  double rx = r[0], ry = r[1], rz = r[2];
  double theta = sqrt(rx * rx + ry * ry + rz * rz);
  R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
  if (theta >= 2.2204460492503131e-16) {
    double c = cos(theta), s = sin(theta), c1 = 1.0 - c, it = 1.0 / theta;
    rx *= it; ry *= it; rz *= it;
    R[0] = c + c1 * rx * rx;        R[1] = c1 * rx * ry - s * rz;  R[2] = c1 * rx * rz + s * ry;
    R[3] = c1 * rx * ry + s * rz;   R[4] = c + c1 * ry * ry;       R[5] = c1 * ry * rz - s * rx;
    R[6] = c1 * rx * rz - s * ry;   R[7] = c1 * ry * rz + s * rx;  R[8] = c + c1 * rz * rz;
  }
}

CFB_HD void mul3(const double* a, const double* b, double* c) {
  double r[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      r[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
#pragma unroll
  for (int i = 0; i < 9; ++i) c[i] = r[i];
}

// f32 3x3 inverse by cofactors
CFB_HD void inverse3f(const float* m, float* inv) {
  float c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8],
        c02 = m[3] * m[7] - m[4] * m[6];
  float id = 1.0f / (m[0] * c00 + m[1] * c01 + m[2] * c02);
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

// K for a pyramid level from f32 level intrinsics (CameraModel::operator(), types.cuh:94-98)
CFB_HD void make_K(float fx, float fy, float cx, float cy, double K[9], double Kinv[9]) {
  K[0] = fx; K[1] = 0; K[2] = cx; K[3] = 0; K[4] = fy; K[5] = cy; K[6] = 0; K[7] = 0; K[8] = 1;
  Kinv[0] = 1.0 / fx; Kinv[1] = 0; Kinv[2] = -(double)cx / fx;
  Kinv[3] = 0; Kinv[4] = 1.0 / fy; Kinv[5] = -(double)cy / fy;
  Kinv[6] = 0; Kinv[7] = 0; Kinv[8] = 1;
}

// resultRt (4x4 rigid, row-major f64) -> krkinv = K R' K^-1, kt = K t' with [R'|t'] = resultRt^-1
CFB_HD void pose_to_warp(const double* resultRt, const double* K, const double* Kinv, float krkinv[9],
                         float kt[3]) {
  double R[9], t[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) R[i * 3 + j] = resultRt[j * 4 + i];
#pragma unroll
  for (int i = 0; i < 3; ++i)
    t[i] = -(R[i * 3] * resultRt[3] + R[i * 3 + 1] * resultRt[7] + R[i * 3 + 2] * resultRt[11]);
  double tmp[9], krk[9];
  mul3(K, R, tmp);
  mul3(tmp, Kinv, krk);
#pragma unroll
  for (int i = 0; i < 9; ++i) krkinv[i] = (float)krk[i];
#pragma unroll
  for (int i = 0; i < 3; ++i)
    kt[i] = (float)(K[i * 3] * t[0] + K[i * 3 + 1] * t[1] + K[i * 3 + 2] * t[2]);
}

// resultRt <- [exp(x[3..5]) | x[0..2]] * resultRt
CFB_HD void update_se3(double* resultRt, const double x[6]) {
  double R[9];
  const double rv[3] = {x[3], x[4], x[5]};
  rodrigues(rv, R);
  double U[16] = {R[0], R[1], R[2], x[0], R[3], R[4], R[5], x[1], R[6], R[7], R[8], x[2], 0, 0, 0, 1};
  double r[16];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      r[i * 4 + j] = U[i * 4] * resultRt[j] + U[i * 4 + 1] * resultRt[4 + j] +
                     U[i * 4 + 2] * resultRt[8 + j] + U[i * 4 + 3] * resultRt[12 + j];
#pragma unroll
  for (int i = 0; i < 16; ++i) resultRt[i] = r[i];
}

// [Rcurr|tcurr] = [Rprev|tprev] * (f32 cast of resultRt)^-1, rotation inverse = transpose
CFB_HD void compose_pose(const float* Rprev, const float* tprev, const double* resultRt, float* Rcurr,
                         float* tcurr) {
  float Ro[9], to[3], ti[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) Ro[r * 3 + c] = (float)resultRt[r * 4 + c];
    to[r] = (float)resultRt[r * 4 + 3];
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) ti[r] = -(Ro[r] * to[0] + Ro[3 + r] * to[1] + Ro[6 + r] * to[2]);
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
      Rcurr[r * 3 + c] =
          Rprev[r * 3] * Ro[c * 3] + Rprev[r * 3 + 1] * Ro[c * 3 + 1] + Rprev[r * 3 + 2] * Ro[c * 3 + 2];
    tcurr[r] = Rprev[r * 3] * ti[0] + Rprev[r * 3 + 1] * ti[1] + Rprev[r * 3 + 2] * ti[2] + tprev[r];
  }
}

// Unpack the 29 packed sums (order aa..ag, bb..bg, ..., ff, fg, residual, inliers; types.cuh:101-112)
// into A (6x6 row-major, symmetric), b (6) as icpStep's host tail does (reduce.cu:484-498).
template <class T>
CFB_HD void unpack_se3(const float* packed, T* A, T* b) {
  int shift = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = i; j < 7; ++j) {
      T v = (T)packed[shift++];
      if (j == 6)
        b[i] = v;
      else
        A[j * 6 + i] = A[i * 6 + j] = v;
    }
}
template <class T>
CFB_HD void unpack_so3(const float* packed, T* A, T* b) {
  int shift = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = i; j < 4; ++j) {
      T v = (T)packed[shift++];
      if (j == 3)
        b[i] = v;
      else
        A[j * 3 + i] = A[i * 3 + j] = v;
    }
}

}  // namespace gn
}  // namespace cfb
