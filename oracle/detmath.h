/* detmath.h -- oracle-side copy of the deterministic f32 exp (TEST INFRASTRUCTURE, see cf_oracle.h).
 * GLSL exp() has implementation-defined precision (depth_bilateral_metric.frag:64, surfels.glsl:45),
 * so the oracle freezes one polynomial realisation; only +,-,*,rint and EXPLICIT fused multiply-adds
 * (C99 fmaf, one rounding) are used, which a GPU kernel reproduces bit for bit with FFMA. */
#ifndef ORC_DETMATH_H_
#define ORC_DETMATH_H_
#include <math.h>
#include <stdint.h>
/* Functions that evaluate orc_expf in a hot loop carry ORC_FMA_CLONES: GCC emits one clone with the
 * FMA instruction inlined (picked at load time on any CPU that has it) and a portable clone that calls
 * libm's fmaf -- both IEEE-exact, so the result does not depend on the host. */
#if defined(__GNUC__) && defined(__x86_64__) && !defined(__clang__)
#define ORC_FMA_CLONES __attribute__((target_clones("fma", "default")))
#else
#define ORC_FMA_CLONES
#endif
static inline float orc_expf(float x) {
  if (!(x >= -87.0f)) return (x != x) ? x : 0.0f;
  if (x > 88.0f) x = 88.0f;
  const float n = rintf(x * 1.44269504088896341f);
  float r = fmaf(n, -0.693359375f, x);
  r = fmaf(n, 2.12194440e-4f, r);
  float p = 1.9875691500E-4f;
  p = fmaf(p, r, 1.3981999507E-3f);
  p = fmaf(p, r, 8.3334519073E-3f);
  p = fmaf(p, r, 4.1665795894E-2f);
  p = fmaf(p, r, 1.6666665459E-1f);
  p = fmaf(p, r, 5.0000001201E-1f);
  const float y = fmaf(p, r * r, r) + 1.0f;
  int e = (int)n;
  union { uint32_t u; float f; } s;
  s.u = (uint32_t)(e + 127) << 23;
  return y * s.f;
}
/* Deterministic double acos (Cephes asin/acos rational approximations, |err| <= 2 ulp): only IEEE + - * /
 * and sqrt in a fixed order without contraction, so the CPU oracle and a GPU kernel agree bit for bit.
 * Model::computeFusionWeight (Model.cpp:391-406) takes the rotation angle through acos. */
static inline double orc_polevl(double x, const double* c, int n) {
  double a = c[0];
  for (int i = 1; i <= n; ++i) a = a * x + c[i];
  return a;
}
static inline double orc_p1evl(double x, const double* c, int n) {
  double a = x + c[0];
  for (int i = 1; i < n; ++i) a = a * x + c[i];
  return a;
}
static inline double orc_asin(double x) {
  static const double P[6] = {4.253011369004428248960E-3, -6.019598008014123785661E-1, 5.444622390564711410273E0,
                              -1.626247967210700244449E1, 1.956261983317594739197E1, -8.198089802484824371615E0};
  static const double Q[5] = {-1.474091372988853791896E1, 7.049610280856842141659E1, -1.471791292232726029859E2,
                              1.395105614657485689735E2, -4.918853881490881290097E1};
  static const double R[5] = {2.967721961301243206100E-3, -5.634242780008963776856E-1, 6.968710824104713396794E0,
                              -2.556901049652824852289E1, 2.853665548261061424989E1};
  static const double S[4] = {-2.194779531642920639778E1, 1.470656354026814941758E2, -3.838770957603691357202E2,
                              3.424398657913078477438E2};
  const double PIO4 = 7.85398163397448309616E-1, MOREBITS = 6.123233995736765886130E-17;
  const double a = x < 0 ? -x : x;
  double z;
  if (a > 0.625) {
    double zz = 1.0 - a;
    const double p = zz * orc_polevl(zz, R, 4) / orc_p1evl(zz, S, 4);
    zz = sqrt(zz + zz);
    z = PIO4 - zz;
    zz = zz * p - MOREBITS;
    z = z - zz;
    z = z + PIO4;
  } else {
    if (a < 1.0e-8) return x;
    const double zz = a * a;
    z = zz * orc_polevl(zz, P, 5) / orc_p1evl(zz, Q, 5);
    z = a * z + a;
  }
  return x < 0 ? -z : z;
}
static inline double orc_acos(double x) { /* x in [-1, 1] */
  const double PIO4 = 7.85398163397448309616E-1, MOREBITS = 6.123233995736765886130E-17;
  if (x > 0.5) return 2.0 * orc_asin(sqrt(0.5 - 0.5 * x));
  double z = PIO4 - orc_asin(x);
  z = z + MOREBITS;
  z = z + PIO4;
  return z;
}
#endif
