// detmath.cuh -- deterministic f32 math used where results must be BIT-EXACT between the CUDA
// kernels and the CPU oracle (bilateral weights, surfel confidence): only +,-,*,/, fma and rint, each
// correctly rounded on both sides (this translation unit is compiled with -fmad=false and the
// default -prec-div=true -prec-sqrt=true -ftz=false; the oracle with -ffp-contract=off).
// GLSL leaves exp() precision implementation-defined (the reference runs it in a shader,
// depth_bilateral_metric.frag:64, surfels.glsl:45), so a fixed polynomial is a valid realisation.
#pragma once

namespace cfb {

// Cephes-style expf: |rel err| ~ 1 ulp on [-87, 88].  Every multiply-add is an EXPLICIT IEEE fused
// multiply-add (one rounding): identical on the device (FFMA, also under -fmad=false, which only stops
// implicit contraction) and in the oracle (C99 fmaf), at half the instructions of the unfused form --
// the 13x13 bilateral filter evaluates it 169 times per pixel and is bound by it.
__host__ __device__ __forceinline__ float det_expf(float x) {
  // branch free: the polynomial always runs on the clamped argument, the out-of-range / NaN cases are
  // selected at the end (same values as `if (!(x >= -87)) return x != x ? x : 0; if (x > 88) x = 88;`)
  const float x_in = x;
  x = fminf(fmaxf(x, -87.0f), 88.0f);
#ifdef __CUDA_ARCH__
  // rint and the float -> int conversion both issue on the quarter-rate conversion pipe, and this function runs 169
  // times per pixel of the bilateral filter.  Adding 1.5 * 2^23 rounds to the nearest integer, ties to even -- exactly
  // rintf for |v| < 2^22 -- and the integer is then the difference of the bit patterns.  Same n, same result.
  const float tmagic = __fadd_rn(__fmul_rn(x, 1.44269504088896341f), 12582912.0f);
  const float n = __fadd_rn(tmagic, -12582912.0f);
#else
  const float n = rintf(x * 1.44269504088896341f);
#endif
  float r = fmaf(n, -0.693359375f, x);
  r = fmaf(n, 2.12194440e-4f, r);
  float p = 1.9875691500E-4f;
  p = fmaf(p, r, 1.3981999507E-3f);
  p = fmaf(p, r, 8.3334519073E-3f);
  p = fmaf(p, r, 4.1665795894E-2f);
  p = fmaf(p, r, 1.6666665459E-1f);
  p = fmaf(p, r, 5.0000001201E-1f);
  const float y = fmaf(p, r * r, r) + 1.0f;
#ifdef __CUDA_ARCH__
  const int e = __float_as_int(tmagic) - 0x4B400000;  // in [-126, 127]
#else
  const int e = (int)n;  // in [-126, 127]
#endif
  union {
    unsigned u;
    float f;
  } s;
  s.u = (unsigned)(e + 127) << 23;
  const float v = y * s.f;
  return (x_in >= -87.0f) ? v : ((x_in != x_in) ? x_in : 0.0f);
}

#ifdef __CUDACC__
// Two det_expf at once on Blackwell's packed f32x2 pipe (fma.rn.f32x2 / mul / add: IEEE per lane, so each lane is
// bit-identical to det_expf): the polynomial, which is most of the work, costs half the issue slots.  For
// arguments <= 0 only (the bilateral weights): the upper clamp of det_expf can never bind there.
__device__ __forceinline__ float2 det_expf2_nonpos(float2 x) {
  const float2 x_in = x;
  x.x = fmaxf(x.x, -87.0f);
  x.y = fmaxf(x.y, -87.0f);
  // n = rint(x * log2 e) by the magic-number addition (see det_expf): packed adds instead of two conversions
  const float2 tm = make_float2(__fadd_rn(__fmul_rn(x.x, 1.44269504088896341f), 12582912.0f),
                                __fadd_rn(__fmul_rn(x.y, 1.44269504088896341f), 12582912.0f));
  const float2 n = make_float2(__fadd_rn(tm.x, -12582912.0f), __fadd_rn(tm.y, -12582912.0f));
  float2 r = __ffma2_rn(n, make_float2(-0.693359375f, -0.693359375f), x);
  r = __ffma2_rn(n, make_float2(2.12194440e-4f, 2.12194440e-4f), r);
  float2 p = make_float2(1.9875691500E-4f, 1.9875691500E-4f);
  p = __ffma2_rn(p, r, make_float2(1.3981999507E-3f, 1.3981999507E-3f));
  p = __ffma2_rn(p, r, make_float2(8.3334519073E-3f, 8.3334519073E-3f));
  p = __ffma2_rn(p, r, make_float2(4.1665795894E-2f, 4.1665795894E-2f));
  p = __ffma2_rn(p, r, make_float2(1.6666665459E-1f, 1.6666665459E-1f));
  p = __ffma2_rn(p, r, make_float2(5.0000001201E-1f, 5.0000001201E-1f));
  const float2 y = __fadd2_rn(__ffma2_rn(p, __fmul2_rn(r, r), r), make_float2(1.0f, 1.0f));
  float2 sc;
  sc.x = __int_as_float((__float_as_int(tm.x) - 0x4B400000 + 127) << 23);
  sc.y = __int_as_float((__float_as_int(tm.y) - 0x4B400000 + 127) << 23);
  float2 v = __fmul2_rn(y, sc);
  v.x = (x_in.x >= -87.0f) ? v.x : ((x_in.x != x_in.x) ? x_in.x : 0.0f);
  v.y = (x_in.y >= -87.0f) ? v.y : ((x_in.y != x_in.y) ? x_in.y : 0.0f);
  return v;
}
#endif

// Deterministic double acos (Cephes asin / acos rational approximations, |err| <= 2 ulp): only IEEE
// + - * / and sqrt in a fixed order, never contracted (explicit round-to-nearest intrinsics on the device,
// no FMA instructions in the host build), so host, device and the CPU oracle agree bit for bit.
// Model::computeFusionWeight (Model.cpp:391-406) takes the rotation angle through acos.
#ifdef __CUDA_ARCH__
#define CFB_DMUL(a, b) __dmul_rn((a), (b))
#define CFB_DADD(a, b) __dadd_rn((a), (b))
#else
#define CFB_DMUL(a, b) ((a) * (b))
#define CFB_DADD(a, b) ((a) + (b))
#endif
__host__ __device__ inline double det_polevl(double x, const double* c, int n) {
  double a = c[0];
  for (int i = 1; i <= n; ++i) a = CFB_DADD(CFB_DMUL(a, x), c[i]);
  return a;
}
__host__ __device__ inline double det_p1evl(double x, const double* c, int n) {
  double a = CFB_DADD(x, c[0]);
  for (int i = 1; i < n; ++i) a = CFB_DADD(CFB_DMUL(a, x), c[i]);
  return a;
}
__host__ __device__ inline double det_asin(double x) {
  const double P[6] = {4.253011369004428248960E-3, -6.019598008014123785661E-1, 5.444622390564711410273E0,
                       -1.626247967210700244449E1, 1.956261983317594739197E1, -8.198089802484824371615E0};
  const double Q[5] = {-1.474091372988853791896E1, 7.049610280856842141659E1, -1.471791292232726029859E2,
                       1.395105614657485689735E2, -4.918853881490881290097E1};
  const double R[5] = {2.967721961301243206100E-3, -5.634242780008963776856E-1, 6.968710824104713396794E0,
                       -2.556901049652824852289E1, 2.853665548261061424989E1};
  const double S[4] = {-2.194779531642920639778E1, 1.470656354026814941758E2, -3.838770957603691357202E2,
                       3.424398657913078477438E2};
  const double PIO4 = 7.85398163397448309616E-1, MOREBITS = 6.123233995736765886130E-17;
  const double a = x < 0 ? -x : x;
  double z;
  if (a > 0.625) {
    double zz = CFB_DADD(1.0, -a);
    const double p = CFB_DMUL(zz, det_polevl(zz, R, 4)) / det_p1evl(zz, S, 4);
    zz = sqrt(CFB_DADD(zz, zz));
    z = CFB_DADD(PIO4, -zz);
    zz = CFB_DADD(CFB_DMUL(zz, p), -MOREBITS);
    z = CFB_DADD(z, -zz);
    z = CFB_DADD(z, PIO4);
  } else {
    if (a < 1.0e-8) return x;
    const double zz = CFB_DMUL(a, a);
    z = CFB_DMUL(zz, det_polevl(zz, P, 5)) / det_p1evl(zz, Q, 5);
    z = CFB_DADD(CFB_DMUL(a, z), a);
  }
  return x < 0 ? -z : z;
}
__host__ __device__ inline double det_acos(double x) {  // x in [-1, 1]
  const double PIO4 = 7.85398163397448309616E-1, MOREBITS = 6.123233995736765886130E-17;
  if (x > 0.5) return CFB_DMUL(2.0, det_asin(sqrt(CFB_DADD(0.5, -CFB_DMUL(0.5, x)))));
  double z = CFB_DADD(PIO4, -det_asin(x));
  z = CFB_DADD(z, MOREBITS);
  z = CFB_DADD(z, PIO4);
  return z;
}

}  // namespace cfb
