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
  const int e = (int)n;  // in [-126, 127]
  union {
    unsigned u;
    float f;
  } s;
  s.u = (unsigned)(e + 127) << 23;
  const float v = y * s.f;
  return (x_in >= -87.0f) ? v : ((x_in != x_in) ? x_in : 0.0f);
}

}  // namespace cfb
