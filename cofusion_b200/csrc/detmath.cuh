// detmath.cuh -- deterministic f32 math used where results must be BIT-EXACT between the CUDA
// kernels and the CPU oracle (bilateral weights, surfel confidence): only +,-,*,/ and rint, each
// correctly rounded on both sides (this translation unit is compiled with -fmad=false and the
// default -prec-div=true -prec-sqrt=true -ftz=false; the oracle with -ffp-contract=off).
// GLSL leaves exp() precision implementation-defined (the reference runs it in a shader,
// depth_bilateral_metric.frag:64, surfels.glsl:45), so a fixed polynomial is a valid realisation.
#pragma once

namespace cfb {

// Cephes-style expf: |rel err| ~ 1 ulp on [-87, 88].
__host__ __device__ __forceinline__ float det_expf(float x) {
  if (!(x >= -87.0f)) return (x != x) ? x : 0.0f;
  if (x > 88.0f) x = 88.0f;
  float t = x * 1.44269504088896341f;
  float n = rintf(t);
  float r = x - n * 0.693359375f;
  r = r - n * -2.12194440e-4f;
  float p = 1.9875691500E-4f;
  p = p * r + 1.3981999507E-3f;
  p = p * r + 8.3334519073E-3f;
  p = p * r + 4.1665795894E-2f;
  p = p * r + 1.6666665459E-1f;
  p = p * r + 5.0000001201E-1f;
  float y = (p * (r * r) + r) + 1.0f;
  int e = (int)n;  // in [-126, 127]
  union {
    unsigned u;
    float f;
  } s;
  s.u = (unsigned)(e + 127) << 23;
  return y * s.f;
}

}  // namespace cfb
