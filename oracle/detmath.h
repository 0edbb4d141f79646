/* detmath.h -- oracle-side copy of the deterministic f32 exp (TEST INFRASTRUCTURE, see cf_oracle.h).
 * GLSL exp() has implementation-defined precision (depth_bilateral_metric.frag:64, surfels.glsl:45),
 * so the oracle freezes one polynomial realisation; only +,-,*,rint are used so that a GPU kernel
 * built without FMA contraction reproduces it bit for bit. */
#ifndef ORC_DETMATH_H_
#define ORC_DETMATH_H_
#include <math.h>
#include <stdint.h>
static inline float orc_expf(float x) {
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
  int e = (int)n;
  union { uint32_t u; float f; } s;
  s.u = (uint32_t)(e + 127) << 23;
  return y * s.f;
}
#endif
