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
#endif
