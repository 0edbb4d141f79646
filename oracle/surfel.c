/*
 * surfel.c -- CPU oracle (TEST INFRASTRUCTURE, see cf_oracle.h) for the surfel-map stage that the
 * reference runs as OpenGL 3.3 shaders: first-frame initialisation, index-map prediction, data
 * association + fusion, cleaning/compaction, splat prediction and fill-in.
 *
 * PARITY UNPINNED: there is no GL context in this image and the reference has no golden vectors, so
 * this line-by-line restatement of the GLSL *defines* the expected result.  GL semantics that the
 * shader text does not fix are frozen here (and documented in DESIGN.md):
 *   F1  textures are sampled NEAREST, clamp-to-edge; texel = floor(coord * size) in f32 with the
 *       shader's own float arithmetic for the coordinate (incl. the half-texel window steps of
 *       data.vert:138-139 / copy_unstable.vert:86-87);
 *   F2  a GL point lands on pixel floor(window xy); it is clipped when its centre is outside
 *       [0,W)x[0,H);
 *   F3  depth test is GL_LESS (GUI/Tools/GUI.h:94-96) on a 24-bit fixed-point buffer:
 *       key = round(depth * (2^24-1)); equal keys -> the earlier primitive (lower surfel id /
 *       earlier pixel in column-major draw order) wins;
 *   F4  a point sprite of size s covers the pixels whose centres (px+.5, py+.5) satisfy
 *       xw - s/2 <= px+.5 < xw + s/2 (same in y); sizes below 1 are raised to 1;
 *   F5  exp() = orc_expf (detmath.h), normalize(v) = v / sqrt(dot(v,v)), round() = half away from
 *       zero, and `acos(c) < 0.5` is evaluated as `c <= 1 && c > cos(0.5)` (acos of c > 1 is NaN in
 *       GLSL, so the comparison is false -- kept);
 *   F6  the per-pixel draw order of data.vert / vertex_feedback.vert is column-major
 *       (Model.cpp:164-170, FeedbackBuffer.cpp:44-50).
 * Compile with -ffp-contract=off.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "cf_oracle.h"
#include "detmath.h"

#define COS_HALF 0.87758256189037276f

struct OrcSurfelMap {
  int W, H;
  float fx, fy, cx, cy;
  unsigned capacity;
  OrcSurfel* buf[2];
  int target, renderSource; /* Model.h:251: swapped after fuse and clean */
  unsigned count;
  OrcSurfel* unstable; /* newUnstableBuffer, W*H */
  unsigned unstableCount;
  /* ModelProjection sparse index targets (ModelProjection.cpp:72-76) */
  uint32_t* index;
  float *vertConf, *colorTime, *normRad; /* float4 each */
  /* combinedPredict targets (:90-94) */
  uint8_t* image;   /* RGBA8 */
  float *splatVertexConf, *splatNormalRad; /* float4 */
  uint16_t* splatTime;
  /* FillIn targets (FillIn.cpp:21-23) */
  uint8_t* fillImage;
  float *fillVertex, *fillNormal;
  /* scratch */
  uint64_t* keys;
  uint32_t* winner;
};

OrcSurfelMap* orc_map_create(int W, int H, float fx, float fy, float cx, float cy, unsigned max_surfels) {
  OrcSurfelMap* m = (OrcSurfelMap*)calloc(1, sizeof(OrcSurfelMap));
  size_t n = (size_t)W * H;
  m->W = W;
  m->H = H;
  m->fx = fx;
  m->fy = fy;
  m->cx = cx;
  m->cy = cy;
  m->capacity = max_surfels;
  m->buf[0] = (OrcSurfel*)calloc(max_surfels, sizeof(OrcSurfel));
  m->buf[1] = (OrcSurfel*)calloc(max_surfels, sizeof(OrcSurfel));
  m->target = 0;
  m->renderSource = 1;
  m->unstable = (OrcSurfel*)calloc(n, sizeof(OrcSurfel));
  m->index = (uint32_t*)calloc(n, 4);
  m->vertConf = (float*)calloc(n * 4, 4);
  m->colorTime = (float*)calloc(n * 4, 4);
  m->normRad = (float*)calloc(n * 4, 4);
  m->image = (uint8_t*)calloc(n * 4, 1);
  m->splatVertexConf = (float*)calloc(n * 4, 4);
  m->splatNormalRad = (float*)calloc(n * 4, 4);
  m->splatTime = (uint16_t*)calloc(n, 2);
  m->fillImage = (uint8_t*)calloc(n * 4, 1);
  m->fillVertex = (float*)calloc(n * 4, 4);
  m->fillNormal = (float*)calloc(n * 4, 4);
  m->keys = (uint64_t*)calloc(n, 8);
  m->winner = (uint32_t*)calloc(max_surfels, 4);
  return m;
}

void orc_map_destroy(OrcSurfelMap* m) {
  if (!m) return;
  free(m->buf[0]);
  free(m->buf[1]);
  free(m->unstable);
  free(m->index);
  free(m->vertConf);
  free(m->colorTime);
  free(m->normRad);
  free(m->image);
  free(m->splatVertexConf);
  free(m->splatNormalRad);
  free(m->splatTime);
  free(m->fillImage);
  free(m->fillVertex);
  free(m->fillNormal);
  free(m->keys);
  free(m->winner);
  free(m);
}

unsigned orc_map_count(const OrcSurfelMap* m) { return m->count; }
const OrcSurfel* orc_map_surfels(const OrcSurfelMap* m) { return m->buf[m->target]; }
unsigned orc_map_unstable_count(const OrcSurfelMap* m) { return m->unstableCount; }
const OrcSurfel* orc_map_unstable(const OrcSurfelMap* m) { return m->unstable; }
void orc_map_set_surfels(OrcSurfelMap* m, const OrcSurfel* s, unsigned n) {
  memcpy(m->buf[m->target], s, (size_t)n * sizeof(OrcSurfel));
  m->count = n;
}
const void* orc_map_view(const OrcSurfelMap* m, int which) {
  switch (which) {
    case 0: return m->index;
    case 1: return m->vertConf;
    case 2: return m->colorTime;
    case 3: return m->normRad;
    case 4: return m->image;
    case 5: return m->splatVertexConf;
    case 6: return m->splatNormalRad;
    case 7: return m->splatTime;
    case 8: return m->fillImage;
    case 9: return m->fillVertex;
    case 10: return m->fillNormal;
  }
  return 0;
}

/* ---------------------------------------------------------------- shared GLSL helpers */
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
/* F1: nearest texel of a normalised coordinate */
static inline int texel(float coord, int size) { return clampi((int)floorf(coord * (float)size), 0, size - 1); }

/* the uv attribute built on the host (Model.cpp:164-170, FeedbackBuffer.cpp:44-50) */
static inline float uv_coord(int i, int n) { return (float)((double)((float)i / (float)n) + 1.0 / (double)(2 * (float)n)); }

/* color_encoding.glsl:19-34 */
static inline float encode_color(float r, float g, float b) {
  int rgb = (int)roundf(r * 255.0f);
  rgb = (rgb << 8) + (int)roundf(g * 255.0f);
  rgb = (rgb << 8) + (int)roundf(b * 255.0f);
  return (float)rgb;
}
static inline void decode_color(float c, float out[3]) {
  out[0] = (float)(((int)c >> 16) & 0xFF) / 255.0f;
  out[1] = (float)(((int)c >> 8) & 0xFF) / 255.0f;
  out[2] = (float)((int)c & 0xFF) / 255.0f;
}
/* surfels.glsl:19-34; cam.zw are INVERSE focal lengths (f32 of 1.0/fx) */
static inline float get_radius(float depth, float norm_z, float inv_fx, float inv_fy) {
  float meanFocal = ((1.0f / fabsf(inv_fx)) + (1.0f / fabsf(inv_fy))) / 2.0f;
  const float sqrt2 = 1.41421356237f;
  float radius = (depth / meanFocal) * sqrt2;
  float radius_n = radius / fabsf(norm_z);
  radius_n = fminf(2.0f * radius, radius_n);
  return radius_n;
}
/* surfels.glsl:36-46 */
ORC_FMA_CLONES static float confidence(float x, float y, float cx, float cy, float weighting) {
  const float maxRadDist = 400, twoSigmaSquared = 0.72f;
  float px = x - cx, py = y - cy;
  float radialDist = sqrtf(px * px + py * py) / maxRadDist;
  return orc_expf((-(radialDist * radialDist) / twoSigmaSquared)) * weighting;
}
static inline void normalize3(float v[3]) {
  float l = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  v[0] /= l;
  v[1] /= l;
  v[2] /= l;
}
static inline void cross3f(const float a[3], const float b[3], float o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
/* geometry.glsl:21-25: vertex of texel (tx,ty) at float pixel coordinate (x,y) */
static inline void get_vertex(const float* depth, int W, int tx, int ty, float x, float y, float cx, float cy,
                              float inv_fx, float inv_fy, float v[3]) {
  float z = depth[ty * W + tx];
  v[0] = (x - cx) * z * inv_fx;
  v[1] = (y - cy) * z * inv_fy;
  v[2] = z;
}
/* geometry.glsl:28-40: central differences (float overload) */
static void get_normal_central(const float* depth, int W, int H, float tcx, float tcy, float x, float y, float cx,
                               float cy, float inv_fx, float inv_fy, const float vpos[3], float n[3]) {
  float cols = (float)W, rows = (float)H;
  float xf[3], xb[3], yf[3], yb[3];
  get_vertex(depth, W, texel(tcx + (1.0f / cols), W), texel(tcy, H), x + 1, y, cx, cy, inv_fx, inv_fy, xf);
  get_vertex(depth, W, texel(tcx - (1.0f / cols), W), texel(tcy, H), x - 1, y, cx, cy, inv_fx, inv_fy, xb);
  get_vertex(depth, W, texel(tcx, W), texel(tcy + (1.0f / rows), H), x, y + 1, cx, cy, inv_fx, inv_fy, yf);
  get_vertex(depth, W, texel(tcx, W), texel(tcy - (1.0f / rows), H), x, y - 1, cx, cy, inv_fx, inv_fy, yb);
  float dx[3], dy[3];
  for (int k = 0; k < 3; ++k) {
    dx[k] = ((xb[k] + vpos[k]) / 2) - ((xf[k] + vpos[k]) / 2);
    dy[k] = ((yb[k] + vpos[k]) / 2) - ((yf[k] + vpos[k]) / 2);
  }
  cross3f(dx, dy, n);
  normalize3(n);
}
static inline void mat4_mul_point(const float* T, const float p[3], float o[3]) {
  for (int r = 0; r < 3; ++r) o[r] = T[r * 4 + 0] * p[0] + T[r * 4 + 1] * p[1] + T[r * 4 + 2] * p[2] + T[r * 4 + 3];
}
static inline void mat3_mul_vec(const float* T, const float p[3], float o[3]) {
  for (int r = 0; r < 3; ++r) o[r] = T[r * 4 + 0] * p[0] + T[r * 4 + 1] * p[1] + T[r * 4 + 2] * p[2];
}
/* Eigen pose.inverse() of a rigid transform, f32 (ModelProjection.cpp:119, Model.cpp:598) */
void orc_pose_inverse(const float T[16], float Ti[16]) {
  memset(Ti, 0, 64);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) Ti[r * 4 + c] = T[c * 4 + r];
  for (int r = 0; r < 3; ++r) Ti[r * 4 + 3] = -(Ti[r * 4 + 0] * T[3] + Ti[r * 4 + 1] * T[7] + Ti[r * 4 + 2] * T[11]);
  Ti[15] = 1;
}
/* F3 */
static inline uint32_t depth_key24(float depth01) {
  double d = (double)depth01;
  if (!(d > 0)) d = 0;
  if (d > 1) d = 1;
  return (uint32_t)floor(d * 16777215.0 + 0.5);
}

/* ---------------------------------------------------------------- a18: first-frame initialisation
 * vertex_feedback.vert/.geom run twice (raw depth, filtered depth; CoFusion.cpp:161-169), then
 * init_unstable.vert pairs the i-th emitted RAW vertex (position, colour) with the i-th emitted
 * FILTERED vertex (normal, radius) (Model.cpp:227-272).  The pairing is by compacted index -- kept. */
static unsigned vertex_feedback(const OrcSurfelMap* m, const uint8_t* rgb, const float* depth, int time,
                                float maxDepth, OrcSurfel* out) {
  const int W = m->W, H = m->H;
  const float inv_fx = 1.0f / m->fx, inv_fy = 1.0f / m->fy; /* FeedbackBuffer.cpp:81-82 (f32 division) */
  unsigned n = 0;
  for (int i = 0; i < W; ++i)
    for (int j = 0; j < H; ++j) { /* F6 column-major */
      float tcx = uv_coord(i, W), tcy = uv_coord(j, H);
      float x = tcx * (float)W, y = tcy * (float)H;
      int tx = texel(tcx, W), ty = texel(tcy, H);
      float v[3], nrm[3];
      get_vertex(depth, W, tx, ty, x, y, m->cx, m->cy, inv_fx, inv_fy, v);
      get_normal_central(depth, W, H, tcx, tcy, x, y, m->cx, m->cy, inv_fx, inv_fy, v, nrm);
      float zVal = (v[2] <= 0 || v[2] > maxDepth) ? 0 : v[2];
      if (!(zVal > 0)) continue;
      OrcSurfel s;
      s.pos[0] = v[0];
      s.pos[1] = v[1];
      s.pos[2] = v[2];
      s.pos[3] = confidence(x, y, m->cx, m->cy, 1.0f);
      const uint8_t* c = &rgb[(ty * W + tx) * 3];
      s.col[0] = encode_color(c[0] / 255.0f, c[1] / 255.0f, c[2] / 255.0f);
      s.col[1] = 0;
      s.col[2] = c[2] / 255.0f; /* vColor.z is left as the blue channel by vertex_feedback.vert */
      s.col[3] = (float)time;
      s.nrm[0] = nrm[0];
      s.nrm[1] = nrm[1];
      s.nrm[2] = nrm[2];
      s.nrm[3] = get_radius(v[2], nrm[2], inv_fx, inv_fy);
      out[n++] = s;
    }
  return n;
}

void orc_map_initialise(OrcSurfelMap* m, const uint8_t* rgb, const float* depthRaw, const float* depthFiltered,
                        int time, float maxDepth) {
  size_t n = (size_t)m->W * m->H;
  OrcSurfel* raw = (OrcSurfel*)calloc(n, sizeof(OrcSurfel));
  OrcSurfel* fil = (OrcSurfel*)calloc(n, sizeof(OrcSurfel));
  unsigned nr = vertex_feedback(m, rgb, depthRaw, time, maxDepth, raw);
  vertex_feedback(m, rgb, depthFiltered, time, maxDepth, fil);
  OrcSurfel* dst = m->buf[m->target];
  for (unsigned i = 0; i < nr && i < m->capacity; ++i) {
    OrcSurfel s = raw[i];
    s.col[1] = 0; /* init_unstable.vert:31-35 */
    s.col[2] = 1;
    memcpy(s.nrm, fil[i].nrm, sizeof(s.nrm));
    dst[i] = s;
  }
  m->count = nr < m->capacity ? nr : m->capacity;
  free(raw);
  free(fil);
}

/* ---------------------------------------------------------------- a13: index map
 * index_map.vert/.frag, ModelProjection::predictIndices (ModelProjection.cpp:105-157). */
void orc_map_predict_indices(OrcSurfelMap* m, const float pose[16], int time, float maxDepth, int timeDelta) {
  const int W = m->W, H = m->H;
  const size_t n = (size_t)W * H;
  float t_inv[16];
  orc_pose_inverse(pose, t_inv);
  for (size_t i = 0; i < n; ++i) m->keys[i] = ~(uint64_t)0;
  const OrcSurfel* S = m->buf[m->target];
  const float cols = (float)W, rows = (float)H;
  for (unsigned id = 0; id < m->count; ++id) {
    const OrcSurfel* s = &S[id];
    float ph[3];
    mat4_mul_point(t_inv, s->pos, ph);
    if (ph[2] > maxDepth || ph[2] < 0 || (float)time - s->col[3] > (float)timeDelta) continue;
    /* NDC -> window exactly as the fixed-function viewport transform of (x_ndc + 1) * W/2 */
    float xn = ((((m->fx * ph[0]) / ph[2]) + m->cx) - (cols * 0.5f)) / (cols * 0.5f);
    float yn = ((((m->fy * ph[1]) / ph[2]) + m->cy) - (rows * 0.5f)) / (rows * 0.5f);
    float zn = ph[2] / maxDepth;
    if (!(xn >= -1.0f && xn <= 1.0f && yn >= -1.0f && yn <= 1.0f)) continue; /* clipped (also NaN) */
    float xw = (xn + 1.0f) * (cols * 0.5f), yw = (yn + 1.0f) * (rows * 0.5f);
    int px = (int)floorf(xw), py = (int)floorf(yw);
    if (px < 0 || py < 0 || px >= W || py >= H) continue;
    uint64_t key = ((uint64_t)depth_key24(zn * 0.5f + 0.5f) << 32) | id;
    if (key < m->keys[py * W + px]) m->keys[py * W + px] = key;
  }
  for (size_t i = 0; i < n; ++i) {
    float* vc = &m->vertConf[i * 4];
    float* ct = &m->colorTime[i * 4];
    float* nr = &m->normRad[i * 4];
    if (m->keys[i] == ~(uint64_t)0) {
      m->index[i] = 0;
      memset(vc, 0, 16);
      memset(ct, 0, 16);
      memset(nr, 0, 16);
      continue;
    }
    unsigned id = (unsigned)(m->keys[i] & 0xffffffffu);
    const OrcSurfel* s = &S[id];
    float ph[3], nl[3];
    mat4_mul_point(t_inv, s->pos, ph);
    mat3_mul_vec(t_inv, s->nrm, nl);
    normalize3(nl);
    m->index[i] = id;
    vc[0] = ph[0];
    vc[1] = ph[1];
    vc[2] = ph[2];
    vc[3] = s->pos[3];
    memcpy(ct, s->col, 16);
    nr[0] = nl[0];
    nr[1] = nl[1];
    nr[2] = nl[2];
    nr[3] = s->nrm[3];
  }
}

/* ---------------------------------------------------------------- a16: fuse
 * pass 1 = data.vert/.geom/.frag (Model.cpp:410-497), pass 2 = update.vert (:499-562). */
float orc_fusion_weight(const float pose[16], const float lastPose[16], float weightMultiplier) {
  /* Model::computeFusionWeight (Model.cpp:391-406): diff = pose^-1 * lastPose */
  float pinv[16], d[16];
  orc_pose_inverse(pose, pinv);
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      float s = 0;
      for (int k = 0; k < 4; ++k) s += pinv[r * 4 + k] * lastPose[k * 4 + c];
      d[r * 4 + c] = s;
    }
  float tn = sqrtf(d[3] * d[3] + d[7] * d[7] + d[11] * d[11]);
  /* rodrigues2 (Model.cpp:816-857) without the SVD re-orthogonalisation */
  double rx = d[9] - d[6], ry = d[2] - d[8], rz = d[4] - d[1];
  double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
  double c = ((double)d[0] + d[5] + d[10] - 1) * 0.5;
  c = c > 1. ? 1. : c < -1. ? -1. : c;
  double theta = orc_acos(c); /* deterministic realisation shared with the GPU side (detmath.h) */
  if (s < 1e-5) {
    if (c > 0)
      rx = ry = rz = 0;
    else {
      double t = (d[0] + 1) * 0.5;
      rx = sqrt(t > 0 ? t : 0);
      t = (d[5] + 1) * 0.5;
      ry = sqrt(t > 0 ? t : 0) * (d[1] < 0 ? -1.0 : 1.0);
      t = (d[10] + 1) * 0.5;
      rz = sqrt(t > 0 ? t : 0) * (d[2] < 0 ? -1.0 : 1.0);
      if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && (d[6] > 0) != (ry * rz > 0)) rz = -rz;
      theta /= sqrt(rx * rx + ry * ry + rz * rz);
      rx *= theta;
      ry *= theta;
      rz *= theta;
    }
  } else {
    double vth = 1 / (2 * s) * theta;
    rx *= vth;
    ry *= vth;
    rz *= vth;
  }
  float rn = sqrtf((float)rx * (float)rx + (float)ry * (float)ry + (float)rz * (float)rz);
  float weighting = tn > rn ? tn : rn;
  const float largest = 0.01f, minWeight = 0.5f;
  if (weighting > largest) weighting = largest;
  float w = 1.0f - (weighting / largest);
  weighting = (w > minWeight ? w : minWeight) * weightMultiplier;
  return weighting;
}

void orc_map_fuse(OrcSurfelMap* m, const float pose[16], int time, const uint8_t* rgb, const uint8_t* mask,
                  const float* depthRaw, const float* depthFiltered, float maxDepth, float weighting,
                  unsigned maskID) {
  const int W = m->W, H = m->H;
  const float cols = (float)W, rows = (float)H, scale = 1.0f;
  const float inv_fx = (float)(1.0 / m->fx), inv_fy = (float)(1.0 / m->fy);
  OrcSurfel* S = m->buf[m->target];
  for (unsigned i = 0; i < m->count; ++i) m->winner[i] = 0xffffffffu;
  /* per-pixel update records, indexed by the column-major ordinal */
  size_t n = (size_t)W * H;
  OrcSurfel* upd = (OrcSurfel*)malloc(n * sizeof(OrcSurfel));
  uint32_t* updBest = (uint32_t*)calloc(n, 4);
  m->unstableCount = 0;
  const float ftime = (float)time;
  for (int i = 0; i < W; ++i)
    for (int j = 0; j < H; ++j) {
      const unsigned ord = (unsigned)i * H + j;
      float tcx = uv_coord(i, W), tcy = uv_coord(j, H);
      float x = tcx * cols, y = tcy * rows;
      int tx = texel(tcx, W), ty = texel(tcy, H);
      float vl[3], vf[3], vg[3];
      get_vertex(depthRaw, W, tx, ty, x, y, m->cx, m->cy, inv_fx, inv_fy, vl);
      mat4_mul_point(pose, vl, vg);
      get_vertex(depthFiltered, W, tx, ty, x, y, m->cx, m->cy, inv_fx, inv_fy, vf);
      /* eligibility (data.vert:116-119) */
      if (!((int)x % 2 == (int)ftime % 2 && (int)y % 2 == (int)ftime % 2)) continue;
      if ((unsigned)mask[ty * W + tx] != maskID) continue;
      /* checkNeighbours on the RAW depth (data.vert:52-71) */
      if (depthRaw[ty * W + texel(tcx - (1.0f / cols), W)] == 0) continue;
      if (depthRaw[texel(tcy - (1.0f / rows), H) * W + tx] == 0) continue;
      if (depthRaw[ty * W + texel(tcx + (1.0f / cols), W)] == 0) continue;
      if (depthRaw[texel(tcy + (1.0f / rows), H) * W + tx] == 0) continue;
      if (!(vl[2] > 0 && vl[2] <= maxDepth)) continue;

      OrcSurfel c;
      float nl[3], ng[3];
      get_normal_central(depthFiltered, W, H, tcx, tcy, x, y, m->cx, m->cy, inv_fx, inv_fy, vf, nl);
      mat3_mul_vec(pose, nl, ng);
      c.pos[0] = vg[0];
      c.pos[1] = vg[1];
      c.pos[2] = vg[2];
      c.pos[3] = confidence(x, y, m->cx, m->cy, weighting);
      const uint8_t* col = &rgb[(ty * W + tx) * 3];
      c.col[0] = encode_color(col[0] / 255.0f, col[1] / 255.0f, col[2] / 255.0f);
      c.col[1] = 0;
      c.col[2] = ftime;
      c.col[3] = 0;
      c.nrm[0] = ng[0];
      c.nrm[1] = ng[1];
      c.nrm[2] = ng[2];
      c.nrm[3] = get_radius(vf[2], nl[2], inv_fx, inv_fy);

      /* association window (data.vert:127-163) */
      int operation = 0;
      uint32_t best = 0;
      float indexXStep = (1.0f / (cols * scale)) * 0.5f, indexYStep = (1.0f / (rows * scale)) * 0.5f;
      float bestDist = 1000;
      const float windowMultiplier = 2;
      float xl = (x - m->cx) * inv_fx, yl = (y - m->cy) * inv_fy;
      float lambda = sqrtf(xl * xl + yl * yl + 1);
      float ray[3] = {xl, yl, 1};
      for (float si = tcx - (scale * indexXStep * windowMultiplier); si < tcx + (scale * indexXStep * windowMultiplier);
           si += indexXStep)
        for (float sj = tcy - (scale * indexYStep * windowMultiplier);
             sj < tcy + (scale * indexYStep * windowMultiplier); sj += indexYStep) {
          int sx = texel(si, W), sy = texel(sj, H);
          uint32_t current = m->index[sy * W + sx];
          if (current > 0U) {
            const float* vc = &m->vertConf[(sy * W + sx) * 4];
            float zdiff = vc[2] - vl[2];
            if (fabsf(zdiff * lambda) < 0.05f) {
              float cr[3];
              cross3f(ray, vc, cr);
              float dist = sqrtf(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
              const float* nr = &m->normRad[(sy * W + sx) * 4];
              /* angleBetween (data.vert:73-76), F5 */
              float cosang = (nr[0] * nl[0] + nr[1] * nl[1] + nr[2] * nl[2]) /
                             (sqrtf(nr[0] * nr[0] + nr[1] * nr[1] + nr[2] * nr[2]) *
                              sqrtf(nl[0] * nl[0] + nl[1] * nl[1] + nl[2] * nl[2]));
              int angle_ok = (cosang <= 1.0f && cosang > COS_HALF);
              if (dist < bestDist && (fabsf(nr[2]) < 0.75f || angle_ok)) {
                operation = 1;
                bestDist = dist;
                best = current;
              }
            }
          }
        }
      if (operation == 1) {
        c.col[3] = -1;
        upd[ord] = c;
        updBest[ord] = best;
        if (ord < m->winner[best]) m->winner[best] = ord; /* first in draw order wins (F3) */
      } else {
        c.col[3] = -2;
      }
      /* data.geom emits every vertex with updateId > 0 into newUnstableBuffer, in draw order */
      m->unstable[m->unstableCount++] = c;
    }

  /* pass 2: update.vert, in place (each surfel is touched by at most one winning pixel) */
  for (unsigned id = 0; id < m->count; ++id) {
    uint32_t ord = m->winner[id];
    if (ord == 0xffffffffu) continue;
    const OrcSurfel* nw = &upd[ord];
    OrcSurfel* o = &S[id];
    float c_k = o->pos[3], a = nw->pos[3];
    if (nw->nrm[3] < (1.0f + 0.5f) * o->nrm[3]) {
      for (int k = 0; k < 3; ++k) o->pos[k] = ((c_k * o->pos[k]) + (a * nw->pos[k])) / (c_k + a);
      o->pos[3] = c_k + a;
      float oc[3], nc[3];
      decode_color(o->col[0], oc);
      decode_color(nw->col[0], nc);
      float avg[3];
      for (int k = 0; k < 3; ++k) avg[k] = ((c_k * oc[k]) + (a * nc[k])) / (c_k + a);
      o->col[0] = encode_color(avg[0], avg[1], avg[2]);
      o->col[3] = ftime;
      float nr[4];
      for (int k = 0; k < 4; ++k) nr[k] = ((c_k * o->nrm[k]) + (a * nw->nrm[k])) / (c_k + a);
      normalize3(nr);
      memcpy(o->nrm, nr, 16);
    } else {
      o->pos[3] = c_k + a;
      o->col[3] = ftime;
    }
  }
  free(upd);
  free(updBest);
  /* the reference ping-pongs the VBOs here (Model.cpp:559); the oracle updates in place */
}

/* ---------------------------------------------------------------- a17: clean
 * copy_unstable.vert/.geom (Model.cpp:565-697); the deformation-graph block (:155-335) is dead
 * (nodes == 0 always). */
static int clean_one(const OrcSurfelMap* m, const float t_inv[16], OrcSurfel* s, int time, float confThreshold,
                     int timeDelta, const float* depthFiltered, const uint8_t* mask, unsigned maskID,
                     float outlierCoeff) {
  const int W = m->W, H = m->H;
  const float cols = (float)W, rows = (float)H, scale = 1.0f;
  int test = 1;
  float lp[3], ln[3];
  mat4_mul_point(t_inv, s->pos, lp);
  float x = ((m->fx * lp[0]) / lp[2]) + m->cx;
  float y = ((m->fy * lp[1]) / lp[2]) + m->cy;
  mat3_mul_vec(t_inv, s->nrm, ln);
  normalize3(ln);
  float x_n = x / cols, y_n = y / rows;
  float stepX = 1.0f / cols, stepY = 1.0f / rows;
  float indexXStep = stepX * 0.5f / scale, indexYStep = stepY * 0.5f / scale;
  const float windowMultiplier = 2;
  int count = 0, zCount = 0, violationCount = 0;
  float avgViolation = 0;
  const float ftime = (float)time;
  if (ftime - s->col[3] < (float)timeDelta && lp[2] > 0 && x > 0 && y > 0 && x < cols && y < rows) {
    for (float i = x_n - (scale * indexXStep * windowMultiplier); i < x_n + (scale * indexXStep * windowMultiplier);
         i += indexXStep)
      for (float j = y_n - (scale * indexYStep * windowMultiplier); j < y_n + (scale * indexYStep * windowMultiplier);
           j += indexYStep) {
        int sx = texel(i, W), sy = texel(j, H);
        uint32_t current = m->index[sy * W + sx];
        if (current > 0U) {
          const float* vc = &m->vertConf[(sy * W + sx) * 4];
          const float* ct = &m->colorTime[(sy * W + sx) * 4];
          float ddx = vc[0] - lp[0], ddy = vc[1] - lp[1];
          if (ct[2] < s->col[2] && vc[3] > confThreshold && vc[2] > lp[2] && vc[2] - lp[2] < 0.01f &&
              sqrtf(ddx * ddx + ddy * ddy) < s->nrm[3] * 1.4f)
            count++;
          if (ct[3] == ftime && vc[3] > confThreshold && vc[2] > lp[2] && vc[2] - lp[2] > 0.01f &&
              fabsf(ln[2]) > 0.85f)
            zCount++;
        }
      }
    for (float i = x_n - stepX; i <= x_n + stepX; i += stepX)
      for (float j = y_n - stepY; j <= y_n + stepY; j += stepY) {
        float d = depthFiltered[texel(j, H) * W + texel(i, W)] - lp[2];
        if (d > 0.03f) {
          violationCount++;
          avgViolation += d;
        }
      }
  }
  if (count > 8 || zCount > 4) test = 0;
  if (s->col[3] == -2) s->col[3] = ftime;
  if ((s->col[3] == -1 || ((ftime - s->col[3]) > 20 && s->pos[3] < confThreshold))) test = 0;
  if (s->col[3] > 0 && ftime - s->col[3] > (float)timeDelta) test = 1;
  if (violationCount > 0) {
    avgViolation /= (float)violationCount;
    s->pos[3] *= 1.0f / (1 + outlierCoeff * avgViolation);
    int sx = texel(x_n, W), sy = texel(y_n, H);
    unsigned maskValue = mask[sy * W + sx];
    float wDepth = depthFiltered[sy * W + sx];
    if (maskValue != maskID && (wDepth > lp[2] - 0.05f && wDepth < lp[2] + 0.05f))
      s->pos[3] *= (0.5f + 0.5f * (1 - outlierCoeff / 10.0f));
  }
  return test;
}

void orc_map_clean(OrcSurfelMap* m, const float pose[16], int time, float confThreshold, int timeDelta,
                   const float* depthFiltered, const uint8_t* mask, unsigned maskID, float outlierCoeff) {
  float t_inv[16];
  orc_pose_inverse(pose, t_inv);
  const OrcSurfel* src = m->buf[m->target];
  OrcSurfel* dst = m->buf[m->renderSource];
  unsigned n = 0;
  for (unsigned i = 0; i < m->count; ++i) {
    OrcSurfel s = src[i];
    if (clean_one(m, t_inv, &s, time, confThreshold, timeDelta, depthFiltered, mask, maskID, outlierCoeff) &&
        n < m->capacity)
      dst[n++] = s;
  }
  for (unsigned i = 0; i < m->unstableCount; ++i) {
    OrcSurfel s = m->unstable[i];
    if (clean_one(m, t_inv, &s, time, confThreshold, timeDelta, depthFiltered, mask, maskID, outlierCoeff) &&
        n < m->capacity)
      dst[n++] = s;
  }
  m->count = n;
  int t = m->target;
  m->target = m->renderSource;
  m->renderSource = t;
}

/* ---------------------------------------------------------------- a14: splat prediction
 * splat.vert + combo_splat.frag, ModelProjection::combinedPredict (ModelProjection.cpp:192-273). */
typedef struct {
  float ph[3], conf, nl[3], rad, colour, initTime;
  float xw, yw, size;
} SplatVtx;

static int splat_vertex(const OrcSurfelMap* m, const float t_inv[16], const OrcSurfel* s, float maxDepth,
                        float confThreshold, int time, int maxTime, int timeDelta, SplatVtx* o) {
  const float cols = (float)m->W, rows = (float)m->H;
  mat4_mul_point(t_inv, s->pos, o->ph);
  if (o->ph[2] > maxDepth || o->ph[2] < 0 || s->pos[3] < confThreshold || (float)time - s->col[3] > (float)timeDelta ||
      s->col[3] > (float)maxTime)
    return 0;
  float xn = ((((m->fx * o->ph[0]) / o->ph[2]) + m->cx) - (cols * 0.5f)) / (cols * 0.5f);
  float yn = ((((m->fy * o->ph[1]) / o->ph[2]) + m->cy) - (rows * 0.5f)) / (rows * 0.5f);
  if (!(xn >= -1.0f && xn <= 1.0f && yn >= -1.0f && yn <= 1.0f)) return 0; /* sprite culled by its centre */
  o->xw = (xn + 1.0f) * (cols * 0.5f);
  o->yw = (yn + 1.0f) * (rows * 0.5f);
  o->conf = s->pos[3];
  o->colour = s->col[0];
  o->initTime = s->col[2];
  mat3_mul_vec(t_inv, s->nrm, o->nl);
  normalize3(o->nl);
  o->rad = s->nrm[3];
  float x1[3] = {(o->nl[1] - o->nl[2]), -o->nl[0], o->nl[0]};
  normalize3(x1);
  for (int k = 0; k < 3; ++k) x1[k] = x1[k] * o->rad * 1.41421356f;
  float y1[3];
  cross3f(o->nl, x1, y1);
  float px[4], py[4];
  for (int q = 0; q < 4; ++q) {
    float p[3];
    for (int k = 0; k < 3; ++k) {
      float off = (q == 0) ? x1[k] : (q == 1) ? y1[k] : (q == 2) ? -y1[k] : -x1[k];
      p[k] = o->ph[k] + off;
    }
    px[q] = ((m->fx * p[0]) / p[2]) + m->cx;
    py[q] = ((m->fy * p[1]) / p[2]) + m->cy;
  }
  float xmin = fminf(px[0], fminf(px[1], fminf(px[2], px[3]))), xmax = fmaxf(px[0], fmaxf(px[1], fmaxf(px[2], px[3])));
  float ymin = fminf(py[0], fminf(py[1], fminf(py[2], py[3]))), ymax = fmaxf(py[0], fmaxf(py[1], fmaxf(py[2], py[3])));
  float xDiff = fabsf(xmax - xmin), yDiff = fabsf(ymax - ymin);
  o->size = fmaxf(0, fmaxf(xDiff, yDiff));
  if (!(o->size >= 1.0f)) o->size = 1.0f; /* F4 (also NaN) */
  if (o->size > 2047.0f) o->size = 2047.0f;
  return 1;
}

/* combo_splat.frag for surfel vertex v at pixel (px,py); returns 0 when discarded */
static int splat_fragment(const OrcSurfelMap* m, const SplatVtx* v, int px, int py, float maxDepth, float out_vc[4],
                          float* fragDepth) {
  float fcx = (float)px + 0.5f, fcy = (float)py + 0.5f;
  float l[3] = {(fcx - m->cx) / m->fx, (fcy - m->cy) / m->fy, 1.0f};
  normalize3(l);
  float k = (v->ph[0] * v->nl[0] + v->ph[1] * v->nl[1] + v->ph[2] * v->nl[2]) /
            (l[0] * v->nl[0] + l[1] * v->nl[1] + l[2] * v->nl[2]);
  float cp[3] = {k * l[0], k * l[1], k * l[2]};
  float sqrRad = v->rad * v->rad;
  float d[3] = {cp[0] - v->ph[0], cp[1] - v->ph[1], cp[2] - v->ph[2]};
  if (!(d[0] * d[0] + d[1] * d[1] + d[2] * d[2] <= sqrRad)) return 0; /* discard (NaN discards too) */
  float z = cp[2];
  out_vc[0] = (fcx - m->cx) * z * (1.f / m->fx);
  out_vc[1] = (fcy - m->cy) * z * (1.f / m->fy);
  out_vc[2] = z;
  out_vc[3] = v->conf;
  *fragDepth = (cp[2] / (2 * maxDepth)) + 0.5f;
  return 1;
}

void orc_map_combined_predict(OrcSurfelMap* m, const float pose[16], float maxDepth, float confThreshold, int time,
                              int maxTime, int timeDelta) {
  const int W = m->W, H = m->H;
  const size_t n = (size_t)W * H;
  float t_inv[16];
  orc_pose_inverse(pose, t_inv);
  for (size_t i = 0; i < n; ++i) m->keys[i] = ~(uint64_t)0;
  const OrcSurfel* S = m->buf[m->target];
  for (unsigned id = 0; id < m->count; ++id) {
    SplatVtx v;
    if (!splat_vertex(m, t_inv, &S[id], maxDepth, confThreshold, time, maxTime, timeDelta, &v)) continue;
    float h = v.size * 0.5f;
    int x0 = (int)ceilf(v.xw - h - 0.5f), x1 = (int)ceilf(v.xw + h - 0.5f) - 1; /* F4 */
    int y0 = (int)ceilf(v.yw - h - 0.5f), y1 = (int)ceilf(v.yw + h - 0.5f) - 1;
    x0 = x0 < 0 ? 0 : x0;
    y0 = y0 < 0 ? 0 : y0;
    x1 = x1 > W - 1 ? W - 1 : x1;
    y1 = y1 > H - 1 ? H - 1 : y1;
    for (int py = y0; py <= y1; ++py)
      for (int px = x0; px <= x1; ++px) {
        float vc[4], fd;
        if (!splat_fragment(m, &v, px, py, maxDepth, vc, &fd)) continue;
        if (!(fd >= 0.0f && fd <= 1.0f)) continue; /* depth clip */
        uint64_t key = ((uint64_t)depth_key24(fd) << 32) | id;
        if (key < m->keys[py * W + px]) m->keys[py * W + px] = key;
      }
  }
  for (int py = 0; py < H; ++py)
    for (int px = 0; px < W; ++px) {
      size_t i = (size_t)py * W + px;
      uint8_t* im = &m->image[i * 4];
      float* vc = &m->splatVertexConf[i * 4];
      float* nr = &m->splatNormalRad[i * 4];
      if (m->keys[i] == ~(uint64_t)0) {
        memset(im, 0, 4);
        memset(vc, 0, 16);
        memset(nr, 0, 16);
        m->splatTime[i] = 0;
        continue;
      }
      unsigned id = (unsigned)(m->keys[i] & 0xffffffffu);
      SplatVtx v;
      splat_vertex(m, t_inv, &S[id], maxDepth, confThreshold, time, maxTime, timeDelta, &v);
      float fd;
      splat_fragment(m, &v, px, py, maxDepth, vc, &fd);
      float col[3];
      decode_color(v.colour, col);
      for (int k = 0; k < 3; ++k) im[k] = (uint8_t)floorf(col[k] * 255.0f + 0.5f); /* RGBA8 UNORM store */
      im[3] = 255;
      nr[0] = v.nl[0];
      nr[1] = v.nl[1];
      nr[2] = v.nl[2];
      nr[3] = v.rad;
      m->splatTime[i] = (uint16_t)(unsigned)v.initTime;
    }
}

/* ---------------------------------------------------------------- a15: fill-in
 * fill_vertex/normal/rgb.frag via Model::performFillIn (Model.cpp:901-909); rawDepth is the
 * FILTERED metric depth (CoFusion.cpp:541). */
void orc_map_fill_in(OrcSurfelMap* m, const uint8_t* rgb, const float* depth, int passthrough_geom,
                     int passthrough_rgb) {
  const int W = m->W, H = m->H;
  const float inv_fx = 1.0f / m->fx, inv_fy = 1.0f / m->fy; /* FillIn.cpp: 1.0f / fx in f32 */
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      size_t i = (size_t)y * W + x;
      const float* sv = &m->splatVertexConf[i * 4];
      float* fv = &m->fillVertex[i * 4];
      if (sv[2] == 0 || passthrough_geom) {
        float z = depth[i];
        fv[0] = ((float)x - m->cx) * z * inv_fx;
        fv[1] = ((float)y - m->cy) * z * inv_fy;
        fv[2] = z;
        fv[3] = 1;
      } else
        memcpy(fv, sv, 16);
      const float* sn = &m->splatNormalRad[i * 4];
      float* fn = &m->fillNormal[i * 4];
      if (sn[2] == 0 || passthrough_geom) {
        /* forward differences, int overload (geometry.glsl:43-60), clamp-to-edge neighbours */
        int xp = x + 1 < W ? x + 1 : W - 1, yp = y + 1 < H ? y + 1 : H - 1;
        float z = depth[i], zx = depth[y * W + xp], zy = depth[yp * W + x];
        float v[3] = {((float)x - m->cx) * z * inv_fx, ((float)y - m->cy) * z * inv_fy, z};
        float vx[3] = {((float)(x + 1) - m->cx) * zx * inv_fx, ((float)y - m->cy) * zx * inv_fy, zx};
        float vy[3] = {((float)x - m->cx) * zy * inv_fx, ((float)(y + 1) - m->cy) * zy * inv_fy, zy};
        float dx[3] = {vx[0] - v[0], vx[1] - v[1], vx[2] - v[2]}, dy[3] = {vy[0] - v[0], vy[1] - v[1], vy[2] - v[2]};
        float nn[3];
        cross3f(dx, dy, nn);
        normalize3(nn);
        fn[0] = nn[0];
        fn[1] = nn[1];
        fn[2] = nn[2];
        fn[3] = 1;
      } else
        memcpy(fn, sn, 16);
      const uint8_t* si = &m->image[i * 4];
      uint8_t* fi = &m->fillImage[i * 4];
      if ((si[0] == 0 && si[1] == 0 && si[2] == 0) || passthrough_rgb) {
        fi[0] = rgb[i * 3];
        fi[1] = rgb[i * 3 + 1];
        fi[2] = rgb[i * 3 + 2];
        fi[3] = 255;
      } else
        memcpy(fi, si, 4);
    }
}

/* CoFusion::requiresFillIn (CoFusion.cpp:547-565): 20x subsampled RGB projection, < ratio filled */
int orc_map_requires_fill_in(const OrcSurfelMap* m, float ratio) {
  const int cons = 20, lw = m->W / cons, lh = m->H / cons;
  int sum = 0;
  for (int j = 0; j < lh; ++j)
    for (int i = 0; i < lw; ++i) {
      int sx = texel(((float)i + 0.5f) / (float)lw, m->W), sy = texel(((float)j + 0.5f) / (float)lh, m->H);
      const uint8_t* p = &m->image[((size_t)sy * m->W + sx) * 4];
      sum += p[0] > 0 && p[1] > 0 && p[2] > 0;
    }
  return (float)sum / (float)(lh * lw) < ratio;
}
