// surfel_kernels.cu -- predict / fuse / clean / splat / fill-in of the surfel map, hand-written for
// sm_100a.  Compiled with -fmad=false (see build.py): every value that feeds an integer decision
// (pixel a surfel lands on, depth-test winner, association, survival) is computed with the same
// IEEE operation sequence as the CPU oracle, so index maps, ids and counts are bit-identical.
//
// How the GL pipeline of the reference is re-expressed:
//   * point rasterisation + GL_LESS depth test  -> one 64-bit atomicMin per surfel (or per sprite
//     fragment) of (24-bit depth key << 32 | surfel id) into a W*H key buffer, then a per-pixel
//     resolve pass that writes the attribute images of the winner only (the reference writes all
//     four render targets for every fragment that passes the test at the time it is drawn);
//   * the 3072^2 "update maps" that Model::fuse clears (453 MB per call, Model.cpp:413-420) and
//     scatters into  -> a per-surfel winner ordinal (atomicMin) + in-place update of touched surfels;
//   * transform-feedback append / geometry-shader compaction -> flag + exclusive scan + scatter,
//     order preserving (ids are observable, SURVEY.md section 7 hard part 2).
// Frozen GL semantics F1-F6 are the ones listed in the oracle (oracle/surfel.c header) / DESIGN.md.
#include "surfel_kernels.cuh"

#include <stdlib.h>

#include "detmath.cuh"

namespace cfb {
namespace {

#define COS_HALF 0.87758256189037276f

// ------------------------------------------------------------------------------- GLSL helpers
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ int texel(float coord, int size) {  // F1
  return clampi((int)floorf(coord * (float)size), 0, size - 1);
}
__device__ __forceinline__ float uv_coord(int i, int n) {  // Model.cpp:164-170
  return (float)((double)((float)i / (float)n) + 1.0 / (double)(2 * (float)n));
}
__device__ __forceinline__ float encode_color(float r, float g, float b) {  // color_encoding.glsl:19-25
  int rgb = (int)roundf(r * 255.0f);
  rgb = (rgb << 8) + (int)roundf(g * 255.0f);
  rgb = (rgb << 8) + (int)roundf(b * 255.0f);
  return (float)rgb;
}
__device__ __forceinline__ float3 decode_color(float c) {  // color_encoding.glsl:27-34
  return make_float3((float)(((int)c >> 16) & 0xFF) / 255.0f, (float)(((int)c >> 8) & 0xFF) / 255.0f,
                     (float)((int)c & 0xFF) / 255.0f);
}
__device__ __forceinline__ float get_radius(float depth, float norm_z, float inv_fx, float inv_fy) {  // surfels.glsl:19-34
  float meanFocal = ((1.0f / fabsf(inv_fx)) + (1.0f / fabsf(inv_fy))) / 2.0f;
  const float sqrt2 = 1.41421356237f;
  float radius = (depth / meanFocal) * sqrt2;
  float radius_n = radius / fabsf(norm_z);
  return fminf(2.0f * radius, radius_n);
}
__device__ __forceinline__ float confidence(float x, float y, float cx, float cy, float weighting) {  // surfels.glsl:36-46
  const float maxRadDist = 400, twoSigmaSquared = 0.72f;
  float px = x - cx, py = y - cy;
  float radialDist = sqrtf(px * px + py * py) / maxRadDist;
  return det_expf((-(radialDist * radialDist) / twoSigmaSquared)) * weighting;
}
__device__ __forceinline__ float3 normalize3(float3 v) {  // F5
  float l = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
  return make_float3(v.x / l, v.y / l, v.z / l);
}
__device__ __forceinline__ float3 cross3(float3 a, float3 b) {
  return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float3 get_vertex(const float* depth, int W, int tx, int ty, float x, float y,
                                             const SurfelGeom& g, float inv_fx, float inv_fy) {  // geometry.glsl:21-25
  float z = __ldg(depth + ty * W + tx);
  return make_float3((x - g.cx) * z * inv_fx, (y - g.cy) * z * inv_fy, z);
}
__device__ __forceinline__ float3 get_normal_central(const float* depth, const SurfelGeom& g, float tcx, float tcy,
                                                     float x, float y, float inv_fx, float inv_fy,
                                                     float3 vpos) {  // geometry.glsl:28-40
  const int W = g.W, H = g.H;
  const float cols = (float)W, rows = (float)H;
  float3 xf = get_vertex(depth, W, texel(tcx + (1.0f / cols), W), texel(tcy, H), x + 1, y, g, inv_fx, inv_fy);
  float3 xb = get_vertex(depth, W, texel(tcx - (1.0f / cols), W), texel(tcy, H), x - 1, y, g, inv_fx, inv_fy);
  float3 yf = get_vertex(depth, W, texel(tcx, W), texel(tcy + (1.0f / rows), H), x, y + 1, g, inv_fx, inv_fy);
  float3 yb = get_vertex(depth, W, texel(tcx, W), texel(tcy - (1.0f / rows), H), x, y - 1, g, inv_fx, inv_fy);
  float3 dx = make_float3(((xb.x + vpos.x) / 2) - ((xf.x + vpos.x) / 2), ((xb.y + vpos.y) / 2) - ((xf.y + vpos.y) / 2),
                          ((xb.z + vpos.z) / 2) - ((xf.z + vpos.z) / 2));
  float3 dy = make_float3(((yb.x + vpos.x) / 2) - ((yf.x + vpos.x) / 2), ((yb.y + vpos.y) / 2) - ((yf.y + vpos.y) / 2),
                          ((yb.z + vpos.z) / 2) - ((yf.z + vpos.z) / 2));
  return normalize3(cross3(dx, dy));
}
__device__ __forceinline__ Pose34 resolve_pose(const PoseRef& r) {
  if (!r.dev) return r.v;
  Pose34 p;
#pragma unroll
  for (int i = 0; i < 12; ++i) p.m[i] = __ldg(&r.dev->m[i]);
  return p;
}
__device__ __forceinline__ float3 xform_point(const Pose34& T, float3 p) {
  return make_float3(T.m[0] * p.x + T.m[1] * p.y + T.m[2] * p.z + T.m[3],
                     T.m[4] * p.x + T.m[5] * p.y + T.m[6] * p.z + T.m[7],
                     T.m[8] * p.x + T.m[9] * p.y + T.m[10] * p.z + T.m[11]);
}
__device__ __forceinline__ float3 xform_vec(const Pose34& T, float3 p) {
  return make_float3(T.m[0] * p.x + T.m[1] * p.y + T.m[2] * p.z, T.m[4] * p.x + T.m[5] * p.y + T.m[6] * p.z,
                     T.m[8] * p.x + T.m[9] * p.y + T.m[10] * p.z);
}
__device__ __forceinline__ unsigned depth_key24(float depth01) {  // F3
  double d = (double)depth01;
  if (!(d > 0)) d = 0;
  if (d > 1) d = 1;
  return (unsigned)floor(d * 16777215.0 + 0.5);
}
__device__ __forceinline__ Surfel load_surfel(const Surfel* s) {
  Surfel r;
  const float4* p = reinterpret_cast<const float4*>(s);
  r.pos = p[0];
  r.col = p[1];
  r.nrm = p[2];
  return r;
}
__device__ __forceinline__ void store_surfel(Surfel* d, const Surfel& s) {
  float4* p = reinterpret_cast<float4*>(d);
  p[0] = s.pos;
  p[1] = s.col;
  p[2] = s.nrm;
}

// ------------------------------------------------------------------------------- flag scan
// ranks[i] = number of set flags before i (tile-local prefix + scanned tile sums), total -> *total.
// Exclusive prefix sum of the survival flags in ONE launch (decoupled look-back): a CTA takes the next tile by a
// ticket, publishes its tile total, adds up the totals of the tiles before it (stopping at the first one that has
// already published an inclusive prefix) and writes the ranks.  Status word: epoch (30) | state (2) | value (32); the
// epoch and the ticket base come from the host, so nothing has to be cleared between scans.
constexpr unsigned kScanTile = 2048;  // 256 threads x 8 items
__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* q) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(q) : "memory");
  return v;
}
__device__ __forceinline__ void st_status(unsigned long long* q, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(q), "l"(v) : "memory");
}
__global__ void __launch_bounds__(256) scan_lookback_kernel(const uint8_t* __restrict__ flags, unsigned n_ub, const unsigned* n_dev,
                                                           unsigned n_extra, unsigned long long* status, unsigned* ticket,
                                                           unsigned ticket_base, unsigned epoch, uint32_t* __restrict__ ranks,
                                                           unsigned* total, unsigned ntiles) {
  pdl_prologue();
  __shared__ unsigned s_tile, s_prefix, ws[8];
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u) - ticket_base;
  __syncthreads();
  const unsigned tile = s_tile;
  const unsigned n = min(n_ub, (n_dev ? *n_dev : n_ub) + n_extra);
  const unsigned base = tile * kScanTile + threadIdx.x * 8;
  unsigned f[8], c = 0;
  if (base + 8 <= n) {
    const uint2 v = *reinterpret_cast<const uint2*>(flags + base);  // 8 flag bytes, 8-byte aligned
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      f[k] = (((k < 4 ? v.x : v.y) >> (8 * (k & 3))) & 0xffu) ? 1u : 0u;
      c += f[k];
    }
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      f[k] = (base + k < n && flags[base + k]) ? 1u : 0u;
      c += f[k];
    }
  }
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (unsigned)o) incl += t;
  }
  if (lane == 31) ws[warp] = incl;
  __syncthreads();
  unsigned woff = 0, agg = 0;
#pragma unroll
  for (unsigned w = 0; w < 8; ++w) {
    if (w < warp) woff += ws[w];
    agg += ws[w];
  }
  const unsigned long long tagE = (unsigned long long)(epoch & 0x3fffffffu) << 34;
  if (warp == 0) {
    unsigned prefix = 0;
    if (tile == 0) {
      if (lane == 0) st_status(status + 0, tagE | (2ull << 32) | agg);
    } else {
      if (lane == 0) st_status(status + tile, tagE | (1ull << 32) | agg);
      int look = (int)tile - 1;
      while (true) {
        const int idx = look - (int)lane;
        unsigned state = 2, val = 0;  // tiles before the first: inclusive prefix 0
        if (idx >= 0) {
          unsigned long long w;
          do {
            w = ld_status(status + idx);
          } while ((w >> 34) != (tagE >> 34) || ((w >> 32) & 3ull) == 0ull);
          state = (unsigned)((w >> 32) & 3ull);
          val = (unsigned)(w & 0xffffffffull);
        }
        const unsigned incl_mask = __ballot_sync(0xffffffffu, state == 2);
        const int first = incl_mask ? (__ffs(incl_mask) - 1) : 32;
        unsigned contrib = ((int)lane <= first) ? val : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
        prefix += contrib;
        if (incl_mask) break;
        look -= 32;
      }
      if (lane == 0) st_status(status + tile, tagE | (2ull << 32) | (prefix + agg));
    }
    if (lane == 0) {
      s_prefix = prefix;
      if (tile == ntiles - 1) *total = prefix + agg;
    }
  }
  __syncthreads();
  unsigned r = s_prefix + woff + incl - c;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (base + k < n) ranks[base + k] = r;
    r += f[k];
  }
}

cudaError_t scan_flags(ScanScratch sc, unsigned n_ub, const unsigned* n_dev, unsigned n_extra, unsigned* total,
                       cudaStream_t s) {
  if (n_ub == 0) return cudaMemsetAsync(total, 0, sizeof(unsigned), s);
  const unsigned nb = (n_ub + kScanTile - 1) / kScanTile;
  unsigned long long* status = reinterpret_cast<unsigned long long*>(sc.blockSums) + 1;
  unsigned* ticket = reinterpret_cast<unsigned*>(sc.blockSums);
  ScanHostState& h = *sc.host;
  h.epoch += 1;
  CFB_PDL(launch_pdl(scan_lookback_kernel, nb, 256, 0, s, sc.flags, n_ub, n_dev, n_extra, status, ticket, h.ticketBase, h.epoch, sc.ranks, total, nb));
  h.ticketBase += nb;
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------- a18 initialise
// vertex_feedback.vert/.geom for one depth image: record + flag per pixel, column-major ordinal.
__global__ void vertex_feedback_kernel(SurfelGeom g, const uint8_t* __restrict__ rgb, const float* __restrict__ depth,
                                       int time, float maxDepth, Surfel* __restrict__ rec, uint8_t* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= g.W || j >= g.H) return;
  const int W = g.W, H = g.H;
  const float inv_fx = 1.0f / g.fx, inv_fy = 1.0f / g.fy;  // FeedbackBuffer.cpp:81-82
  const float tcx = uv_coord(i, W), tcy = uv_coord(j, H);
  const float x = tcx * (float)W, y = tcy * (float)H;
  const int tx = texel(tcx, W), ty = texel(tcy, H);
  float3 v = get_vertex(depth, W, tx, ty, x, y, g, inv_fx, inv_fy);
  float3 n = get_normal_central(depth, g, tcx, tcy, x, y, inv_fx, inv_fy, v);
  const unsigned ord = (unsigned)i * H + j;
  const bool ok = !(v.z <= 0 || v.z > maxDepth);
  flags[ord] = ok ? 1 : 0;
  if (!ok) return;
  Surfel s;
  s.pos = make_float4(v.x, v.y, v.z, confidence(x, y, g.cx, g.cy, 1.0f));
  const uint8_t* c = rgb + (ty * W + tx) * 3;
  s.col = make_float4(encode_color(c[0] / 255.0f, c[1] / 255.0f, c[2] / 255.0f), 0.f, c[2] / 255.0f, (float)time);
  s.nrm = make_float4(n.x, n.y, n.z, get_radius(v.z, n.z, inv_fx, inv_fy));
  store_surfel(rec + ord, s);
}
// init_unstable.vert: surfel k takes position+colour of the k-th raw vertex and the normal+radius
// of the k-th filtered vertex (Model.cpp:230-241)
__global__ void init_scatter_kernel(unsigned n, const uint8_t* __restrict__ flags, const uint32_t* __restrict__ ranks,
                                    const Surfel* __restrict__ rec, Surfel* __restrict__ dst, unsigned capacity,
                                    int which) {
  unsigned o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n || !flags[o]) return;
  unsigned k = ranks[o];
  if (k >= capacity) return;
  Surfel s = load_surfel(rec + o);
  float4* d = reinterpret_cast<float4*>(dst + k);
  if (which == 0) {
    s.col.y = 0.f;  // init_unstable.vert:31-35
    s.col.z = 1.f;
    d[0] = s.pos;
    d[1] = s.col;
  } else {
    d[2] = s.nrm;
  }
}
__global__ void set_count_kernel(MapCounters* c, unsigned capacity) {
  c->count = min(c->scanTotal, capacity);
  c->unstableCount = 0;
}

// ------------------------------------------------------------------------------- a13 index map
__global__ void index_project_kernel(SurfelGeom g, const Surfel* __restrict__ surfels, unsigned n_ub,
                                     const MapCounters* __restrict__ ctr, PoseRef t_inv_ref, int time, float maxDepth,
                                     int timeDelta, unsigned long long* __restrict__ keys) {
  pdl_prologue();
  const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n_ub || id >= ctr->count) return;
  const Pose34 t_inv = resolve_pose(t_inv_ref);
  const float4 pos = __ldg(&surfels[id].pos);
  const float lastTime = __ldg(&surfels[id].col.w);
  float3 ph = xform_point(t_inv, make_float3(pos.x, pos.y, pos.z));
  if (ph.z > maxDepth || ph.z < 0 || (float)time - lastTime > (float)timeDelta) return;
  const float cols = (float)g.W, rows = (float)g.H;
  float xn = ((((g.fx * ph.x) / ph.z) + g.cx) - (cols * 0.5f)) / (cols * 0.5f);
  float yn = ((((g.fy * ph.y) / ph.z) + g.cy) - (rows * 0.5f)) / (rows * 0.5f);
  float zn = ph.z / maxDepth;
  if (!(xn >= -1.0f && xn <= 1.0f && yn >= -1.0f && yn <= 1.0f)) return;
  float xw = (xn + 1.0f) * (cols * 0.5f), yw = (yn + 1.0f) * (rows * 0.5f);
  int px = (int)floorf(xw), py = (int)floorf(yw);
  if (px < 0 || py < 0 || px >= g.W || py >= g.H) return;
  unsigned long long key = ((unsigned long long)depth_key24(zn * 0.5f + 0.5f) << 32) | id;
  atomicMin(&keys[py * g.W + px], key);
}
__global__ void index_resolve_kernel(SurfelGeom g, const Surfel* __restrict__ surfels, PoseRef t_inv_ref,
                                     unsigned long long* __restrict__ keys, IndexMaps out) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.W * g.H) return;
  const Pose34 t_inv = resolve_pose(t_inv_ref);
  const unsigned long long k = keys[i];
  keys[i] = ~0ull;  // the buffer is left as the next projection pass expects it (no memset between passes)
  if (k == ~0ull) {
    out.index[i] = 0;
    const float4 z = make_float4(0, 0, 0, 0);
    out.vertConf[i] = z;
    out.colorTime[i] = z;
    out.normRad[i] = z;
    return;
  }
  const unsigned id = (unsigned)(k & 0xffffffffu);
  Surfel s = load_surfel(surfels + id);
  float3 ph = xform_point(t_inv, make_float3(s.pos.x, s.pos.y, s.pos.z));
  float3 nl = normalize3(xform_vec(t_inv, make_float3(s.nrm.x, s.nrm.y, s.nrm.z)));
  out.index[i] = id;
  out.vertConf[i] = make_float4(ph.x, ph.y, ph.z, s.pos.w);
  out.colorTime[i] = s.col;
  out.normRad[i] = make_float4(nl.x, nl.y, nl.z, s.nrm.w);
}

// ------------------------------------------------------------------------------- a16 fuse
// data.vert for the eligible pixels only ((x%2,y%2) == (t%2,t%2), data.vert:116); e = column-major
// ordinal among eligible pixels (monotonic in the reference's draw order).
__global__ void fuse_associate_kernel(SurfelGeom g, PoseRef pose_ref, int time, const uint8_t* __restrict__ rgb,
                                      const uint8_t* __restrict__ mask, const float* __restrict__ depthRaw,
                                      const float* __restrict__ depthFiltered, float maxDepth, WeightRef weight_ref,
                                      unsigned maskID, IndexMaps idx, uint32_t* __restrict__ winner,
                                      Surfel* __restrict__ cand, uint32_t* __restrict__ candBest,
                                      uint8_t* __restrict__ flags, int par, int W2, int H2) {
  pdl_prologue();
  const int a = blockIdx.x * blockDim.x + threadIdx.x;  // eligible column index (fast: coalesced image reads)
  const int b = blockIdx.y * blockDim.y + threadIdx.y;  // eligible row index
  if (a >= W2 || b >= H2) return;
  const Pose34 pose = resolve_pose(pose_ref);
  const float weighting = weight_ref.dev ? __ldg(weight_ref.dev) * weight_ref.mult : weight_ref.v;
  const int i = 2 * a + par, j = 2 * b + par;
  const unsigned e = (unsigned)a * H2 + b;
  const int W = g.W, H = g.H;
  const float cols = (float)W, rows = (float)H, scale = 1.0f;
  const float inv_fx = (float)(1.0 / (double)g.fx), inv_fy = (float)(1.0 / (double)g.fy);  // Model.cpp:436-437
  const float ftime = (float)time;
  flags[e] = 0;
  if (i >= W || j >= H) return;
  const float tcx = uv_coord(i, W), tcy = uv_coord(j, H);
  const float x = tcx * cols, y = tcy * rows;
  const int tx = texel(tcx, W), ty = texel(tcy, H);
  if (!((int)x % 2 == (int)ftime % 2 && (int)y % 2 == (int)ftime % 2)) return;
  if ((unsigned)__ldg(mask + ty * W + tx) != maskID) return;
  if (__ldg(depthRaw + ty * W + texel(tcx - (1.0f / cols), W)) == 0) return;  // checkNeighbours, data.vert:52-71
  if (__ldg(depthRaw + texel(tcy - (1.0f / rows), H) * W + tx) == 0) return;
  if (__ldg(depthRaw + ty * W + texel(tcx + (1.0f / cols), W)) == 0) return;
  if (__ldg(depthRaw + texel(tcy + (1.0f / rows), H) * W + tx) == 0) return;
  const float3 vl = get_vertex(depthRaw, W, tx, ty, x, y, g, inv_fx, inv_fy);
  if (!(vl.z > 0 && vl.z <= maxDepth)) return;
  const float3 vg = xform_point(pose, vl);
  const float3 vf = get_vertex(depthFiltered, W, tx, ty, x, y, g, inv_fx, inv_fy);
  const float3 nl = get_normal_central(depthFiltered, g, tcx, tcy, x, y, inv_fx, inv_fy, vf);
  const float3 ng = xform_vec(pose, nl);
  Surfel c;
  c.pos = make_float4(vg.x, vg.y, vg.z, confidence(x, y, g.cx, g.cy, weighting));
  const uint8_t* col = rgb + (ty * W + tx) * 3;
  c.col = make_float4(encode_color(col[0] / 255.0f, col[1] / 255.0f, col[2] / 255.0f), 0.f, ftime, 0.f);
  c.nrm = make_float4(ng.x, ng.y, ng.z, get_radius(vf.z, nl.z, inv_fx, inv_fy));

  int operation = 0;
  uint32_t best = 0;
  const float indexXStep = (1.0f / (cols * scale)) * 0.5f, indexYStep = (1.0f / (rows * scale)) * 0.5f;
  float bestDist = 1000;
  const float windowMultiplier = 2;
  const float xl = (x - g.cx) * inv_fx, yl = (y - g.cy) * inv_fy;
  const float lambda = sqrtf(xl * xl + yl * yl + 1);
  const float3 ray = make_float3(xl, yl, 1);
  for (float si = tcx - (scale * indexXStep * windowMultiplier); si < tcx + (scale * indexXStep * windowMultiplier);
       si += indexXStep)
    for (float sj = tcy - (scale * indexYStep * windowMultiplier); sj < tcy + (scale * indexYStep * windowMultiplier);
         sj += indexYStep) {
      const int sp = texel(sj, H) * W + texel(si, W);
      const uint32_t current = __ldg(idx.index + sp);
      if (current > 0U) {
        const float4 vc = __ldg(idx.vertConf + sp);
        const float zdiff = vc.z - vl.z;
        if (fabsf(zdiff * lambda) < 0.05f) {
          const float3 cr = cross3(ray, make_float3(vc.x, vc.y, vc.z));
          const float dist = sqrtf(cr.x * cr.x + cr.y * cr.y + cr.z * cr.z);
          const float4 nr = __ldg(idx.normRad + sp);
          const float cosang = (nr.x * nl.x + nr.y * nl.y + nr.z * nl.z) /
                               (sqrtf(nr.x * nr.x + nr.y * nr.y + nr.z * nr.z) *
                                sqrtf(nl.x * nl.x + nl.y * nl.y + nl.z * nl.z));
          const bool angle_ok = (cosang <= 1.0f && cosang > COS_HALF);  // F5
          if (dist < bestDist && (fabsf(nr.z) < 0.75f || angle_ok)) {
            operation = 1;
            bestDist = dist;
            best = current;
          }
        }
      }
    }
  if (operation == 1) {
    c.col.w = -1.f;
    candBest[e] = best;
    atomicMin(&winner[best], e);  // first pixel in draw order wins the surfel (F3)
    flags[e] = 1;
  } else {
    c.col.w = -2.f;
    candBest[e] = 0;
    flags[e] = 2;
  }
  store_surfel(cand + e, c);
}
// After the scan: every candidate (flag 1 = associated, 2 = new) joins the unstable list at its rank (data.geom), and
// the winning pixel of each touched surfel updates it in place (update.vert).  One launch for both; the winner also
// puts its surfel's entry of `winner` back to all ones, so the array needs no clearing pass before the next fuse
// (a losing candidate that reads the entry after that still sees "not me").
__global__ void fuse_apply_kernel(unsigned n, int time, const uint8_t* __restrict__ flags, const uint32_t* __restrict__ ranks,
                                  const uint32_t* __restrict__ candBest, uint32_t* __restrict__ winner,
                                  const Surfel* __restrict__ cand, Surfel* __restrict__ unstable, Surfel* __restrict__ surfels,
                                  MapCounters* ctr) {
  pdl_prologue();
  unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e == 0) ctr->unstableCount = ctr->scanTotal;
  if (e >= n || !flags[e]) return;
  const Surfel nw = load_surfel(cand + e);
  store_surfel(unstable + ranks[e], nw);
  if (flags[e] != 1) return;
  const uint32_t id = candBest[e];
  if (winner[id] != e) return;
  Surfel o = load_surfel(surfels + id);
  const float c_k = o.pos.w, a = nw.pos.w, ftime = (float)time;
  if (nw.nrm.w < (1.0f + 0.5f) * o.nrm.w) {
    o.pos.x = ((c_k * o.pos.x) + (a * nw.pos.x)) / (c_k + a);
    o.pos.y = ((c_k * o.pos.y) + (a * nw.pos.y)) / (c_k + a);
    o.pos.z = ((c_k * o.pos.z) + (a * nw.pos.z)) / (c_k + a);
    o.pos.w = c_k + a;
    const float3 oc = decode_color(o.col.x), nc = decode_color(nw.col.x);
    o.col.x = encode_color(((c_k * oc.x) + (a * nc.x)) / (c_k + a), ((c_k * oc.y) + (a * nc.y)) / (c_k + a),
                           ((c_k * oc.z) + (a * nc.z)) / (c_k + a));
    o.col.w = ftime;
    float4 nr = make_float4(((c_k * o.nrm.x) + (a * nw.nrm.x)) / (c_k + a), ((c_k * o.nrm.y) + (a * nw.nrm.y)) / (c_k + a),
                            ((c_k * o.nrm.z) + (a * nw.nrm.z)) / (c_k + a), ((c_k * o.nrm.w) + (a * nw.nrm.w)) / (c_k + a));
    const float3 nn = normalize3(make_float3(nr.x, nr.y, nr.z));
    o.nrm = make_float4(nn.x, nn.y, nn.z, nr.w);
  } else {
    o.pos.w = c_k + a;
    o.col.w = ftime;
  }
  store_surfel(surfels + id, o);
  winner[id] = 0xffffffffu;
}

// ------------------------------------------------------------------------------- a17 clean
// copy_unstable.vert for item i (old surfels first, then candidates); the modified record is written
// back in place and its survival flag recorded for the stable compaction.
__global__ void clean_evaluate_kernel(SurfelGeom g, Surfel* __restrict__ src, Surfel* __restrict__ unstable,
                                      unsigned n_ub, const MapCounters* __restrict__ ctr, PoseRef t_inv_ref, int time,
                                      float confThreshold, int timeDelta, const float* __restrict__ depthFiltered,
                                      const uint8_t* __restrict__ mask, unsigned maskID, float outlierCoeff,
                                      IndexMaps idx, uint8_t* __restrict__ flags) {
  pdl_prologue();
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned count = ctr->count, total = count + ctr->unstableCount;
  if (i >= n_ub) return;
  if (i >= total) {
    flags[i] = 0;  // the scan runs over the host-side upper bound
    return;
  }
  const Pose34 t_inv = resolve_pose(t_inv_ref);
  Surfel* rec = (i < count) ? (src + i) : (unstable + (i - count));
  Surfel s = load_surfel(rec);
  const int W = g.W, H = g.H;
  const float cols = (float)W, rows = (float)H, scale = 1.0f;
  int test = 1;
  const float3 lp = xform_point(t_inv, make_float3(s.pos.x, s.pos.y, s.pos.z));
  const float x = ((g.fx * lp.x) / lp.z) + g.cx;
  const float y = ((g.fy * lp.y) / lp.z) + g.cy;
  const float3 ln = normalize3(xform_vec(t_inv, make_float3(s.nrm.x, s.nrm.y, s.nrm.z)));
  const float x_n = x / cols, y_n = y / rows;
  const float stepX = 1.0f / cols, stepY = 1.0f / rows;
  const float indexXStep = stepX * 0.5f / scale, indexYStep = stepY * 0.5f / scale;
  const float windowMultiplier = 2;
  int count_ = 0, zCount = 0, violationCount = 0;
  float avgViolation = 0;
  const float ftime = (float)time;
  if (ftime - s.col.w < (float)timeDelta && lp.z > 0 && x > 0 && y > 0 && x < cols && y < rows) {
    // The shader walks a (nominally 4x4) window with float loop counters; the trip counts are whatever
    // the float accumulation gives (4 or 5).  Enumerate the texels first, then issue the loads of all
    // of them together: 3 dependent round trips (index -> vertex -> colour/time) instead of up to 48.
    constexpr int kMaxTap = 5;  // 4 nominal steps, +1 when the accumulated counter lands just below the bound
    int tx[kMaxTap], ty[kMaxTap], nx = 0, ny = 0;
    for (float si = x_n - (scale * indexXStep * windowMultiplier); si < x_n + (scale * indexXStep * windowMultiplier);
         si += indexXStep)
      if (nx < kMaxTap) tx[nx++] = texel(si, W);
    for (float sj = y_n - (scale * indexYStep * windowMultiplier); sj < y_n + (scale * indexYStep * windowMultiplier);
         sj += indexYStep)
      if (ny < kMaxTap) ty[ny++] = texel(sj, H);
    // the counters advance by HALF a texel: consecutive taps often name the same texel.  A repeated
    // column / row reuses the values already loaded (it is still counted once per tap, as the shader does).
    bool dupx[kMaxTap], dupy[kMaxTap];
#pragma unroll
    for (int a = 0; a < kMaxTap; ++a) {
      dupx[a] = a > 0 && a < nx && tx[a] == tx[a - 1];
      dupy[a] = a > 0 && a < ny && ty[a] == ty[a - 1];
    }
    uint32_t cur[kMaxTap][kMaxTap];
#pragma unroll
    for (int a = 0; a < kMaxTap; ++a)
#pragma unroll
      for (int b = 0; b < kMaxTap; ++b)
        cur[a][b] = (a < nx && b < ny && !dupx[a] && !dupy[b]) ? __ldg(idx.index + ty[b] * W + tx[a]) : 0U;
#pragma unroll
    for (int a = 0; a < kMaxTap; ++a)
#pragma unroll
      for (int b = 0; b < kMaxTap; ++b) {
        if (b > 0 && dupy[b]) cur[a][b] = cur[a][b - 1];
        if (a > 0 && dupx[a]) cur[a][b] = cur[a - 1][b];
      }
    float4 vprev[kMaxTap];
#pragma unroll
    for (int a = 0; a < kMaxTap; ++a) {
      float4 vcs[kMaxTap];
#pragma unroll
      for (int b = 0; b < kMaxTap; ++b)
        vcs[b] = (cur[a][b] > 0U && !dupx[a] && !dupy[b]) ? __ldg(idx.vertConf + ty[b] * W + tx[a])
                                                         : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int b = 0; b < kMaxTap; ++b) {
        if (b > 0 && dupy[b]) vcs[b] = vcs[b - 1];
        if (a > 0 && dupx[a]) vcs[b] = vprev[b];
      }
#pragma unroll
      for (int b = 0; b < kMaxTap; ++b) {
        vprev[b] = vcs[b];
        const float4 vc = vcs[b];
        // both counters need a confident map surfel BEHIND this one: only then fetch colour/time
        if (a < nx && b < ny && cur[a][b] > 0U && vc.w > confThreshold && vc.z > lp.z) {
          const float4 ct = __ldg(idx.colorTime + ty[b] * W + tx[a]);
          const float ddx = vc.x - lp.x, ddy = vc.y - lp.y;
          if (ct.z < s.col.z && vc.z - lp.z < 0.01f && sqrtf(ddx * ddx + ddy * ddy) < s.nrm.w * 1.4f) count_++;
          if (ct.w == ftime && vc.z - lp.z > 0.01f && fabsf(ln.z) > 0.85f) zCount++;
        }
      }
    }
    float dd[3][3];
    int nsx = 0, nsy = 0, sx[3], sy[3];
    for (float si = x_n - stepX; si <= x_n + stepX; si += stepX)
      if (nsx < 3) sx[nsx++] = texel(si, W);
    for (float sj = y_n - stepY; sj <= y_n + stepY; sj += stepY)
      if (nsy < 3) sy[nsy++] = texel(sj, H);
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) dd[a][b] = (a < nsx && b < nsy) ? __ldg(depthFiltered + sy[b] * W + sx[a]) : 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a)  // same visiting order as the shader: si outer, sj inner (the float sum is ordered)
#pragma unroll
      for (int b = 0; b < 3; ++b)
        if (a < nsx && b < nsy) {
          const float d = dd[a][b] - lp.z;
          if (d > 0.03f) {
            violationCount++;
            avgViolation += d;
          }
        }
  }
  if (count_ > 8 || zCount > 4) test = 0;
  if (s.col.w == -2.f) s.col.w = ftime;
  if ((s.col.w == -1.f || ((ftime - s.col.w) > 20 && s.pos.w < confThreshold))) test = 0;
  if (s.col.w > 0 && ftime - s.col.w > (float)timeDelta) test = 1;
  if (violationCount > 0) {
    avgViolation /= (float)violationCount;
    s.pos.w *= 1.0f / (1 + outlierCoeff * avgViolation);
    const int sp = texel(y_n, H) * W + texel(x_n, W);
    const unsigned maskValue = __ldg(mask + sp);
    const float wDepth = __ldg(depthFiltered + sp);
    if (maskValue != maskID && (wDepth > lp.z - 0.05f && wDepth < lp.z + 0.05f))
      s.pos.w *= (0.5f + 0.5f * (1 - outlierCoeff / 10.0f));
  }
  flags[i] = (uint8_t)test;
  if (test) {  // only confidence and time stamp can change
    rec->pos.w = s.pos.w;
    rec->col.w = s.col.w;
  }
}
__global__ void clean_scatter_kernel(const Surfel* __restrict__ src, const Surfel* __restrict__ unstable,
                                     Surfel* __restrict__ dst, unsigned n_ub, unsigned capacity, MapCounters* ctr,
                                     const uint8_t* __restrict__ flags, const uint32_t* __restrict__ ranks, int time) {
  pdl_prologue();
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned count = ctr->count, total = count + ctr->unstableCount;
  if (i < n_ub && i < total && flags[i]) {
    const unsigned r = ranks[i];
    if (r < capacity) {
      const Surfel* rec = (i < count) ? (src + i) : (unstable + (i - count));
      store_surfel(dst + r, load_surfel(rec));
    }
  }
  // the last block to get here closes the pass (what a one-thread launch did before): every block has read the
  // counters by then
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&ctr->cleanTicket, 1u) == gridDim.x - 1) {
      ctr->count = min(ctr->scanTotal, capacity);
      ctr->unstableCount = 0;
      ctr->cleanTick = (unsigned)time;
      ctr->cleanTicket = 0;
    }
  }
}
__global__ void clean_finish_kernel(MapCounters* ctr, unsigned capacity, int time) {
  pdl_prologue();
  ctr->count = min(ctr->scanTotal, capacity);
  ctr->unstableCount = 0;
  ctr->cleanTick = (unsigned)time;
}

// ------------------------------------------------------------------------------- a14 splat
struct SplatVtx {
  float3 ph, nl;
  float conf, rad, colour, initTime, xw, yw, size;
};
__device__ __forceinline__ bool splat_vertex(const SurfelGeom& g, const Pose34& t_inv, const Surfel& s, float maxDepth,
                                             float confThreshold, int time, int maxTime, int timeDelta, SplatVtx& o) {
  const float cols = (float)g.W, rows = (float)g.H;
  o.ph = xform_point(t_inv, make_float3(s.pos.x, s.pos.y, s.pos.z));
  if (o.ph.z > maxDepth || o.ph.z < 0 || s.pos.w < confThreshold || (float)time - s.col.w > (float)timeDelta ||
      s.col.w > (float)maxTime)
    return false;
  const float xn = ((((g.fx * o.ph.x) / o.ph.z) + g.cx) - (cols * 0.5f)) / (cols * 0.5f);
  const float yn = ((((g.fy * o.ph.y) / o.ph.z) + g.cy) - (rows * 0.5f)) / (rows * 0.5f);
  if (!(xn >= -1.0f && xn <= 1.0f && yn >= -1.0f && yn <= 1.0f)) return false;
  o.xw = (xn + 1.0f) * (cols * 0.5f);
  o.yw = (yn + 1.0f) * (rows * 0.5f);
  o.conf = s.pos.w;
  o.colour = s.col.x;
  o.initTime = s.col.z;
  o.nl = normalize3(xform_vec(t_inv, make_float3(s.nrm.x, s.nrm.y, s.nrm.z)));
  o.rad = s.nrm.w;
  float3 x1 = normalize3(make_float3((o.nl.y - o.nl.z), -o.nl.x, o.nl.x));
  x1 = make_float3(x1.x * o.rad * 1.41421356f, x1.y * o.rad * 1.41421356f, x1.z * o.rad * 1.41421356f);
  const float3 y1 = cross3(o.nl, x1);
  float px[4], py[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float3 off = (q == 0) ? x1 : (q == 1) ? y1 : (q == 2) ? make_float3(-y1.x, -y1.y, -y1.z)
                                                               : make_float3(-x1.x, -x1.y, -x1.z);
    const float3 p = make_float3(o.ph.x + off.x, o.ph.y + off.y, o.ph.z + off.z);
    px[q] = ((g.fx * p.x) / p.z) + g.cx;
    py[q] = ((g.fy * p.y) / p.z) + g.cy;
  }
  const float xmin = fminf(px[0], fminf(px[1], fminf(px[2], px[3]))), xmax = fmaxf(px[0], fmaxf(px[1], fmaxf(px[2], px[3])));
  const float ymin = fminf(py[0], fminf(py[1], fminf(py[2], py[3]))), ymax = fmaxf(py[0], fmaxf(py[1], fmaxf(py[2], py[3])));
  o.size = fmaxf(0.f, fmaxf(fabsf(xmax - xmin), fabsf(ymax - ymin)));
  if (!(o.size >= 1.0f)) o.size = 1.0f;  // F4
  if (o.size > 2047.0f) o.size = 2047.0f;
  return true;
}
__device__ __forceinline__ bool splat_fragment(const SurfelGeom& g, const SplatVtx& v, int px, int py, float maxDepth,
                                               float4& vc, float& fragDepth) {
  const float fcx = (float)px + 0.5f, fcy = (float)py + 0.5f;
  const float3 l = normalize3(make_float3((fcx - g.cx) / g.fx, (fcy - g.cy) / g.fy, 1.0f));
  const float k = (v.ph.x * v.nl.x + v.ph.y * v.nl.y + v.ph.z * v.nl.z) / (l.x * v.nl.x + l.y * v.nl.y + l.z * v.nl.z);
  const float3 cp = make_float3(k * l.x, k * l.y, k * l.z);
  const float sqrRad = v.rad * v.rad;
  const float3 d = make_float3(cp.x - v.ph.x, cp.y - v.ph.y, cp.z - v.ph.z);
  if (!(d.x * d.x + d.y * d.y + d.z * d.z <= sqrRad)) return false;
  const float z = cp.z;
  vc = make_float4((fcx - g.cx) * z * (1.f / g.fx), (fcy - g.cy) * z * (1.f / g.fy), z, v.conf);
  fragDepth = (cp.z / (2 * maxDepth)) + 0.5f;
  return true;
}
__global__ void splat_raster_kernel(SurfelGeom g, const Surfel* __restrict__ surfels, unsigned n_ub,
                                    const MapCounters* __restrict__ ctr, PoseRef t_inv_ref, float maxDepth,
                                    float confThreshold, int time, int maxTime, int timeDelta,
                                    unsigned long long* __restrict__ keys) {
  pdl_prologue();
  const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n_ub || id >= ctr->count) return;
  const Pose34 t_inv = resolve_pose(t_inv_ref);
  SplatVtx v;
  const Surfel s = load_surfel(surfels + id);
  if (!splat_vertex(g, t_inv, s, maxDepth, confThreshold, time, maxTime, timeDelta, v)) return;
  const float h = v.size * 0.5f;
  int x0 = (int)ceilf(v.xw - h - 0.5f), x1 = (int)ceilf(v.xw + h - 0.5f) - 1;  // F4
  int y0 = (int)ceilf(v.yw - h - 0.5f), y1 = (int)ceilf(v.yw + h - 0.5f) - 1;
  x0 = max(x0, 0);
  y0 = max(y0, 0);
  x1 = min(x1, g.W - 1);
  y1 = min(y1, g.H - 1);
  for (int py = y0; py <= y1; ++py)
    for (int px = x0; px <= x1; ++px) {
      float4 vc;
      float fd;
      if (!splat_fragment(g, v, px, py, maxDepth, vc, fd)) continue;
      if (!(fd >= 0.0f && fd <= 1.0f)) continue;
      const unsigned long long key = ((unsigned long long)depth_key24(fd) << 32) | id;
      atomicMin(&keys[py * g.W + px], key);
    }
}
__global__ void splat_resolve_kernel(SurfelGeom g, const Surfel* __restrict__ surfels, PoseRef t_inv_ref, float maxDepth,
                                     float confThreshold, int time, int maxTime, int timeDelta,
                                     unsigned long long* __restrict__ keys, SplatMaps out) {
  pdl_prologue();
  const int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
  if (px >= g.W || py >= g.H) return;
  const Pose34 t_inv = resolve_pose(t_inv_ref);
  const int i = py * g.W + px;
  const unsigned long long k = keys[i];
  keys[i] = ~0ull;  // the buffer is left as the next projection pass expects it (no memset between passes)
  if (k == ~0ull) {
    out.image[i] = make_uchar4(0, 0, 0, 0);
    out.vertexConf[i] = make_float4(0, 0, 0, 0);
    out.normalRad[i] = make_float4(0, 0, 0, 0);
    out.time[i] = 0;
    return;
  }
  const unsigned id = (unsigned)(k & 0xffffffffu);
  SplatVtx v;
  splat_vertex(g, t_inv, load_surfel(surfels + id), maxDepth, confThreshold, time, maxTime, timeDelta, v);
  float4 vc;
  float fd;
  splat_fragment(g, v, px, py, maxDepth, vc, fd);
  const float3 col = decode_color(v.colour);
  out.image[i] = make_uchar4((unsigned char)floorf(col.x * 255.0f + 0.5f), (unsigned char)floorf(col.y * 255.0f + 0.5f),
                             (unsigned char)floorf(col.z * 255.0f + 0.5f), 255);
  out.vertexConf[i] = vc;
  out.normalRad[i] = make_float4(v.nl.x, v.nl.y, v.nl.z, v.rad);
  out.time[i] = (uint16_t)(unsigned)v.initTime;
}

// ------------------------------------------------------------------------------- a15 fill-in
__device__ __forceinline__ void fill_in_pixel(const SurfelGeom& g, const SplatMaps& splat, const uint8_t* __restrict__ rgb,
                                              const float* __restrict__ depth, int pt_geom, int pt_rgb, const FillMaps& out,
                                              int x, int y, int* s_cnt) {
  const int W = g.W, H = g.H, i = y * W + x;
  const float inv_fx = 1.0f / g.fx, inv_fy = 1.0f / g.fy;  // FillIn.cpp:73-74
  const float4 sv = splat.vertexConf[i];
  if (sv.z == 0 || pt_geom) {
    const float z = __ldg(depth + i);
    out.vertex[i] = make_float4(((float)x - g.cx) * z * inv_fx, ((float)y - g.cy) * z * inv_fy, z, 1.f);
  } else
    out.vertex[i] = sv;
  const float4 sn = splat.normalRad[i];
  if (sn.z == 0 || pt_geom) {
    const int xp = min(x + 1, W - 1), yp = min(y + 1, H - 1);
    const float z = __ldg(depth + i), zx = __ldg(depth + y * W + xp), zy = __ldg(depth + yp * W + x);
    const float3 v = make_float3(((float)x - g.cx) * z * inv_fx, ((float)y - g.cy) * z * inv_fy, z);
    const float3 vx = make_float3(((float)(x + 1) - g.cx) * zx * inv_fx, ((float)y - g.cy) * zx * inv_fy, zx);
    const float3 vy = make_float3(((float)x - g.cx) * zy * inv_fx, ((float)(y + 1) - g.cy) * zy * inv_fy, zy);
    const float3 n = normalize3(cross3(make_float3(vx.x - v.x, vx.y - v.y, vx.z - v.z),
                                       make_float3(vy.x - v.x, vy.y - v.y, vy.z - v.z)));
    out.normal[i] = make_float4(n.x, n.y, n.z, 1.f);
  } else
    out.normal[i] = sn;
  const uchar4 si = splat.image[i];
  if ((si.x == 0 && si.y == 0 && si.z == 0) || pt_rgb)
    out.image[i] = make_uchar4(rgb[i * 3], rgb[i * 3 + 1], rgb[i * 3 + 2], 255);
  else
    out.image[i] = si;
  {  // is this pixel one of requiresFillIn's samples?  (sample q of a row of lw sits at texel((q + 0.5) / lw))
    const int cons = 20, lw = W / cons, lh = H / cons;
    if (lw > 0 && lh > 0) {
      const int qi = min(x * lw / W, lw - 1), qj = min(y * lh / H, lh - 1);
      if (texel(((float)qi + 0.5f) / (float)lw, W) == x && texel(((float)qj + 0.5f) / (float)lh, H) == y &&
          si.x > 0 && si.y > 0 && si.z > 0)
        atomicAdd(s_cnt, 1);
    }
  }
}
// CoFusion::requiresFillIn (CoFusion.cpp:547-565) rides along: the 20-pixel sample grid of the splat image is counted
// by the threads that own those pixels, the last block forms the ratio (a 1-CTA launch before).
__global__ void fill_in_kernel(SurfelGeom g, SplatMaps splat, const uint8_t* __restrict__ rgb,
                               const float* __restrict__ depth, int pt_geom, int pt_rgb, FillMaps out, MapCounters* ctr,
                               float ratio) {
  pdl_prologue();
  __shared__ int s_cnt;
  if (threadIdx.x == 0 && threadIdx.y == 0) s_cnt = 0;
  __syncthreads();
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x < g.W && y < g.H) fill_in_pixel(g, splat, rgb, depth, pt_geom, pt_rgb, out, x, y, &s_cnt);
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    if (s_cnt) atomicAdd(&ctr->fillSamples, (unsigned)s_cnt);
    __threadfence();
    if (atomicAdd(&ctr->fillTicket, 1u) == gridDim.x * gridDim.y - 1) {
      const int cons = 20, lw = g.W / cons, lh = g.H / cons;
      const unsigned total = atomicExch(&ctr->fillSamples, 0u);
      ctr->fillInRequired = ((float)total / (float)(lh * lw) < ratio) ? 1u : 0u;
      ctr->fillTicket = 0;
    }
  }
}

inline unsigned cdiv(unsigned a, unsigned b) { return (a + b - 1) / b; }

}  // namespace

#define RET_IF(e)                       \
  do {                                  \
    cudaError_t e__ = (e);              \
    if (e__ != cudaSuccess) return e__; \
  } while (0)

cudaError_t launch_surfel_initialise(const SurfelGeom& g, const uint8_t* rgb, const float* depthRaw,
                                     const float* depthFiltered, int time, float maxDepth, Surfel* dst,
                                     unsigned capacity, Surfel* stagingRaw, Surfel* stagingFil, ScanScratch sc,
                                     MapCounters* ctr, cudaStream_t s) {
  const unsigned n = (unsigned)g.W * g.H;
  const dim3 b(32, 8), gr(cdiv(g.W, 32), cdiv(g.H, 8));
  // filtered pass first so that the final scanTotal (= surfel count) is the RAW one
  vertex_feedback_kernel<<<gr, b, 0, s>>>(g, rgb, depthFiltered, time, maxDepth, stagingFil, sc.flags);
  RET_IF(scan_flags(sc, n, nullptr, 0, &ctr->scanTotal, s));
  init_scatter_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, sc.flags, sc.ranks, stagingFil, dst, capacity, 1);
  vertex_feedback_kernel<<<gr, b, 0, s>>>(g, rgb, depthRaw, time, maxDepth, stagingRaw, sc.flags);
  RET_IF(scan_flags(sc, n, nullptr, 0, &ctr->scanTotal, s));
  init_scatter_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, sc.flags, sc.ranks, stagingRaw, dst, capacity, 0);
  set_count_kernel<<<1, 1, 0, s>>>(ctr, capacity);
  return cudaGetLastError();
}

cudaError_t launch_predict_indices(const SurfelGeom& g, const Surfel* surfels, unsigned count_ub,
                                   const MapCounters* ctr, const PoseRef& t_inv, int time, float maxDepth,
                                   int timeDelta, unsigned long long* keys, IndexMaps out, cudaStream_t s) {
  const unsigned n = (unsigned)g.W * g.H;
  // keys: all ones on entry (Model construction, then every resolve pass restores it)
  if (count_ub)
    CFB_PDL(launch_pdl(index_project_kernel, cdiv(count_ub, 256), 256, 0, s, g, surfels, count_ub, ctr, t_inv, time, maxDepth,
                                                             timeDelta, keys));
  CFB_PDL(launch_pdl(index_resolve_kernel, cdiv(n, 256), 256, 0, s, g, surfels, t_inv, keys, out));
  return cudaGetLastError();
}

cudaError_t launch_fuse(const SurfelGeom& g, Surfel* surfels, unsigned count_ub, MapCounters* ctr,
                        const PoseRef& pose, int time, const uint8_t* rgb, const uint8_t* mask, const float* depthRaw,
                        const float* depthFiltered, float maxDepth, const WeightRef& weighting, unsigned maskID, IndexMaps idx,
                        uint32_t* winner, Surfel* cand, uint32_t* candBest, Surfel* unstable, ScanScratch sc,
                        cudaStream_t s) {
  const int par = ((time % 2) + 2) % 2;
  const int W2 = (g.W - par + 1) / 2, H2 = (g.H - par + 1) / 2;
  const unsigned ne = (unsigned)W2 * H2;
  // `winner` is all ones on entry (Model construction; the winners of the last fuse restored their entries)
  const dim3 b(32, 8), gr(cdiv(W2, 32), cdiv(H2, 8));
  CFB_PDL(launch_pdl(fuse_associate_kernel, gr, b, 0, s, g, pose, time, rgb, mask, depthRaw, depthFiltered, maxDepth, weighting, maskID,
                                         idx, winner, cand, candBest, sc.flags, par, W2, H2));
  RET_IF(scan_flags(sc, ne, nullptr, 0, &ctr->scanTotal, s));
  CFB_PDL(launch_pdl(fuse_apply_kernel, cdiv(ne, 256), 256, 0, s, ne, time, sc.flags, sc.ranks, candBest, winner, cand, unstable, surfels, ctr));
  return cudaGetLastError();
}

cudaError_t launch_clean(const SurfelGeom& g, Surfel* src, Surfel* unstable, Surfel* dst, unsigned count_ub,
                         unsigned cand_ub, unsigned capacity, MapCounters* ctr, const PoseRef& t_inv, int time,
                         float confThreshold, int timeDelta, const float* depthFiltered, const uint8_t* mask,
                         unsigned maskID, float outlierCoeff, IndexMaps idx, ScanScratch sc, cudaStream_t s) {
  const unsigned n_ub = count_ub + cand_ub;
  if (n_ub) {
    CFB_PDL(launch_pdl(clean_evaluate_kernel, cdiv(n_ub, 128), 128, 0, s, g, src, unstable, n_ub, ctr, t_inv, time, confThreshold,
                                                          timeDelta, depthFiltered, mask, maskID, outlierCoeff, idx,
                                                          sc.flags));
    RET_IF(scan_flags(sc, n_ub, &ctr->count, cand_ub, &ctr->scanTotal, s));
    CFB_PDL(launch_pdl(clean_scatter_kernel, cdiv(n_ub, 256), 256, 0, s, src, unstable, dst, n_ub, capacity, ctr, sc.flags, sc.ranks, time));
  } else {
    RET_IF(cudaMemsetAsync(&ctr->scanTotal, 0, sizeof(unsigned), s));
    CFB_PDL(launch_pdl(clean_finish_kernel, 1, 1, 0, s, ctr, capacity, time));
  }
  return cudaGetLastError();
}

cudaError_t launch_combined_predict(const SurfelGeom& g, const Surfel* surfels, unsigned count_ub, MapCounters* ctr,
                                    const PoseRef& t_inv, float maxDepth, float confThreshold, int time, int maxTime,
                                    int timeDelta, unsigned long long* keys, SplatMaps out, cudaStream_t s) {
  if (count_ub)
    CFB_PDL(launch_pdl(splat_raster_kernel, cdiv(count_ub, 128), 128, 0, s, g, surfels, count_ub, ctr, t_inv, maxDepth, confThreshold,
                                                            time, maxTime, timeDelta, keys));
  const dim3 b(32, 8), gr(cdiv(g.W, 32), cdiv(g.H, 8));
  CFB_PDL(launch_pdl(splat_resolve_kernel, gr, b, 0, s, g, surfels, t_inv, maxDepth, confThreshold, time, maxTime, timeDelta, keys, out));
  return cudaGetLastError();
}


cudaError_t launch_fill_in(const SurfelGeom& g, SplatMaps splat, const uint8_t* rgb, const float* depthFiltered,
                           int pt_geom, int pt_rgb, FillMaps out, MapCounters* ctr, float ratio, cudaStream_t s) {
  const dim3 b(32, 8), gr(cdiv(g.W, 32), cdiv(g.H, 8));
  CFB_PDL(launch_pdl(fill_in_kernel, gr, b, 0, s, g, splat, rgb, depthFiltered, pt_geom, pt_rgb, out, ctr, ratio));
  return cudaGetLastError();
}

}  // namespace cfb
