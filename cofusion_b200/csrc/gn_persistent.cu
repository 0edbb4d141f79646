// gn_persistent.cu -- the whole tracker optimisation (SO(3) pre-alignment + 3-level ICP/RGB
// Gauss-Newton) as ONE persistent cooperative kernel: one CTA per SM, grid-wide barriers in
// software, the FP64 Gauss-Newton step executed redundantly by every CTA on its own copy of the state.
//
// Why (measured on B200, profiles/): at 640x480 a GN iteration touches <= 34 MB that sits in the
// 126 MB L2, so the work of an iteration is a few microseconds while every separate reduction kernel
// costs ~9 us of launch + prologue + last-block election latency; the 48-launch graph version spends
// two thirds of its 0.58 ms in those fixed costs.  Inside one kernel an iteration is
//     residual -> arrive A -> ICP rows -> wait A -> RGB rows -> barrier B -> fold partial rows -> solve
// with the iteration-invariant inputs of a thread's pixels staged in shared memory once per level and
// the photometric correspondences of a pixel never leaving the SM (the reference round-trips a
// 16-byte DataTerm image through memory, reduce.cu:862 / :524).
//
// Arithmetic is the same per-pixel code as the stand-alone steps (tracker_device.cuh); sums are
// folded in a fixed order (per-CTA partials, then CTA order) -> bit-reproducible run to run.
// gn_batched.cu is the same kernel over several models of one frame.
#include "gn_serial.cuh"
#include "image_kernels.cuh"

namespace cfb {
namespace {
using namespace dev;

constexpr int kPT = 512;   // threads per CTA, one CTA per SM

struct LevelData {
  const float *vmap_curr, *nmap_curr, *vmap_g_prev, *nmap_g_prev;
  const float *lastDepth, *nextDepth;
  const unsigned char *lastImage, *nextImage;
  const short *dIdx, *dIdy;
  const unsigned char* cand;
  DataTerm* corres;  // only used when a thread owns more than kMaxPP pixels
  int w, h;
  LevelK k;
};

struct GridSync {   // zeroed by rgb_prepare_all_kernel before every launch
  unsigned arrive;  // monotonic arrival counter of the software grid barrier
  unsigned pad[31];
  int counts[32][2];  // per GN iteration: {RGB correspondences, sum of floor(diff^2)} (integer atomics)
};

struct PersistParams {
  LevelData L[3];
  const unsigned char *so3_last, *so3_next;
  GNState* g;   // global copy of the final state (pose, stats), written by CTA 0
  StepScratch* sc;
  GridSync* gs;
  const float* pose_in;
  float* err;
  size_t err_pitch;
  float distThres, angleThres, maxDepthDelta, sobelScale, icpWeight;
  int use_so3;
  int iters[3];
  unsigned long long* dbg;  // optional %globaltimer trace of CTA 0, tools only
};

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define DBG_MARK(slot)                                                       \
  do {                                                                       \
    if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) p.dbg[(slot)] = gtime(); \
  } while (0)

// All shared state lives in ONE dynamic allocation with a fixed layout, so that the (deliberately
// not inlined: one copy of the FP64 solve instead of five) phase functions can re-derive typed
// shared-memory references from `extern __shared__` -- LDS/STS with immediate offsets instead of
// generic loads through pointers, and no local-memory copy of the kernel parameters.
enum StagePlane { SP_VX, SP_VY, SP_VZ, SP_NX, SP_NY, SP_NZ, SP_D1, SP_SOB, SP_FLAGS, SP_ZERO, SP_DIFF, SP_D0, SP_COUNT };
constexpr int kStagePP = 5;  // 640x480 on 148 x 512 threads: ceil(307200 / 75776)
constexpr unsigned kNoCorr = 0xffffffffu;
struct Smem {
  float stage[SP_COUNT * kStagePP * kPT];
  float red[2 * (kPT / 32) * 32];
  float out64[64];
  GNState S;  // every CTA keeps (and identically updates) its own copy of the GN state
  int scnt[kPT / 32], ssig[kPT / 32];
  PersistParams prm;
  int sched[20];
  int nsched;
};
#define SMEM_REF()                                                \
  extern __shared__ __align__(16) unsigned char dyn_smem_raw[]; \
  Smem& sm = *reinterpret_cast<Smem*>(dyn_smem_raw);              \
  const PersistParams& p = sm.prm

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Split software grid barrier on a monotonic counter: arrive() publishes this CTA's global writes,
// wait(target) blocks until `target` arrivals have been counted.  Work placed between the two calls
// hides the barrier latency.  No CTA is special: every CTA folds the per-CTA partial sums itself (in
// the same fixed order -> identical totals everywhere) and runs the FP64 Gauss-Newton step on its own
// shared-memory copy of the state, so there is no "finaliser -> flag -> everyone re-reads" hop.
// (Measured alternative: relaxed polling + one acq_rel fence after the wait is SLOWER -- the fence is
// a full MEMBAR.ALL.GPU, +0.6 us per barrier -- than polling with acquire loads.)
__device__ __forceinline__ void grid_arrive(GridSync* gs) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(&gs->arrive, 1u);
  }
}
__device__ __forceinline__ void grid_wait(GridSync* gs, unsigned target) {
  if (threadIdx.x == 0) {
    while (ld_acquire(&gs->arrive) < target) __nanosleep(20);
  }
  __syncthreads();
}

// fixed-order sum of `nsets` consecutive sets of per-CTA partial rows -> out[set*32 + lane] (shared)
template <int NSETS>
__device__ __forceinline__ void sum_partials(const float* partials, unsigned rows_per_set, float* smem, float* out) {
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int set = 0; set < NSETS; ++set) {
    const float* base = partials + (size_t)set * rows_per_set * 32;
    float v[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) {  // all loads in flight before the first add
      const unsigned b = warp + k * nw;
      v[k] = (b < gridDim.x) ? __ldcg(&base[b * 32 + lane]) : 0.f;
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 12; ++k) s += v[k];
    for (unsigned b = warp + 12 * nw; b < gridDim.x; b += nw) s += __ldcg(&base[b * 32 + lane]);
    smem[(set * (kPT / 32) + warp) * 32 + lane] = s;
  }
  __syncthreads();
  if (warp < NSETS) {
    float tot = 0.f;
    for (unsigned w = 0; w < nw; ++w) tot += smem[(warp * (kPT / 32) + w) * 32 + lane];
    out[warp * 32 + lane] = tot;
  }
  __syncthreads();
}

// RGB Jacobian row with the cloud point recomputed from lastDepth (same expression as
// projectPointsKernel, cudafuncs.cu:731-735 -> identical values, 8 bytes less traffic per row)
__device__ __forceinline__ void rgb_step_from_depth(const LevelData& L, float sigma, float sobelScale, bool valid,
                                                    unsigned zero, float diff, int x, int y, float (&acc)[32]) {
  float row[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (valid) {
    float w = sigma + fabsf(diff);
    w = w > 1.19209290E-07F ? 1.0f / w : 1.0f;
    if (sigma == -1.f) w = 1.f;
    row[6] = -w * diff;
    const int zx = (int)(zero & 0xffff), zy = (int)(zero >> 16);
    const float z = __ldg(L.lastDepth + zy * L.w + zx);
    const float invFx = 1.0f / L.k.fx, invFy = 1.0f / L.k.fy;
    float3 P = make_float3(((float)zx - L.k.cx) * z * invFx, ((float)zy - L.k.cy) * z * invFy, z);
    float invz = (float)(1.0 / (double)P.z);
    float dI_dx_val = w * sobelScale * (float)__ldg(L.dIdx + y * L.w + x);
    float dI_dy_val = w * sobelScale * (float)__ldg(L.dIdy + y * L.w + x);
    float v0 = dI_dx_val * L.k.fx * invz;
    float v1 = dI_dy_val * L.k.fy * invz;
    float v2 = -(v0 * P.x + v1 * P.y) * invz;
    row[0] = v0;
    row[1] = v1;
    row[2] = v2;
    row[3] = -P.z * v1 + P.y * v2;
    row[4] = P.z * v0 - P.x * v2;
    row[5] = -P.y * v0 + P.x * v1;
  }
  accumulate_se3(acc, row, valid);
}

// block reduce that lets warps without any contribution skip the 31-shuffle transpose
__device__ __forceinline__ float block_reduce32_sparse(float (&v)[32], bool warp_has_work, float* smem) {
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float t = 0.f;
  if (warp_has_work) t = warp_transpose_reduce32(v);
  smem[warp * 32 + lane] = t;
  __syncthreads();
  float s = 0.f;
  if (warp == 0)
    for (unsigned w = 0; w < nw; ++w) s += smem[w * 32 + lane];
  return s;
}

// ---------------------------------------------------------------------------------------------
// Per-level staging.  The pixels a thread owns (tid + k * nthreads) do not change within a level, and
// half of what an iteration reads about them does not change either: the current-frame vertex and
// normal, nextDepth / nextImage / Sobel of the photometric term, the candidate gate.  They are loaded
// ONCE per level into shared memory (plane-major [plane][k][thread], conflict free); an iteration
// then only issues the pose-dependent gathers (model vertex / normal, lastDepth / lastImage), all of a
// thread's pixels in flight together, and the RGB Jacobian pass touches no global memory at all.
struct LevelCtx {  // everything the phases of one level need, built once per level
  IcpArgs ia;
  RgbResidualArgs ra;
};

__device__ __forceinline__ void make_level_ctx(const PersistParams& p, const LevelData& L, LevelCtx& c) {
  const size_t pitch = (size_t)L.w * 4;
  c.ia.vmap_curr = {L.vmap_curr, pitch};
  c.ia.nmap_curr = {L.nmap_curr, pitch};
  c.ia.vmap_g_prev = {L.vmap_g_prev, pitch};
  c.ia.nmap_g_prev = {L.nmap_g_prev, pitch};
  c.ia.intr = Intr{L.k.fx, L.k.fy, L.k.cx, L.k.cy};
  c.ia.distThres = p.distThres;
  c.ia.angleThres = p.angleThres;
  c.ia.cols = L.w;
  c.ia.rows = L.h;
  c.ia.error_map = nullptr;
  c.ia.error_pitch = p.err_pitch;
  c.ra.minScale = 0.f;
  c.ra.maxDepthDelta = p.maxDepthDelta;
  c.ra.dIdx = L.dIdx;
  c.ra.dIdy = L.dIdy;
  c.ra.grad_pitch = (size_t)L.w * 2;
  c.ra.lastDepth = L.lastDepth;
  c.ra.nextDepth = L.nextDepth;
  c.ra.depth_pitch = pitch;
  c.ra.lastImage = L.lastImage;
  c.ra.nextImage = L.nextImage;
  c.ra.img_pitch = (size_t)L.w;
  c.ra.corres = L.corres;
  c.ra.cols = L.w;
  c.ra.rows = L.h;
}

// the common tail of a GN iteration: barrier B, fixed-order fold of the partial rows, FP64 solve
__device__ __noinline__ void finish_iteration(unsigned& barriers, int q, float tmpError, int tot_cnt) {
  SMEM_REF();
  GridSync* gs = p.gs;
  StepScratch* sc = p.sc;
  const int nsched = sm.nsched;
  const int* sched = sm.sched;
  const unsigned G = gridDim.x;
  grid_arrive(gs);  // barrier B
  grid_wait(gs, ++barriers * G);
  DBG_MARK(8 + q * 8 + 5);
  const float* rows = sc->partials + (size_t)(((q & 1) * 2) * G) * 32;
  sum_partials<2>(rows, G, sm.red, sm.out64);
  DBG_MARK(8 + q * 8 + 6);
  const int is_last = (q + 1 == nsched);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < 32; ++i) sm.S.icp_result[i] = sm.out64[i];
    gn_solve_serial(&sm.S, nullptr, sm.out64 + 32, p.icpWeight, p.L[is_last ? sched[q] : sched[q + 1]].k, is_last,
                    tmpError, tot_cnt);
  }
  __syncthreads();
  DBG_MARK(8 + q * 8 + 7);
}

// publish the integer photometric count / sigma of this CTA, then arrive at barrier A
__device__ __forceinline__ void publish_counts(int q, int cnt, int sig) {
  SMEM_REF();
  GridSync* gs = p.gs;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    sig += __shfl_xor_sync(0xffffffffu, sig, o);
  }
  if (lane == 0) {
    sm.scnt[warp] = cnt;
    sm.ssig[warp] = sig;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int c = 0, s = 0;
    for (int w = 0; w < kPT / 32; ++w) {
      c += sm.scnt[w];
      s += sm.ssig[w];
    }
    atomicAdd(&gs->counts[q][0], c);  // integer sums commute exactly
    atomicAdd(&gs->counts[q][1], s);
  }
  grid_arrive(gs);  // barrier A: its latency is hidden behind the ICP pass
}

__device__ __forceinline__ float fetch_sigma(int q, float* tmpError, int* tot) {
  SMEM_REF();
  GridSync* gs = p.gs;
  if (threadIdx.x == 0) {
    sm.scnt[0] = __ldcg(&gs->counts[q][0]);
    sm.ssig[0] = __ldcg(&gs->counts[q][1]);
  }
  __syncthreads();
  const int tot_cnt = sm.scnt[0], tot_sig = sm.ssig[0];
  *tot = tot_cnt;
  return rgb_sigma_from_counts(tot_cnt, tot_sig, tmpError);
}

template <int PP>
__device__ __noinline__ void run_level_staged(int lvl, int q0, int nit, unsigned& barriers) {
  SMEM_REF();
  GridSync* gs = p.gs;
  StepScratch* sc = p.sc;
  const LevelData& L = p.L[lvl];
  const int N = L.w * L.h, W = L.w;
  const int tid = blockIdx.x * kPT + threadIdx.x, nthreads = gridDim.x * kPT;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, G = gridDim.x;
  auto at = [&](int plane, int k) -> float& { return sm.stage[(plane * PP + k) * kPT + threadIdx.x]; };
  auto prow = [&](int set, int parity) { return sc->partials + (size_t)((parity * 2 + set) * G) * 32; };
  LevelCtx c;
  make_level_ctx(p, L, c);
  const GNState& S = sm.S;

  // ---- stage the iteration-invariant inputs of this thread's pixels
  bool any_px = false;
#pragma unroll
  for (int k = 0; k < PP; ++k) {
    const int px = tid + k * nthreads;
    unsigned flags = 0;
    if (px < N) {
      any_px = true;
      const int y = px / W, x = px - y * W;
      at(SP_VX, k) = ldplane(c.ia.vmap_curr, y, x);
      at(SP_VY, k) = ldplane(c.ia.vmap_curr, y + L.h, x);
      at(SP_VZ, k) = ldplane(c.ia.vmap_curr, y + 2 * L.h, x);
      at(SP_NX, k) = ldplane(c.ia.nmap_curr, y, x);
      at(SP_NY, k) = ldplane(c.ia.nmap_curr, y + L.h, x);
      at(SP_NZ, k) = ldplane(c.ia.nmap_curr, y + 2 * L.h, x);
      at(SP_D1, k) = __ldg(L.nextDepth + px);
      const unsigned sob = (unsigned)(unsigned short)__ldg(L.dIdx + px) | ((unsigned)(unsigned short)__ldg(L.dIdy + px) << 16);
      at(SP_SOB, k) = __uint_as_float(sob);
      flags = 0x10000u | (__ldg(L.cand + px) ? 1u : 0u) | ((unsigned)__ldg(L.nextImage + px) << 8);
    }
    at(SP_FLAGS, k) = __uint_as_float(flags);
  }
  const bool warp_work = __any_sync(0xffffffffu, any_px);

  for (int it = 0; it < nit; ++it) {
    const int q = q0 + it;
    DBG_MARK(8 + q * 8 + 0);
    const IcpPose& P = S.pose;
    const RgbWarp& Wp = S.warp;
    const bool last_of_l0 = (lvl == 0 && it + 1 == nit);
    float* const error_map = last_of_l0 ? p.err : nullptr;

    // -------- phase 1: photometric correspondences; all gathers of the thread in flight together
    int cnt = 0, sig = 0;
    {
      int u0[PP], v0[PP];
      float td1[PP];
      bool ok[PP];
#pragma unroll
      for (int k = 0; k < PP; ++k) {
        const unsigned flags = __float_as_uint(at(SP_FLAGS, k));
        ok[k] = (flags & 1u) != 0;
        u0[k] = v0[k] = 0;
        td1[k] = 0.f;
        if (ok[k]) {
          const int px = tid + k * nthreads, y = px / W, x = px - y * W;
          const float d1 = at(SP_D1, k);
          const float* kk = Wp.krkinv.m;
          td1[k] = d1 * (kk[6] * x + kk[7] * y + kk[8]) + Wp.kt[2];
          u0[k] = __float2int_rn((d1 * (kk[0] * x + kk[1] * y + kk[2]) + Wp.kt[0]) / td1[k]);
          v0[k] = __float2int_rn((d1 * (kk[3] * x + kk[4] * y + kk[5]) + Wp.kt[1]) / td1[k]);
          ok[k] = (u0[k] >= 0 && v0[k] >= 0 && u0[k] < W && v0[k] < L.h);
        }
      }
      float d0[PP];
      unsigned char li[PP];
#pragma unroll
      for (int k = 0; k < PP; ++k) {
        d0[k] = ok[k] ? __ldg(L.lastDepth + v0[k] * W + u0[k]) : 0.f;
        li[k] = ok[k] ? __ldg(L.lastImage + v0[k] * W + u0[k]) : (unsigned char)0;
      }
#pragma unroll
      for (int k = 0; k < PP; ++k) {
        unsigned zero = kNoCorr;
        if (ok[k] && d0[k] > 0 && fabsf(td1[k] - d0[k]) <= p.maxDepthDelta && li[k] != 0) {
          const unsigned flags = __float_as_uint(at(SP_FLAGS, k));
          const float diff = (float)((flags >> 8) & 0xffu) - (float)li[k];
          cnt += 1;
          sig += (int)(diff * diff);  // float -> int truncation, reduce.cu:851
          zero = (unsigned)(unsigned short)u0[k] | ((unsigned)(unsigned short)v0[k] << 16);
          at(SP_DIFF, k) = diff;
          at(SP_D0, k) = d0[k];
        }
        at(SP_ZERO, k) = __uint_as_float(zero);
      }
    }
    publish_counts(q, cnt, sig);
    DBG_MARK(8 + q * 8 + 1);

    // -------- phase 2: ICP rows (independent of the count)
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
    {
      const float3 tcurr = make_float3(P.tcurr[0], P.tcurr[1], P.tcurr[2]);
      const float3 tprev = make_float3(P.tprev[0], P.tprev[1], P.tprev[2]);
      int ux[PP], uy[PP];
      bool ok[PP];
#pragma unroll
      for (int k = 0; k < PP; ++k) {
        const int px = tid + k * nthreads;
        ok[k] = false;
        ux[k] = uy[k] = 0;
        if (px < N) {
          const float3 vcurr = make_float3(at(SP_VX, k), at(SP_VY, k), at(SP_VZ, k));
          const float3 vcurr_g = mul(P.Rcurr, vcurr) + tcurr;
          const float3 vcurr_cp = mul(P.Rprev_inv, vcurr_g - tprev);
          ux[k] = __float2int_rn(vcurr_cp.x * c.ia.intr.fx / vcurr_cp.z + c.ia.intr.cx);
          uy[k] = __float2int_rn(vcurr_cp.y * c.ia.intr.fy / vcurr_cp.z + c.ia.intr.cy);
          ok[k] = !(ux[k] < 0 || uy[k] < 0 || ux[k] >= W || uy[k] >= L.h || vcurr_cp.z < 0);
          if (!ok[k] && error_map) {
            const int y = px / W, x = px - y * W;
            row_ptr(error_map, p.err_pitch, y)[x] = 0.0f;
          }
        }
      }
      float3 vp[PP], np[PP];
#pragma unroll
      for (int k = 0; k < PP; ++k) {
        vp[k] = np[k] = make_float3(0.f, 0.f, 0.f);
        if (ok[k]) {
          vp[k] = make_float3(ldplane(c.ia.vmap_g_prev, uy[k], ux[k]), ldplane(c.ia.vmap_g_prev, uy[k] + L.h, ux[k]),
                              ldplane(c.ia.vmap_g_prev, uy[k] + 2 * L.h, ux[k]));
          np[k] = make_float3(ldplane(c.ia.nmap_g_prev, uy[k], ux[k]), ldplane(c.ia.nmap_g_prev, uy[k] + L.h, ux[k]),
                              ldplane(c.ia.nmap_g_prev, uy[k] + 2 * L.h, ux[k]));
        }
      }
#pragma unroll
      for (int k = 0; k < PP; ++k) {
        if (!ok[k]) continue;
        const float3 vcurr = make_float3(at(SP_VX, k), at(SP_VY, k), at(SP_VZ, k));
        const float3 ncurr = make_float3(at(SP_NX, k), at(SP_NY, k), at(SP_NZ, k));
        const float3 vcurr_g = mul(P.Rcurr, vcurr) + tcurr;
        const float3 vcurr_cp = mul(P.Rprev_inv, vcurr_g - tprev);
        const float3 ncurr_g = mul(P.Rcurr, ncurr);
        const float dist = norm(vp[k] - vcurr_g);
        const float sine = norm(cross(ncurr_g, np[k]));
        if (error_map) {
          const int px = tid + k * nthreads, y = px / W, x = px - y * W;
          row_ptr(error_map, p.err_pitch, y)[x] = isfinite(dist) ? dist : 0.0f;
        }
        const bool found = (sine < p.angleThres && dist <= p.distThres && !isnan(ncurr.x) && !isnan(np[k].x));
        if (found) {
          const float3 d_cp = mul(P.Rprev_inv, vp[k] - tprev);
          const float3 n_cp = mul(P.Rprev_inv, np[k]);
          const float3 cr = cross(vcurr_cp, n_cp);
          const float row[7] = {n_cp.x, n_cp.y, n_cp.z, cr.x, cr.y, cr.z, dot(n_cp, vcurr_cp - d_cp)};
          accumulate_se3(acc, row, true);
        }
      }
    }
    float bt = block_reduce32_sparse(acc, warp_work, sm.red);
    if (warp == 0) prow(0, q & 1)[blockIdx.x * 32 + lane] = bt;
    DBG_MARK(8 + q * 8 + 2);
    grid_wait(gs, ++barriers * G);
    DBG_MARK(8 + q * 8 + 3);

    // -------- phase 3: RGB rows weighted with the global count -- shared memory only
    float tmpError;
    int tot_cnt;
    const float sigma = fetch_sigma(q, &tmpError, &tot_cnt);
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
    bool any_valid = false;
#pragma unroll
    for (int k = 0; k < PP; ++k) {
      const unsigned zero = __float_as_uint(at(SP_ZERO, k));
      if (zero == kNoCorr) continue;
      any_valid = true;
      const float diff = at(SP_DIFF, k), z = at(SP_D0, k);
      const unsigned sob = __float_as_uint(at(SP_SOB, k));
      float w = sigma + fabsf(diff);
      w = w > 1.19209290E-07F ? 1.0f / w : 1.0f;
      if (sigma == -1.f) w = 1.f;
      const int zx = (int)(zero & 0xffff), zy = (int)(zero >> 16);
      const float invFx = 1.0f / L.k.fx, invFy = 1.0f / L.k.fy;
      const float3 Pt = make_float3(((float)zx - L.k.cx) * z * invFx, ((float)zy - L.k.cy) * z * invFy, z);
      const float invz = (float)(1.0 / (double)Pt.z);
      const float dI_dx_val = w * p.sobelScale * (float)(short)(sob & 0xffff);
      const float dI_dy_val = w * p.sobelScale * (float)(short)(sob >> 16);
      const float v0 = dI_dx_val * L.k.fx * invz;
      const float v1 = dI_dy_val * L.k.fy * invz;
      const float v2 = -(v0 * Pt.x + v1 * Pt.y) * invz;
      const float row[7] = {v0, v1, v2, -Pt.z * v1 + Pt.y * v2, Pt.z * v0 - Pt.x * v2, -Pt.y * v0 + Pt.x * v1, -w * diff};
      accumulate_se3(acc, row, true);
    }
    // invalid correspondences contribute nothing to any of the 29 sums (reduce.cu:562-601)
    bt = block_reduce32_sparse(acc, __any_sync(0xffffffffu, any_valid), sm.red);
    if (warp == 0) prow(1, q & 1)[blockIdx.x * 32 + lane] = bt;
    DBG_MARK(8 + q * 8 + 4);
    finish_iteration(barriers, q, tmpError, tot_cnt);
  }
}

// any image size: correspondences round-trip through the DataTerm image, nothing is staged
__device__ __noinline__ void run_level_generic(int lvl, int q0, int nit, unsigned& barriers) {
  SMEM_REF();
  GridSync* gs = p.gs;
  StepScratch* sc = p.sc;
  const LevelData& L = p.L[lvl];
  const int N = L.w * L.h;
  const int tid = blockIdx.x * kPT + threadIdx.x, nthreads = gridDim.x * kPT;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, G = gridDim.x;
  auto prow = [&](int set, int parity) { return sc->partials + (size_t)((parity * 2 + set) * G) * 32; };
  LevelCtx c;
  make_level_ctx(p, L, c);
  const GNState& S = sm.S;
  for (int it = 0; it < nit; ++it) {
    const int q = q0 + it;
    DBG_MARK(8 + q * 8 + 0);
    const IcpPose& P = S.pose;
    const RgbWarp& Wp = S.warp;
    c.ia.error_map = (lvl == 0 && it + 1 == nit) ? p.err : nullptr;
    int cnt = 0, sig = 0;
    for (int px = tid; px < N; px += nthreads) {
      int y = px / L.w, x = px - y * L.w;
      DataTerm ct;
      ct.valid = false;
      ct.zero = make_short2(0, 0);
      ct.diff = 0.f;
      int sq;
      if (__ldg(L.cand + px) && rgb_residual_cand(c.ra, Wp, x, y, ct, sq)) {
        cnt += 1;
        sig += sq;
      }
      int4 raw;
      raw.x = (int)((unsigned short)ct.zero.x | ((unsigned)(unsigned short)ct.zero.y << 16));
      raw.y = 0;
      raw.z = __float_as_int(ct.diff);
      raw.w = ct.valid ? 1 : 0;
      reinterpret_cast<int4*>(L.corres)[px] = raw;
    }
    publish_counts(q, cnt, sig);
    DBG_MARK(8 + q * 8 + 1);
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
    bool work = false;
    for (int px = tid; px < N; px += nthreads) {
      int y = px / L.w, x = px - y * L.w;
      icp_pixel(c.ia, P, x, y, acc);
      work = true;
    }
    float bt = block_reduce32_sparse(acc, __any_sync(0xffffffffu, work), sm.red);
    if (warp == 0) prow(0, q & 1)[blockIdx.x * 32 + lane] = bt;
    DBG_MARK(8 + q * 8 + 2);
    grid_wait(gs, ++barriers * G);
    DBG_MARK(8 + q * 8 + 3);
    float tmpError;
    int tot_cnt;
    const float sigma = fetch_sigma(q, &tmpError, &tot_cnt);
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
    for (int px = tid; px < N; px += nthreads) {
      int4 raw = reinterpret_cast<const int4*>(L.corres)[px];
      if (raw.w & 0xff) {
        int y = px / L.w, x = px - y * L.w;
        rgb_step_from_depth(L, sigma, p.sobelScale, true, (unsigned)raw.x, __int_as_float(raw.z), x, y, acc);
      }
    }
    bt = block_reduce32_sparse(acc, true, sm.red);
    if (warp == 0) prow(1, q & 1)[blockIdx.x * 32 + lane] = bt;
    DBG_MARK(8 + q * 8 + 4);
    finish_iteration(barriers, q, tmpError, tot_cnt);
  }
}

// SO(3) pre-alignment on level 2 (RGBDOdometry.cpp:239-310)
__device__ __noinline__ void run_so3(unsigned& barriers) {
  SMEM_REF();
  GridSync* gs = p.gs;
  StepScratch* sc = p.sc;
  GNState& S = sm.S;
  const int tid = blockIdx.x * kPT + threadIdx.x, nthreads = gridDim.x * kPT;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, G = gridDim.x;
  auto prow = [&](int set, int parity) { return sc->partials + (size_t)((parity * 2 + set) * G) * 32; };
  const LevelData& L = p.L[2];
  const int N = L.w * L.h;
  for (int it = 0; it < 10; ++it) {
    if (S.so3_done) break;  // identical in every CTA
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
    bool work = false;
    for (int q = tid; q < N; q += nthreads) {
      int y = q / L.w, x = q - y * L.w;
      so3_pixel(p.so3_last, p.so3_next, (size_t)L.w, L.w, L.h, S.so3_imageBasis, S.so3_kinv, S.so3_krlr, x, y, acc);
      work = true;
    }
    float bt = block_reduce32_sparse(acc, __any_sync(0xffffffffu, work), sm.red);
    if (warp == 0) prow(0, it & 1)[blockIdx.x * 32 + lane] = bt;
    grid_arrive(gs);
    grid_wait(gs, ++barriers * G);
    sum_partials<1>(prow(0, it & 1), G, sm.red, sm.out64);
    if (threadIdx.x == 0) {
      so3_update_serial(&S, sm.out64, L.k);
      if (S.so3_done || it == 9) gn_begin_serial(&S, 1, p.L[sm.nsched ? sm.sched[0] : 0].k);
    }
    __syncthreads();
  }
  // the Gauss-Newton loop reuses the partial rows (parity 0 first): no CTA may still be folding SO(3) rows
  grid_arrive(gs);
  grid_wait(gs, ++barriers * G);
}

__global__ void __launch_bounds__(kPT, 1) gn_persistent_kernel(const PersistParams kp) {
  extern __shared__ __align__(16) unsigned char dyn_smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(dyn_smem_raw);
  {  // parameters -> shared memory (the phase functions are not inlined)
    const int* src = reinterpret_cast<const int*>(&kp);
    int* dst = reinterpret_cast<int*>(&sm.prm);
    for (int i = threadIdx.x; i < (int)(sizeof(PersistParams) / 4); i += kPT) dst[i] = src[i];
  }
  if (threadIdx.x == 0) {
    int n = 0;
    for (int i = 2; i >= 0; --i)
      for (int j = 0; j < kp.iters[i] && n < 19; ++j) sm.sched[n++] = i;
    sm.nsched = n;
  }
  __syncthreads();
  const PersistParams& p = sm.prm;
  const int nthreads = gridDim.x * kPT;
  unsigned barriers = 0;  // barriers completed so far -> target = (barriers + 1) * G
  const int nsched = sm.nsched;

  DBG_MARK(0);
  if (threadIdx.x == 0) {
    gn_init_serial(&sm.S, nullptr, p.pose_in, p.L[2].k);
    if (!p.use_so3) gn_begin_serial(&sm.S, 0, p.L[nsched ? sm.sched[0] : 0].k);
  }
  __syncthreads();
  DBG_MARK(1);
  if (p.use_so3) run_so3(barriers);

  DBG_MARK(2);
  // ---- Gauss-Newton iterations, coarse to fine (RGBDOdometry.cpp:331-461)
  int q0 = 0;
  for (int lvl = 2; lvl >= 0; --lvl) {
    int nit = p.iters[lvl];
    if (q0 + nit > nsched) nit = nsched - q0;
    if (nit <= 0) continue;
    const int need = (p.L[lvl].w * p.L[lvl].h + nthreads - 1) / nthreads;
    if (need <= 1)
      run_level_staged<1>(lvl, q0, nit, barriers);
    else if (need <= 2)
      run_level_staged<2>(lvl, q0, nit, barriers);
    else if (need <= kStagePP)
      run_level_staged<kStagePP>(lvl, q0, nit, barriers);
    else
      run_level_generic(lvl, q0, nit, barriers);
    q0 += nit;
  }
  // ---- CTA 0 publishes pose + stats
  if (blockIdx.x == 0) {
    const float* src = (const float*)&sm.S;
    float* dst = (float*)p.g;
    for (int i = threadIdx.x; i < (int)(sizeof(GNState) / 4); i += kPT) dst[i] = src[i];
  }
  DBG_MARK(3);
}

// sobel + candidate gates for all three levels in one launch
struct PrepLevel {
  const unsigned char* img;
  const float* nextDepth;
  short *dx, *dy;
  unsigned char* cand;
  int w, h;
  float minScale;
};
struct PrepParams {
  PrepLevel L[3];
  unsigned* grid_sync;  // GridSync of the persistent kernel launched next, zeroed here
};
__global__ void rgb_prepare_all_kernel(const PrepParams pp) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < (int)(sizeof(GridSync) / 4)) pp.grid_sync[q] = 0;
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    const PrepLevel& L = pp.L[l];
    const int n = L.w * L.h;
    if (q < n) {
      int y = q / L.w, x = q - y * L.w;
      rgb_prepare_pixel(L.img, L.w, L.h, L.nextDepth, L.minScale, L.dx, L.dy, L.cand, x, y);
      return;
    }
    q -= n;
  }
}

}  // namespace

#define RET_IF(e)                       \
  do {                                  \
    cudaError_t e__ = (e);              \
    if (e__ != cudaSuccess) return e__; \
  } while (0)

cudaError_t RGBDOdometry::enqueuePrepare(cudaStream_t s) {
  PrepParams pp;
  int total = 0;
  for (int i = 0; i < NUM_PYRS; ++i) {
    const int w = width >> i, h = height >> i;
    pp.L[i] = PrepLevel{nextImage[i], (next_is_last_ ? lastDepth[i] : nextDepth[i]), nextdIdx[i], nextdIdy[i], rgbCand[i], w, h,
                        (float)(pow(minimumGradientMagnitudes[i], 2.0) / pow(sobelScale, 2.0))};
    total += w * h;
  }
  pp.grid_sync = (unsigned*)grid_sync_;
  rgb_prepare_all_kernel<<<(total + 255) / 256, 256, 0, s>>>(pp);
  return cudaGetLastError();
}

cudaError_t RGBDOdometry::enqueuePersistent(float icpWeight, bool pyramid, bool fastOdom, bool so3, float* err,
                                            size_t err_pitch, cudaStream_t s) {
  float* h_in = (float*)((char*)h_pinned + 1536);
  RET_IF(cudaMemcpyAsync(d_pose_in, h_in, 12 * sizeof(float), cudaMemcpyHostToDevice, s));
  RET_IF(enqueuePrepare(s));
  PersistParams p;
  for (int i = 0; i < NUM_PYRS; ++i) {
    const int w = width >> i, h = height >> i;
    const Intr k = intr.level(i);
    LevelData& L = p.L[i];
    L.vmap_curr = vmaps_curr_[i];
    L.nmap_curr = nmaps_curr_[i];
    L.vmap_g_prev = vmaps_g_prev_[i];
    L.nmap_g_prev = nmaps_g_prev_[i];
    L.lastDepth = lastDepth[i];
    L.nextDepth = (next_is_last_ ? lastDepth[i] : nextDepth[i]);
    L.lastImage = lastImage[i];
    L.nextImage = nextImage[i];
    L.dIdx = nextdIdx[i];
    L.dIdy = nextdIdy[i];
    L.cand = rgbCand[i];
    L.corres = corresImg[i];
    L.w = w;
    L.h = h;
    L.k = LevelK{k.fx, k.fy, k.cx, k.cy};
  }
  p.so3_last = lastNextImage[2];
  p.so3_next = nextImage[2];
  p.g = gn;
  p.sc = scratch;
  p.gs = (GridSync*)grid_sync_;
  p.pose_in = d_pose_in;
  p.err = err;
  p.err_pitch = err_pitch;
  p.distThres = distThres_;
  p.angleThres = angleThres_;
  p.maxDepthDelta = maxDepthDeltaRGB;
  p.sobelScale = sobelScale;
  p.icpWeight = icpWeight;
  p.use_so3 = so3 ? 1 : 0;
  p.iters[0] = fastOdom ? 3 : 10;
  p.iters[1] = pyramid ? 5 : 0;
  p.iters[2] = pyramid ? 4 : 0;
  p.dbg = (unsigned long long*)dbg_trace_;
  int grid = num_sms();
  if (grid > kMaxBlocks) grid = kMaxBlocks;
  void* args[] = {(void*)&p};
  if (time_kernel_) RET_IF(cudaEventRecord(ev_k0_, s));
  const size_t stage_bytes = sizeof(Smem);
  {  // the opt-in to > 48 KB of dynamic shared memory is per device: once per device and process
    static bool attr_set[64] = {};
    int dev = 0;
    RET_IF(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
      RET_IF(cudaFuncSetAttribute(gn_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes));
      if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
  }
  RET_IF(cudaLaunchCooperativeKernel((const void*)gn_persistent_kernel, dim3(grid), dim3(kPT), args, stage_bytes, s));
  if (time_kernel_) {
    RET_IF(cudaEventRecord(ev_k1_, s));
    ev_pending_ = true;
  }
  struct Out {
    float trans[3];
    float rot[9];
    TrackStats st;
  };
  Out* ho = (Out*)((char*)h_pinned + 2048);
  RET_IF(cudaMemcpyAsync(ho->trans, gn->out_trans, 12 * sizeof(float), cudaMemcpyDeviceToHost, s));
  RET_IF(cudaMemcpyAsync(&ho->st, &gn->stats, sizeof(TrackStats), cudaMemcpyDeviceToHost, s));
  return cudaSuccess;
}

void RGBDOdometry::enableKernelTiming(bool on) {
  if (on && !ev_k0_) {
    cudaEventCreate(&ev_k0_);
    cudaEventCreate(&ev_k1_);
  }
  time_kernel_ = on && ev_k0_ && ev_k1_;
}

void RGBDOdometry::kernelTiming(double* sum_ms, int* launches, bool reset) {
  if (ev_pending_ && cudaEventSynchronize(ev_k1_) == cudaSuccess) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ev_k0_, ev_k1_) == cudaSuccess) {
      kernel_ms_sum_ += ms;
      kernel_launches_++;
    }
    ev_pending_ = false;
  }
  if (sum_ms) *sum_ms = kernel_ms_sum_;
  if (launches) *launches = kernel_launches_;
  if (reset) {
    kernel_ms_sum_ = 0;
    kernel_launches_ = 0;
  }
}

}  // namespace cfb
