// gn_batched.cu -- the tracker optimisation of SEVERAL models of one frame in ONE persistent
// cooperative kernel (reference: `for (auto model : models) model->performTracking(...)`,
// Core/CoFusion.cpp:213-218, one RGBDOdometry::getIncrementalTransformation per model).
//
// Why: a single-model launch (gn_persistent.cu) spends ~40 % of its 0.37 ms in costs that do not
// depend on the amount of pixel work -- 38 grid barriers, 19 folds of per-CTA partial rows, 19 serial
// FP64 solves, the SO(3) pre-alignment.  An object model covers a few percent of the image, so with
// five models four of the five launches are almost pure fixed cost.  Here every Gauss-Newton
// iteration is executed for all models between the SAME two grid barriers:
//     residual(m0..mk) -> arrive A -> ICP(m0..mk) -> wait A -> RGB rows(m0..mk) -> barrier B
//     -> fold the rows of all models -> warp m solves model m (the FP64 solves run side by side)
// and the frame-side inputs staged in shared memory (current vertex/normal, Sobel, grey) are shared by
// all models.  Per-pixel arithmetic is the single-model code line for line, sums are folded in the same
// fixed order: every model gets bit-identical results to its own single-model launch.
#include "gn_serial.cuh"
#include "image_kernels.cuh"

namespace cfb {
namespace {
using namespace dev;

constexpr int kPT = 512;     // threads per CTA, one CTA per SM
constexpr int kStagePP = 5;  // pixels per thread that can be staged (640x480 on 148 x 512 threads)
constexpr int kMaxB = RGBDOdometry::kMaxBatch;
constexpr unsigned kNoCorr = 0xffffffffu;

struct BLevel {  // per model, per level
  const float *vmap_g_prev, *nmap_g_prev, *lastDepth, *nextDepth;
  const unsigned char *lastImage, *cand;
};
struct BModel {
  BLevel L[3];
  const unsigned char *so3_last, *so3_next;
  GNState* g;
  const float* pose_in;
  float* err;
};
struct FLevel {  // frame side, shared by all models
  const float *vmap_curr, *nmap_curr;
  const unsigned char* nextImage;
  const short *dIdx, *dIdy;
  int w, h;
  LevelK k;
};
struct BatchSync {  // zeroed before every launch
  unsigned arrive;
  unsigned pad[31];
  int counts[32][kMaxB][2];
};
struct BatchParams {
  BModel M[kMaxB];
  FLevel F[3];
  int nmodels;
  float* partials;  // [parity 2][model][set 2][G][32]
  BatchSync* gs;
  size_t err_pitch;
  float distThres, angleThres, maxDepthDelta, sobelScale, icpWeight;
  int use_so3;
  int iters[3];
};

enum SharedPlane { FP_VX, FP_VY, FP_VZ, FP_NX, FP_NY, FP_NZ, FP_SOB, FP_FLAGS, FP_COUNT };
struct BSmem {
  float stage[FP_COUNT * kStagePP * kPT];   // frame-side planes, [plane][k][thread]
  float corr[kMaxB][2][kStagePP * kPT];     // per model: packed correspondence, depth of the matched point
  float red[4 * (kPT / 32) * 32];
  float wrow[kMaxB][kPT / 32][32];  // per model, per warp: the warp's 32 partial sums (one barrier for all models)
  int cntw[kMaxB][kPT / 32], sigw[kMaxB][kPT / 32];
  float out[kMaxB][64];
  GNState S[kMaxB];  // every CTA keeps (and identically updates) its own copy of every model's state
  int scnt[kPT / 32], ssig[kPT / 32];
  BatchParams prm;
  int sched[20];
  int nsched;
};
#define BSMEM_REF()                                               \
  extern __shared__ __align__(16) unsigned char dyn_smem_raw[]; \
  BSmem& sm = *reinterpret_cast<BSmem*>(dyn_smem_raw);            \
  const BatchParams& p = sm.prm

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void grid_arrive(BatchSync* gs) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(&gs->arrive, 1u);
  }
}
__device__ __forceinline__ void grid_wait(BatchSync* gs, unsigned target) {
  if (threadIdx.x == 0) {
    while (ld_acquire(&gs->arrive) < target) __nanosleep(20);
  }
  __syncthreads();
}

// block reduce that lets warps without any contribution skip the 31-shuffle transpose
__device__ __forceinline__ float block_reduce32_sparse(float (&v)[32], bool warp_has_work, float* smem) {
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float t = 0.f;
  if (warp_has_work) t = warp_transpose_reduce32(v);
  __syncthreads();  // the previous use of `smem` (this CTA's last reduction) is complete
  smem[warp * 32 + lane] = t;
  __syncthreads();
  float s = 0.f;
  if (warp == 0)
    for (unsigned w = 0; w < nw; ++w) s += smem[w * 32 + lane];
  return s;
}

// partial rows of (parity, model, set): G consecutive rows of 32 floats
__device__ __forceinline__ float* prow(const BatchParams& p, int parity, int m, int set) {
  return p.partials + (size_t)(((parity * kMaxB + m) * 2 + set) * gridDim.x) * 32;
}

// fixed-order fold of the per-CTA rows of `nsets` sets (set s -> rows base + s * G * 32) into
// out[s * 32 + lane]; same order as the single-model kernel: warp w adds rows w, w+16, ..., then the 16
// warp sums are added in warp order.  Four sets per round share the staging buffer.
__device__ __forceinline__ void fold_sets(const float* base, int nsets, float* smem, float* out) {
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5, G = gridDim.x;
  for (int s0 = 0; s0 < nsets; s0 += 4) {
    const int ns = min(4, nsets - s0);
    for (int s = 0; s < ns; ++s) {
      const float* rows = base + (size_t)(s0 + s) * G * 32;
      float v[12];
#pragma unroll
      for (int k = 0; k < 12; ++k) {  // all loads in flight before the first add
        const unsigned b = warp + k * nw;
        v[k] = (b < G) ? __ldcg(&rows[b * 32 + lane]) : 0.f;
      }
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 12; ++k) acc += v[k];
      for (unsigned b = warp + 12 * nw; b < G; b += nw) acc += __ldcg(&rows[b * 32 + lane]);
      smem[(s * (kPT / 32) + warp) * 32 + lane] = acc;
    }
    __syncthreads();
    if ((int)warp < ns) {
      float tot = 0.f;
      for (unsigned w = 0; w < nw; ++w) tot += smem[(warp * (kPT / 32) + w) * 32 + lane];
      out[(s0 + warp) * 32 + lane] = tot;
    }
    __syncthreads();
  }
}

// correspondence of one pixel packed into 32 bits: u0 (11) | v0 (11) | diff + 256 (10).  diff is the
// difference of two 8-bit intensities, i.e. an integer in [-255, 255]: the packing is lossless.
__device__ __forceinline__ unsigned pack_corr(int u0, int v0, float diff) {
  return (unsigned)u0 | ((unsigned)v0 << 11) | ((unsigned)((int)diff + 256) << 22);
}

template <int PP>
__device__ __noinline__ void run_level_batched(int lvl, int q0, int nit, unsigned& barriers) {
  BSMEM_REF();
  BatchSync* gs = p.gs;
  const FLevel& F = p.F[lvl];
  const int N = F.w * F.h, W = F.w, H = F.h, NM = p.nmodels;
  const int tid = blockIdx.x * kPT + threadIdx.x, nthreads = gridDim.x * kPT;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, G = gridDim.x;
  auto at = [&](int plane, int k) -> float& { return sm.stage[(plane * PP + k) * kPT + threadIdx.x]; };
  auto cz = [&](int m, int k) -> float& { return sm.corr[m][0][k * kPT + threadIdx.x]; };
  auto cd = [&](int m, int k) -> float& { return sm.corr[m][1][k * kPT + threadIdx.x]; };
  const size_t pitch = (size_t)W * 4;

  // ---- stage the frame-side, iteration-invariant inputs of this thread's pixels (shared by all models)
  bool any_px = false;
#pragma unroll
  for (int k = 0; k < PP; ++k) {
    const int px = tid + k * nthreads;
    unsigned flags = 0;
    if (px < N) {
      any_px = true;
      const int y = px / W, x = px - y * W;
      const PlanarMap vm{F.vmap_curr, pitch}, nm{F.nmap_curr, pitch};
      at(FP_VX, k) = ldplane(vm, y, x);
      at(FP_VY, k) = ldplane(vm, y + H, x);
      at(FP_VZ, k) = ldplane(vm, y + 2 * H, x);
      at(FP_NX, k) = ldplane(nm, y, x);
      at(FP_NY, k) = ldplane(nm, y + H, x);
      at(FP_NZ, k) = ldplane(nm, y + 2 * H, x);
      const unsigned sob = (unsigned)(unsigned short)__ldg(F.dIdx + px) | ((unsigned)(unsigned short)__ldg(F.dIdy + px) << 16);
      at(FP_SOB, k) = __uint_as_float(sob);
      flags = 0x10000u | ((unsigned)__ldg(F.nextImage + px) << 8);
    }
    at(FP_FLAGS, k) = __uint_as_float(flags);
  }
  const bool warp_work = __any_sync(0xffffffffu, any_px);

  for (int it = 0; it < nit; ++it) {
    const int q = q0 + it;
    const bool last_of_l0 = (lvl == 0 && it + 1 == nit);

    // -------- phase 1: photometric correspondences of every model
    for (int m = 0; m < NM; ++m) {
      const BLevel& L = p.M[m].L[lvl];
      const RgbWarp& Wp = sm.S[m].warp;
      int cnt = 0, sig = 0;
      int u0[PP], v0[PP];
      float td1[PP];
      bool ok[PP];
#pragma unroll
      for (int k = 0; k < PP; ++k) {
        const int px = tid + k * nthreads;
        ok[k] = px < N && __ldg(L.cand + px) != 0;
        u0[k] = v0[k] = 0;
        td1[k] = 0.f;
        if (ok[k]) {
          const int y = px / W, x = px - y * W;
          const float d1 = __ldg(L.nextDepth + px);
          const float* kk = Wp.krkinv.m;
          td1[k] = d1 * (kk[6] * x + kk[7] * y + kk[8]) + Wp.kt[2];
          u0[k] = __float2int_rn((d1 * (kk[0] * x + kk[1] * y + kk[2]) + Wp.kt[0]) / td1[k]);
          v0[k] = __float2int_rn((d1 * (kk[3] * x + kk[4] * y + kk[5]) + Wp.kt[1]) / td1[k]);
          ok[k] = (u0[k] >= 0 && v0[k] >= 0 && u0[k] < W && v0[k] < H);
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
          const unsigned flags = __float_as_uint(at(FP_FLAGS, k));
          const float diff = (float)((flags >> 8) & 0xffu) - (float)li[k];
          cnt += 1;
          sig += (int)(diff * diff);  // float -> int truncation, reduce.cu:851
          zero = pack_corr(u0[k], v0[k], diff);
          cd(m, k) = d0[k];
        }
        cz(m, k) = __uint_as_float(zero);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        sig += __shfl_xor_sync(0xffffffffu, sig, o);
      }
      if (lane == 0) {
        sm.cntw[m][warp] = cnt;
        sm.sigw[m][warp] = sig;
      }
    }
    __syncthreads();
    if ((int)threadIdx.x < NM) {  // one thread per model folds the 16 warp counts (integer sums commute exactly)
      int c = 0, sg = 0;
      for (int w = 0; w < kPT / 32; ++w) {
        c += sm.cntw[threadIdx.x][w];
        sg += sm.sigw[threadIdx.x][w];
      }
      if (c) atomicAdd(&gs->counts[q][threadIdx.x][0], c);
      if (sg) atomicAdd(&gs->counts[q][threadIdx.x][1], sg);
    }
    grid_arrive(gs);  // barrier A: its latency is hidden behind the ICP passes

    // -------- phase 2: ICP rows of every model (independent of the counts)
    for (int m = 0; m < NM; ++m) {
      const BLevel& L = p.M[m].L[lvl];
      const IcpPose& P = sm.S[m].pose;
      float* const error_map = last_of_l0 ? p.M[m].err : nullptr;
      const PlanarMap vprev{L.vmap_g_prev, pitch}, nprev{L.nmap_g_prev, pitch};
      float acc[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = 0.f;
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
          const float3 vcurr = make_float3(at(FP_VX, k), at(FP_VY, k), at(FP_VZ, k));
          const float3 vcurr_g = mul(P.Rcurr, vcurr) + tcurr;
          const float3 vcurr_cp = mul(P.Rprev_inv, vcurr_g - tprev);
          ux[k] = __float2int_rn(vcurr_cp.x * F.k.fx / vcurr_cp.z + F.k.cx);
          uy[k] = __float2int_rn(vcurr_cp.y * F.k.fy / vcurr_cp.z + F.k.cy);
          ok[k] = !(ux[k] < 0 || uy[k] < 0 || ux[k] >= W || uy[k] >= H || vcurr_cp.z < 0);
          if (!ok[k] && error_map) {
            const int y = px / W, x = px - y * W;
            row_ptr(error_map, p.err_pitch, y)[x] = 0.0f;
          }
        }
      }
      float3 vp[PP], np[PP];
      if (m == 0) {  // the camera model sees (nearly) every pixel: all six gathers at once
#pragma unroll
        for (int k = 0; k < PP; ++k) {
          vp[k] = np[k] = make_float3(0.f, 0.f, 0.f);
          if (ok[k]) {
            vp[k] = make_float3(ldplane(vprev, uy[k], ux[k]), ldplane(vprev, uy[k] + H, ux[k]),
                                ldplane(vprev, uy[k] + 2 * H, ux[k]));
            np[k] = make_float3(ldplane(nprev, uy[k], ux[k]), ldplane(nprev, uy[k] + H, ux[k]),
                                ldplane(nprev, uy[k] + 2 * H, ux[k]));
          }
        }
      } else {
        // An object model predicts a few percent of the image; everywhere else its vertex map is NaN.
        // Gather the x plane first: a NaN there makes dist NaN, i.e. no correspondence and error 0
        // whatever the other five planes hold -- they are only fetched for the pixels that hit the object.
        float vx[PP];
#pragma unroll
        for (int k = 0; k < PP; ++k) vx[k] = ok[k] ? ldplane(vprev, uy[k], ux[k]) : 0.f;
#pragma unroll
        for (int k = 0; k < PP; ++k) {
          vp[k] = np[k] = make_float3(0.f, 0.f, 0.f);
          if (ok[k] && isnan(vx[k])) {
            ok[k] = false;
            if (error_map) {
              const int px = tid + k * nthreads, y = px / W, x = px - y * W;
              row_ptr(error_map, p.err_pitch, y)[x] = 0.0f;
            }
          }
          if (ok[k]) {
            vp[k] = make_float3(vx[k], ldplane(vprev, uy[k] + H, ux[k]), ldplane(vprev, uy[k] + 2 * H, ux[k]));
            np[k] = make_float3(ldplane(nprev, uy[k], ux[k]), ldplane(nprev, uy[k] + H, ux[k]),
                                ldplane(nprev, uy[k] + 2 * H, ux[k]));
          }
        }
      }
      bool any_found = false;
#pragma unroll
      for (int k = 0; k < PP; ++k) {
        if (!ok[k]) continue;
        const float3 vcurr = make_float3(at(FP_VX, k), at(FP_VY, k), at(FP_VZ, k));
        const float3 ncurr = make_float3(at(FP_NX, k), at(FP_NY, k), at(FP_NZ, k));
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
          any_found = true;
          const float3 d_cp = mul(P.Rprev_inv, vp[k] - tprev);
          const float3 n_cp = mul(P.Rprev_inv, np[k]);
          const float3 cr = cross(vcurr_cp, n_cp);
          const float row[7] = {n_cp.x, n_cp.y, n_cp.z, cr.x, cr.y, cr.z, dot(n_cp, vcurr_cp - d_cp)};
          accumulate_se3(acc, row, true);
        }
      }
      // a warp without any correspondence contributes exact zeros: skip its transpose reduction
      sm.wrow[m][warp][lane] = (warp_work && __any_sync(0xffffffffu, any_found)) ? warp_transpose_reduce32(acc) : 0.f;
    }
    __syncthreads();
    if ((int)warp < NM) {  // warp m adds the 16 warp rows of model m in warp order (as the single-model kernel)
      float sacc = 0.f;
      for (int w = 0; w < kPT / 32; ++w) sacc += sm.wrow[warp][w][lane];
      prow(p, q & 1, warp, 0)[blockIdx.x * 32 + lane] = sacc;
    }
    grid_wait(gs, ++barriers * G);

    // -------- phase 3: RGB rows of every model, weighted with its global count -- shared memory only
    if (threadIdx.x < NM) {
      sm.scnt[threadIdx.x] = __ldcg(&gs->counts[q][threadIdx.x][0]);
      sm.ssig[threadIdx.x] = __ldcg(&gs->counts[q][threadIdx.x][1]);
    }
    __syncthreads();
    float tmpErr[kMaxB];
    int totCnt[kMaxB];
    for (int m = 0; m < NM; ++m) {
      totCnt[m] = sm.scnt[m];
      const float sigma = rgb_sigma_from_counts(sm.scnt[m], sm.ssig[m], &tmpErr[m]);
      float acc[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = 0.f;
      bool any_valid = false;
#pragma unroll
      for (int k = 0; k < PP; ++k) {
        const unsigned zero = __float_as_uint(cz(m, k));
        if (zero == kNoCorr) continue;
        any_valid = true;
        const float diff = (float)((int)(zero >> 22) - 256), z = cd(m, k);
        const unsigned sob = __float_as_uint(at(FP_SOB, k));
        float w = sigma + fabsf(diff);
        w = w > 1.19209290E-07F ? 1.0f / w : 1.0f;
        if (sigma == -1.f) w = 1.f;
        const int zx = (int)(zero & 0x7ffu), zy = (int)((zero >> 11) & 0x7ffu);
        const float invFx = 1.0f / F.k.fx, invFy = 1.0f / F.k.fy;
        const float3 Pt = make_float3(((float)zx - F.k.cx) * z * invFx, ((float)zy - F.k.cy) * z * invFy, z);
        const float invz = (float)(1.0 / (double)Pt.z);
        const float dI_dx_val = w * p.sobelScale * (float)(short)(sob & 0xffff);
        const float dI_dy_val = w * p.sobelScale * (float)(short)(sob >> 16);
        const float v0 = dI_dx_val * F.k.fx * invz;
        const float v1 = dI_dy_val * F.k.fy * invz;
        const float v2 = -(v0 * Pt.x + v1 * Pt.y) * invz;
        const float row[7] = {v0, v1, v2, -Pt.z * v1 + Pt.y * v2, Pt.z * v0 - Pt.x * v2, -Pt.y * v0 + Pt.x * v1, -w * diff};
        accumulate_se3(acc, row, true);
      }
      sm.wrow[m][warp][lane] = __any_sync(0xffffffffu, any_valid) ? warp_transpose_reduce32(acc) : 0.f;
    }
    __syncthreads();
    if ((int)warp < NM) {
      float sacc = 0.f;
      for (int w = 0; w < kPT / 32; ++w) sacc += sm.wrow[warp][w][lane];
      prow(p, q & 1, warp, 1)[blockIdx.x * 32 + lane] = sacc;
    }
    grid_arrive(gs);  // barrier B
    grid_wait(gs, ++barriers * G);
    // rows of model m: sets 2m (ICP), 2m+1 (RGB) are consecutive -> out[m][0..63]
    fold_sets(prow(p, q & 1, 0, 0), 2 * NM, sm.red, &sm.out[0][0]);
    const int is_last = (q + 1 == sm.nsched);
    if ((int)warp < NM && lane == 0) {  // the FP64 solves of the models run side by side, one warp each
      const int m = (int)warp;
#pragma unroll
      for (int i = 0; i < 32; ++i) sm.S[m].icp_result[i] = sm.out[m][i];
      gn_solve_serial(&sm.S[m], nullptr, &sm.out[m][32], p.icpWeight, p.F[is_last ? sm.sched[q] : sm.sched[q + 1]].k, is_last,
                      tmpErr[m], totCnt[m]);
    }
    __syncthreads();
  }
}

// SO(3) pre-alignment of every model on level 2 (RGBDOdometry.cpp:239-310)
__device__ __noinline__ void run_so3_batched(unsigned& barriers) {
  BSMEM_REF();
  BatchSync* gs = p.gs;
  const FLevel& F = p.F[2];
  const int N = F.w * F.h, NM = p.nmodels;
  const int tid = blockIdx.x * kPT + threadIdx.x, nthreads = gridDim.x * kPT;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, G = gridDim.x;
  for (int it = 0; it < 10; ++it) {
    bool all_done = true;  // identical in every CTA
    for (int m = 0; m < NM; ++m) all_done = all_done && sm.S[m].so3_done;
    if (all_done) break;
    for (int m = 0; m < NM; ++m) {
      if (sm.S[m].so3_done) continue;
      const GNState& S = sm.S[m];
      float acc[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = 0.f;
      bool work = false;
      for (int q = tid; q < N; q += nthreads) {
        int y = q / F.w, x = q - y * F.w;
        so3_pixel(p.M[m].so3_last, p.M[m].so3_next, (size_t)F.w, F.w, F.h, S.so3_imageBasis, S.so3_kinv, S.so3_krlr, x, y,
                  acc);
        work = true;
      }
      const float bt = block_reduce32_sparse(acc, __any_sync(0xffffffffu, work), sm.red);
      if (warp == 0) prow(p, it & 1, m, 0)[blockIdx.x * 32 + lane] = bt;
    }
    grid_arrive(gs);
    grid_wait(gs, ++barriers * G);
    // only set 0 of every model is used here; folding both sets keeps one code path (set 1 holds stale rows)
    fold_sets(prow(p, it & 1, 0, 0), 2 * NM, sm.red, &sm.out[0][0]);
    if ((int)warp < NM && lane == 0) {
      GNState& S = sm.S[warp];
      if (!S.so3_done) {
        so3_update_serial(&S, sm.out[warp], F.k);
        if (S.so3_done || it == 9) gn_begin_serial(&S, 1, p.F[sm.nsched ? sm.sched[0] : 0].k);
      }
    }
    __syncthreads();
  }
  // the Gauss-Newton loop reuses the partial rows (parity 0 first): no CTA may still be folding SO(3) rows
  grid_arrive(gs);
  grid_wait(gs, ++barriers * G);
}

__global__ void __launch_bounds__(kPT, 1) gn_batched_kernel(const BatchParams kp) {
  extern __shared__ __align__(16) unsigned char dyn_smem_raw[];
  BSmem& sm = *reinterpret_cast<BSmem*>(dyn_smem_raw);
  {
    const int* src = reinterpret_cast<const int*>(&kp);
    int* dst = reinterpret_cast<int*>(&sm.prm);
    for (int i = threadIdx.x; i < (int)(sizeof(BatchParams) / 4); i += kPT) dst[i] = src[i];
  }
  if (threadIdx.x == 0) {
    int n = 0;
    for (int i = 2; i >= 0; --i)
      for (int j = 0; j < kp.iters[i] && n < 19; ++j) sm.sched[n++] = i;
    sm.nsched = n;
  }
  __syncthreads();
  const BatchParams& p = sm.prm;
  const int nthreads = gridDim.x * kPT, NM = p.nmodels;
  unsigned barriers = 0;
  const int nsched = sm.nsched;
  if ((threadIdx.x & 31) == 0 && (int)(threadIdx.x >> 5) < NM) {
    const int m = threadIdx.x >> 5;
    gn_init_serial(&sm.S[m], nullptr, p.M[m].pose_in, p.F[2].k);
    if (!p.use_so3) gn_begin_serial(&sm.S[m], 0, p.F[nsched ? sm.sched[0] : 0].k);
  }
  __syncthreads();
  if (p.use_so3) run_so3_batched(barriers);
  int q0 = 0;
  for (int lvl = 2; lvl >= 0; --lvl) {
    int nit = p.iters[lvl];
    if (q0 + nit > nsched) nit = nsched - q0;
    if (nit <= 0) continue;
    const int need = (p.F[lvl].w * p.F[lvl].h + nthreads - 1) / nthreads;
    if (need <= 1)
      run_level_batched<1>(lvl, q0, nit, barriers);
    else if (need <= 2)
      run_level_batched<2>(lvl, q0, nit, barriers);
    else
      run_level_batched<kStagePP>(lvl, q0, nit, barriers);  // the host checked need <= kStagePP
    q0 += nit;
  }
  // ---- CTA 0 publishes pose + stats of every model
  if (blockIdx.x == 0)
    for (int m = 0; m < NM; ++m) {
      const float* src = (const float*)&sm.S[m];
      float* dst = (float*)p.M[m].g;
      for (int i = threadIdx.x; i < (int)(sizeof(GNState) / 4); i += kPT) dst[i] = src[i];
    }
}

}  // namespace

#define RET_IF(e)                       \
  do {                                  \
    cudaError_t e__ = (e);              \
    if (e__ != cudaSuccess) return e__; \
  } while (0)

size_t RGBDOdometry::batchScratchBytes() {
  return sizeof(BatchSync) + (size_t)2 * kMaxB * 2 * kMaxBlocks * 32 * sizeof(float);
}

bool RGBDOdometry::canBatch(int n) const {
  const int threads = num_sms() * kPT;
  return n >= 2 && n <= kMaxB && mode_ == 0 && (width * height + threads - 1) / threads <= kStagePP &&
         width < 2048 && height < 2048;
}

// All odometry objects belong to one frame (same geometry, same frame-side inputs, initAll() done on
// stream s).  trans / rot: n x 3 / n x 9 host arrays, in/out.  scratch: batchScratchBytes() of device
// memory owned by the caller (the context).
cudaError_t RGBDOdometry::trackBatched(RGBDOdometry* const* od, int n, float (*trans)[3], float (*rot)[9], float icpWeight,
                                       bool pyramid, bool fastOdom, bool so3, float* const* err, size_t err_pitch,
                                       void* scratch, cudaStream_t s) {
  if (n < 2 || n > kMaxB || !scratch) return cudaErrorInvalidValue;
  struct Out {
    float trans[3];
    float rot[9];
    TrackStats st;
  };
  BatchParams p;
  memset(&p, 0, sizeof(p));
  RGBDOdometry& f = *od[0];
  for (int m = 0; m < n; ++m) {
    RGBDOdometry& o = *od[m];
    float* h_in = (float*)((char*)o.h_pinned + 1536);
    memcpy(h_in, trans[m], 3 * sizeof(float));
    memcpy(h_in + 3, rot[m], 9 * sizeof(float));
    RET_IF(cudaMemcpyAsync(o.d_pose_in, h_in, 12 * sizeof(float), cudaMemcpyHostToDevice, s));
    RET_IF(o.enqueuePrepare(s));  // Sobel images + candidate gates of this model's "next" pyramid
    BModel& M = p.M[m];
    for (int i = 0; i < NUM_PYRS; ++i) {
      BLevel& L = M.L[i];
      L.vmap_g_prev = o.vmaps_g_prev_[i];
      L.nmap_g_prev = o.nmaps_g_prev_[i];
      L.lastDepth = o.lastDepth[i];
      L.nextDepth = o.next_is_last_ ? o.lastDepth[i] : o.nextDepth[i];
      L.lastImage = o.lastImage[i];
      L.cand = o.rgbCand[i];
    }
    M.so3_last = o.lastNextImage[2];
    M.so3_next = o.nextImage[2];
    M.g = o.gn;
    M.pose_in = o.d_pose_in;
    M.err = err ? err[m] : nullptr;
  }
  for (int i = 0; i < NUM_PYRS; ++i) {
    FLevel& F = p.F[i];
    const Intr k = f.intr.level(i);
    F.vmap_curr = f.vmaps_curr_[i];
    F.nmap_curr = f.nmaps_curr_[i];
    F.nextImage = f.nextImage[i];
    F.dIdx = f.nextdIdx[i];
    F.dIdy = f.nextdIdy[i];
    F.w = f.width >> i;
    F.h = f.height >> i;
    F.k = LevelK{k.fx, k.fy, k.cx, k.cy};
  }
  p.nmodels = n;
  p.gs = (BatchSync*)scratch;
  p.partials = (float*)((char*)scratch + sizeof(BatchSync));
  p.err_pitch = err_pitch;
  p.distThres = f.distThres_;
  p.angleThres = f.angleThres_;
  p.maxDepthDelta = f.maxDepthDeltaRGB;
  p.sobelScale = f.sobelScale;
  p.icpWeight = icpWeight;
  p.use_so3 = so3 ? 1 : 0;
  p.iters[0] = fastOdom ? 3 : 10;
  p.iters[1] = pyramid ? 5 : 0;
  p.iters[2] = pyramid ? 4 : 0;
  RET_IF(cudaMemsetAsync(p.gs, 0, sizeof(BatchSync), s));
  int grid = num_sms();
  if (grid > kMaxBlocks) grid = kMaxBlocks;
  {  // the opt-in to > 48 KB of dynamic shared memory is per device: once per device and process
    static bool attr_set[64] = {};
    int dev = 0;
    RET_IF(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
      RET_IF(cudaFuncSetAttribute(gn_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BSmem)));
      if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
  }
  void* args[] = {(void*)&p};
  if (f.time_kernel_) RET_IF(cudaEventRecord(f.ev_k0_, s));
  RET_IF(cudaLaunchCooperativeKernel((const void*)gn_batched_kernel, dim3(grid), dim3(kPT), args, sizeof(BSmem), s));
  if (f.time_kernel_) {
    RET_IF(cudaEventRecord(f.ev_k1_, s));
    f.ev_pending_ = true;
  }
  for (int m = 0; m < n; ++m) {
    RGBDOdometry& o = *od[m];
    Out* ho = (Out*)((char*)o.h_pinned + 2048);
    RET_IF(cudaMemcpyAsync(ho->trans, o.gn->out_trans, 12 * sizeof(float), cudaMemcpyDeviceToHost, s));
    RET_IF(cudaMemcpyAsync(&ho->st, &o.gn->stats, sizeof(TrackStats), cudaMemcpyDeviceToHost, s));
  }
  RET_IF(cudaStreamSynchronize(s));
  if (f.time_kernel_) f.kernelTiming(nullptr, nullptr, false);
  for (int m = 0; m < n; ++m) {
    RGBDOdometry& o = *od[m];
    Out* ho = (Out*)((char*)o.h_pinned + 2048);
    memcpy(trans[m], ho->trans, sizeof(float) * 3);
    memcpy(rot[m], ho->rot, sizeof(float) * 9);
    o.stats_ = ho->st;
    if (so3) {
      for (int i = 0; i < NUM_PYRS; i++) {
        unsigned char* t = o.lastNextImage[i];
        o.lastNextImage[i] = o.nextImage[i];
        o.nextImage[i] = t;
      }
      o.parity_ ^= 1;
    }
  }
  return cudaSuccess;
}

}  // namespace cfb
