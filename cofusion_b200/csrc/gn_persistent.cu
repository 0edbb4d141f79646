// gn_persistent.cu -- the whole tracker optimisation (SO(3) pre-alignment + 3-level ICP/RGB
// Gauss-Newton) as ONE persistent cooperative kernel: one CTA per SM, grid-wide barriers in
// software, the FP64 Gauss-Newton step executed by the last CTA to reach each barrier.
//
// Why (measured on B200, profiles/): at 640x480 a GN iteration touches <= 34 MB that sits in the
// 126 MB L2, so the work of an iteration is 3-8 us while every separate reduction kernel costs
// ~9 us of launch + prologue + last-block election latency; the 48-launch graph version spends
// two thirds of its 0.58 ms in those fixed costs.  Inside one kernel an iteration is
//     pass1 (ICP rows + RGB correspondences) -> barrier A -> pass2 (RGB rows) -> barrier B (+solve)
// and the photometric correspondences of a pixel never leave the registers of the thread that owns
// it (the reference round-trips a 16-byte DataTerm image through memory, reduce.cu:862 / :524).
//
// Arithmetic is the same per-pixel code as the stand-alone steps (tracker_device.cuh); sums are
// folded in a fixed order (per-CTA partials, then CTA order) -> bit-reproducible run to run.
#include "gn_serial.cuh"
#include "image_kernels.cuh"

namespace cfb {
namespace {
using namespace dev;

constexpr int kPT = 512;   // threads per CTA, one CTA per SM
constexpr int kMaxPP = 6;  // pixels per thread whose correspondences stay in registers

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

__global__ void __launch_bounds__(kPT, 1) gn_persistent_kernel(const PersistParams p) {
  __shared__ float red[2 * (kPT / 32) * 32];
  __shared__ float out64[64];
  __shared__ GNState S;  // every CTA keeps (and identically updates) its own copy of the GN state
  __shared__ int scnt[kPT / 32], ssig[kPT / 32];

  GridSync* gs = p.gs;
  StepScratch* sc = p.sc;
  const int tid = blockIdx.x * kPT + threadIdx.x, nthreads = gridDim.x * kPT;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned G = gridDim.x;
  unsigned barriers = 0;  // barriers completed so far -> target = (barriers + 1) * G

  int sched[19], nsched = 0;
  for (int i = 2; i >= 0; --i)
    for (int j = 0; j < p.iters[i] && nsched < 19; ++j) sched[nsched++] = i;

  DBG_MARK(0);
  if (threadIdx.x == 0) {
    gn_init_serial(&S, nullptr, p.pose_in, p.L[2].k);
    if (!p.use_so3) gn_begin_serial(&S, 0, p.L[nsched ? sched[0] : 0].k);
  }
  __syncthreads();
  // partial rows: set s in {0: ICP / SO3, 1: RGB}, double buffered by iteration parity
  auto prow = [&](int set, int parity) { return sc->partials + (size_t)((parity * 2 + set) * G) * 32; };

  DBG_MARK(1);
  // ---- SO(3) pre-alignment on level 2 (RGBDOdometry.cpp:239-310)
  if (p.use_so3) {
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
      float bt = block_reduce32_sparse(acc, __any_sync(0xffffffffu, work), red);
      if (warp == 0) prow(0, it & 1)[blockIdx.x * 32 + lane] = bt;
      grid_arrive(gs);
      grid_wait(gs, ++barriers * G);
      sum_partials<1>(prow(0, it & 1), G, red, out64);
      if (threadIdx.x == 0) {
        so3_update_serial(&S, out64, L.k);
        if (S.so3_done || it == 9) gn_begin_serial(&S, 1, p.L[nsched ? sched[0] : 0].k);
      }
      __syncthreads();
    }
  }

  DBG_MARK(2);
  // ---- Gauss-Newton iterations, coarse to fine (RGBDOdometry.cpp:331-461)
  for (int q = 0; q < nsched; ++q) {
    DBG_MARK(8 + q * 8 + 0);
    const LevelData& L = p.L[sched[q]];
    const int N = L.w * L.h;
    const bool keep = (N <= nthreads * kMaxPP);  // correspondences stay in registers
    const IcpPose& P = S.pose;
    const RgbWarp& Wp = S.warp;

    IcpArgs ia;
    const size_t pitch = (size_t)L.w * 4;
    ia.vmap_curr = {L.vmap_curr, pitch};
    ia.nmap_curr = {L.nmap_curr, pitch};
    ia.vmap_g_prev = {L.vmap_g_prev, pitch};
    ia.nmap_g_prev = {L.nmap_g_prev, pitch};
    ia.intr = Intr{L.k.fx, L.k.fy, L.k.cx, L.k.cy};
    ia.distThres = p.distThres;
    ia.angleThres = p.angleThres;
    ia.cols = L.w;
    ia.rows = L.h;
    const bool last_of_l0 = (sched[q] == 0 && (q + 1 == nsched || sched[q + 1] != 0));
    ia.error_map = last_of_l0 ? p.err : nullptr;
    ia.error_pitch = p.err_pitch;
    RgbResidualArgs ra;
    ra.minScale = 0.f;
    ra.maxDepthDelta = p.maxDepthDelta;
    ra.dIdx = L.dIdx;
    ra.dIdy = L.dIdy;
    ra.grad_pitch = (size_t)L.w * 2;
    ra.lastDepth = L.lastDepth;
    ra.nextDepth = L.nextDepth;
    ra.depth_pitch = pitch;
    ra.lastImage = L.lastImage;
    ra.nextImage = L.nextImage;
    ra.img_pitch = (size_t)L.w;
    ra.corres = L.corres;
    ra.cols = L.w;
    ra.rows = L.h;

    // -------- phase 1: photometric correspondences (count needed by every CTA before phase 3)
    int cnt = 0, sig = 0;
    unsigned kzero[kMaxPP];
    float kdiff[kMaxPP];
    unsigned kvalid = 0;
    if (keep) {
#pragma unroll
      for (int k = 0; k < kMaxPP; ++k) {
        const int px = tid + k * nthreads;
        kzero[k] = 0;
        kdiff[k] = 0.f;
        if (px < N && __ldg(L.cand + px)) {
          int y = px / L.w, x = px - y * L.w;
          DataTerm c;
          int sq;
          if (rgb_residual_cand(ra, Wp, x, y, c, sq)) {
            cnt += 1;
            sig += sq;
            kvalid |= 1u << k;
            kzero[k] = (unsigned)(unsigned short)c.zero.x | ((unsigned)(unsigned short)c.zero.y << 16);
            kdiff[k] = c.diff;
          }
        }
      }
    } else {
      for (int px = tid; px < N; px += nthreads) {
        int y = px / L.w, x = px - y * L.w;
        DataTerm c;
        c.valid = false;
        c.zero = make_short2(0, 0);
        c.diff = 0.f;
        int sq;
        if (__ldg(L.cand + px) && rgb_residual_cand(ra, Wp, x, y, c, sq)) {
          cnt += 1;
          sig += sq;
        }
        int4 raw;
        raw.x = (int)((unsigned short)c.zero.x | ((unsigned)(unsigned short)c.zero.y << 16));
        raw.y = 0;
        raw.z = __float_as_int(c.diff);
        raw.w = c.valid ? 1 : 0;
        reinterpret_cast<int4*>(L.corres)[px] = raw;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      sig += __shfl_xor_sync(0xffffffffu, sig, o);
    }
    if (lane == 0) {
      scnt[warp] = cnt;
      ssig[warp] = sig;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int c = 0, s = 0;
      for (int w = 0; w < kPT / 32; ++w) {
        c += scnt[w];
        s += ssig[w];
      }
      atomicAdd(&gs->counts[q][0], c);  // integer sums commute exactly
      atomicAdd(&gs->counts[q][1], s);
    }
    grid_arrive(gs);  // barrier A: its latency is hidden behind the ICP pass below
    DBG_MARK(8 + q * 8 + 1);

    // -------- phase 2: ICP rows (independent of the count)
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
    bool work = false;
    for (int px = tid; px < N; px += nthreads) {
      int y = px / L.w, x = px - y * L.w;
      icp_pixel(ia, P, x, y, acc);
      work = true;
    }
    float bt = block_reduce32_sparse(acc, __any_sync(0xffffffffu, work), red);
    if (warp == 0) prow(0, q & 1)[blockIdx.x * 32 + lane] = bt;
    DBG_MARK(8 + q * 8 + 2);
    grid_wait(gs, ++barriers * G);
    DBG_MARK(8 + q * 8 + 3);

    // -------- phase 3: RGB rows weighted with the global count
    if (threadIdx.x == 0) {
      scnt[0] = __ldcg(&gs->counts[q][0]);
      ssig[0] = __ldcg(&gs->counts[q][1]);
    }
    __syncthreads();
    const int tot_cnt = scnt[0], tot_sig = ssig[0];
    float tmpError;
    const float sigma = rgb_sigma_from_counts(tot_cnt, tot_sig, &tmpError);
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
    if (keep) {
#pragma unroll
      for (int k = 0; k < kMaxPP; ++k) {
        const int px = tid + k * nthreads;
        if (px < N && ((kvalid >> k) & 1u)) {
          int y = px / L.w, x = px - y * L.w;
          rgb_step_from_depth(L, sigma, p.sobelScale, true, kzero[k], kdiff[k], x, y, acc);
        }
      }
    } else {
      for (int px = tid; px < N; px += nthreads) {
        int4 raw = reinterpret_cast<const int4*>(L.corres)[px];
        if (raw.w & 0xff) {
          int y = px / L.w, x = px - y * L.w;
          rgb_step_from_depth(L, sigma, p.sobelScale, true, (unsigned)raw.x, __int_as_float(raw.z), x, y, acc);
        }
      }
    }
    // invalid correspondences contribute nothing to any of the 29 sums (reduce.cu:562-601)
    bt = block_reduce32_sparse(acc, __any_sync(0xffffffffu, kvalid != 0 || !keep), red);
    if (warp == 0) prow(1, q & 1)[blockIdx.x * 32 + lane] = bt;
    DBG_MARK(8 + q * 8 + 4);
    grid_arrive(gs);  // barrier B
    grid_wait(gs, ++barriers * G);
    DBG_MARK(8 + q * 8 + 5);
    sum_partials<2>(prow(0, q & 1), G, red, out64);
    DBG_MARK(8 + q * 8 + 6);
    const int is_last = (q + 1 == nsched);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int i = 0; i < 32; ++i) S.icp_result[i] = out64[i];
      gn_solve_serial(&S, nullptr, out64 + 32, p.icpWeight, p.L[is_last ? sched[q] : sched[q + 1]].k, is_last, tmpError,
                      tot_cnt);
    }
    __syncthreads();
    DBG_MARK(8 + q * 8 + 7);
  }
  // ---- CTA 0 publishes pose + stats
  if (blockIdx.x == 0) {
    const float* src = (const float*)&S;
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

cudaError_t RGBDOdometry::enqueuePersistent(float icpWeight, bool pyramid, bool fastOdom, bool so3, float* err,
                                            size_t err_pitch, cudaStream_t s) {
  float* h_in = (float*)((char*)h_pinned + 1536);
  RET_IF(cudaMemcpyAsync(d_pose_in, h_in, 12 * sizeof(float), cudaMemcpyHostToDevice, s));
  PrepParams pp;
  PersistParams p;
  int total = 0;
  for (int i = 0; i < NUM_PYRS; ++i) {
    const int w = width >> i, h = height >> i;
    const Intr k = intr.level(i);
    pp.L[i] = PrepLevel{nextImage[i], (next_is_last_ ? lastDepth[i] : nextDepth[i]), nextdIdx[i], nextdIdy[i], rgbCand[i], w, h,
                        (float)(pow(minimumGradientMagnitudes[i], 2.0) / pow(sobelScale, 2.0))};
    total += w * h;
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
  pp.grid_sync = (unsigned*)grid_sync_;
  rgb_prepare_all_kernel<<<(total + 255) / 256, 256, 0, s>>>(pp);
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
  RET_IF(cudaLaunchCooperativeKernel((const void*)gn_persistent_kernel, dim3(grid), dim3(kPT), args, 0, s));
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
