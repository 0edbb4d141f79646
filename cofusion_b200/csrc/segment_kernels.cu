// segment_kernels.cu -- motion segmentation on the device: SLIC super-pixels, super-pixel
// reductions, unaries, fully connected CRF mean field, connected components + post-processing,
// label upsampling.  Reference: Core/Segmentation/{Segmentation.cpp:124-706, Slic.h, Slic.cpp,
// ConnectedLabels.hpp} (CPU, single thread) + gSLICr + densecrf (un-vendored, unpinned).
//
// Compiled with -fmad=false; every float sum whose order matters is accumulated in the reference's
// order (pixel / super-pixel index ascending), so label maps, component ids and ModelData are
// bit-identical to the CPU restatement.  The CRF message passing evaluates the Gaussian kernels
// exactly over the 1200 (4800 @1280x960) super-pixels -- K is built once per frame (two N x N f32
// matrices, 11.5 MB, L2 resident) and applied 10 x 2 times; densecrf's permutohedral lattice
// approximates the same product (frozen choice shared with the oracle).
#include "segment_kernels.cuh"

#include <stdlib.h>

#include "detmath.cuh"

namespace cfb {
namespace {

constexpr int kSp = 16;  // super-pixel size (Segmentation.cpp:55)

// ------------------------------------------------------------------------------------- SLIC
// One CTA per 16x16 grid cell: every pixel of the cell searches the same 3x3 centres, so they are
// staged in shared memory once.  The centre update of the previous iteration is folded into the
// prologue (sum / count from the integer accumulators of the previous launch -- exact, order free)
// and this launch accumulates the sums the next one needs: 6 launches instead of 1 + 6 + 5 + 1.
__global__ void __launch_bounds__(128) slic_iter_kernel(const uint8_t* __restrict__ rgb, int W, int H, int mx, int my,
                                                        int it, const int* __restrict__ sums_prev,
                                                        const float* __restrict__ ctr_prev, float* __restrict__ ctr_cur,
                                                        int* __restrict__ sums_cur, float coh_weight, float max_xy_dist,
                                                        float max_color_dist, int* __restrict__ labels) {
  __shared__ float sc[9][5];
  __shared__ int sv[9];
  __shared__ int sacc[9][6];
  const int cx = blockIdx.x, cy = blockIdx.y, tid = threadIdx.y * 16 + threadIdx.x;
  if (tid < 9) {
    const int cxc = cx + tid % 3 - 1, cyc = cy + tid / 3 - 1;
    const bool ok = cxc >= 0 && cyc >= 0 && cxc < mx && cyc < my;
    sv[tid] = ok ? cyc * mx + cxc : -1;
    if (ok) {
      const int s = cyc * mx + cxc;
      float c[5];
      if (it == 0) {  // gSLICr Init_Cluster_Centers
        int ix = cxc * kSp + kSp / 2, iy = cyc * kSp + kSp / 2;
        ix = ix >= W ? (cxc * kSp + W) / 2 : ix;
        iy = iy >= H ? (cyc * kSp + H) / 2 : iy;
        c[0] = (float)ix;
        c[1] = (float)iy;
        for (int k = 0; k < 3; ++k) c[2 + k] = (float)rgb[(iy * W + ix) * 3 + k];
      } else {  // Update_Cluster_Center
        const int* a = sums_prev + s * 6;
        if (a[5] != 0) {
          const float n = (float)a[5];
          for (int k = 0; k < 5; ++k) c[k] = (float)a[k] / n;
        } else {
          for (int k = 0; k < 5; ++k) c[k] = ctr_prev[s * 5 + k];
        }
      }
      for (int k = 0; k < 5; ++k) sc[tid][k] = c[k];
      if (tid == 4)
        for (int k = 0; k < 5; ++k) ctr_cur[s * 5 + k] = c[k];
    }
  }
  if (tid < 54) (&sacc[0][0])[tid] = 0;
  __syncthreads();
  // 16 x 8 threads, two pixels each (rows ty and ty + 8): 16 CTAs fit one SM, so the 1200 cells of a
  // VGA frame are a single wave
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int x = cx * kSp + threadIdx.x, y = cy * kSp + threadIdx.y + 8 * pass;
    const uint8_t* p = rgb + (y * W + x) * 3;
    const int q0 = p[0], q1 = p[1], q2 = p[2];
    const float p0 = (float)q0, p1 = (float)q1, p2 = (float)q2;
    int best = 4;
    float dist = 999999.9999f;
#pragma unroll
    for (int c = 0; c < 9; ++c) {  // same visiting order as gSLICr: rows -1..1, columns -1..1
      if (sv[c] >= 0) {
        const float dcolor = (p0 - sc[c][2]) * (p0 - sc[c][2]) + (p1 - sc[c][3]) * (p1 - sc[c][3]) +
                             (p2 - sc[c][4]) * (p2 - sc[c][4]);
        const float dxy =
            ((float)x - sc[c][0]) * ((float)x - sc[c][0]) + ((float)y - sc[c][1]) * ((float)y - sc[c][1]);
        const float cdist = sqrtf(dcolor * max_color_dist + coh_weight * dxy * max_xy_dist);
        if (cdist < dist) {
          dist = cdist;
          best = c;
        }
      }
    }
    labels[y * W + x] = sv[best];
    // integer sums of this assignment, warp-aggregated per chosen centre
    const unsigned peers = __match_any_sync(0xffffffffu, best);
    const int sx = __reduce_add_sync(peers, x), sy = __reduce_add_sync(peers, y), s0 = __reduce_add_sync(peers, q0),
              s1 = __reduce_add_sync(peers, q1), s2 = __reduce_add_sync(peers, q2);
    if ((tid & 31) == __ffs(peers) - 1) {
      atomicAdd(&sacc[best][0], sx);
      atomicAdd(&sacc[best][1], sy);
      atomicAdd(&sacc[best][2], s0);
      atomicAdd(&sacc[best][3], s1);
      atomicAdd(&sacc[best][4], s2);
      atomicAdd(&sacc[best][5], __popc(peers));
    }
  }
  __syncthreads();
  if (tid < 54) {
    const int c = tid / 6, k = tid % 6, v = sacc[c][k];
    if (v != 0 && sv[c] >= 0) atomicAdd(&sums_cur[sv[c] * 6 + k], v);
  }
}

// --------------------------------------------------------------------------- Slic.h reductions
// Float sums in PIXEL INDEX ORDER (Slic.h:64-72, :103-108).  One warp per super-pixel walks its
// 3x3-cell window row by row: the lanes load 32 labels / values at once, a ballot finds the pixels
// of this super-pixel and the (warp-uniform) accumulation visits them in raster order through
// shuffles -- the rounding sequence of the reference's loop with ~1/9 of the chain length.
// map 0 = thresholded depth, then icp / confidence per model.
struct DsMap {
  const float* img;
  int channels, channel;
  float threshold;  // < 0: plain downsample, >= 0: downsampleThresholded
};
constexpr int kMaxMaps = 2 * SegLimits::kMaxModels + 1;
struct DsArgs {
  DsMap m[kMaxMaps];
  int nmaps;
};
// sum += v of every lane whose bit is set in `mask`, in lane order.  The shuffles do not depend on the
// running sum, so they pipeline; only the (predicated) adds form the serial chain.
__device__ __forceinline__ void ordered_add(float& sum, float v, unsigned mask) {
#pragma unroll
  for (int b = 0; b < 32; ++b) {
    const float t = __shfl_sync(0xffffffffu, v, b);
    if ((mask >> b) & 1u) sum += t;
  }
}
__global__ void __launch_bounds__(256) spixel_sum_kernel(const DsArgs a, const int* __restrict__ labels, int W, int H,
                                                         int mx, int my, float* __restrict__ sums /* [nmaps][N] */,
                                                         unsigned* __restrict__ dcounts) {
  const int lane = threadIdx.x & 31, s = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int N = mx * my, mi = blockIdx.y;
  if (s >= N) return;
  const int gx = s % mx, gy = s / mx;
  const int x0 = max((gx - 1) * kSp, 0), x1 = min((gx + 2) * kSp, W), y0 = max((gy - 1) * kSp, 0),
            y1 = min((gy + 2) * kSp, H);
  const DsMap mp = a.m[mi];
  float sum = 0.f;
  int cnt0 = 0;
  constexpr int RB = 4;  // rows in flight: all their loads are issued before the ordered accumulation
  for (int y = y0; y < y1; y += RB) {
    bool match[RB][2];
    bool any = false;
#pragma unroll
    for (int r = 0; r < RB; ++r)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int x = x0 + c * 32 + lane, yy = y + r;
        match[r][c] = yy < y1 && x < x1 && __ldg(labels + yy * W + x) == s;
        any |= match[r][c];
      }
    if (!__any_sync(0xffffffffu, any)) continue;
    float val[RB][2];
#pragma unroll
    for (int r = 0; r < RB; ++r)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int x = x0 + c * 32 + lane, yy = y + r;
        float v = match[r][c] ? __ldg(mp.img + (size_t)(yy * W + x) * mp.channels + mp.channel) : 0.f;
        if (mp.threshold >= 0.f) {  // values at or below the threshold are skipped
          match[r][c] = match[r][c] && v > mp.threshold;
          cnt0 += match[r][c];
        }
        val[r][c] = v;
      }
#pragma unroll
    for (int r = 0; r < RB; ++r)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const unsigned mask = __ballot_sync(0xffffffffu, match[r][c]);
        if (mask) ordered_add(sum, val[r][c], mask);
      }
  }
  cnt0 = __reduce_add_sync(0xffffffffu, cnt0);
  if (lane == 0) {
    sums[mi * N + s] = sum;
    if (mi == 0) dcounts[s] = (unsigned)cnt0;
  }
}

struct SegState {
  int W, H, mx, my, N, numModels, numLabels, allowNew;
  unsigned char modelIds[SegLimits::kMaxModels + 1];
  unsigned char nextModelID;
  SegParams prm;
};

__device__ int resample_empty_index(const SegState& st, const int* labels, unsigned index) {
  // Slic.h:193-209 incl. the index / spixelY quirk
  const int x = (int)(index % (unsigned)st.mx), y = (int)(index / (unsigned)st.my);
  int cx = (int)(x * kSp + kSp * 0.5), cy = (int)(y * kSp + kSp * 0.5);
  if (cy >= st.H) cy = st.H - 1;
  if (cx >= st.W) cx = st.W - 1;
  return labels[cx + cy * st.W];
}

// Low-res maps, depth range, average confidences, unaries and CRF features: one CTA.  The division
// by the pixel count is independent per super-pixel; the reference's in-place resolution of EMPTY
// super-pixels (Slic.h:74-83, :110-123) is order dependent and replayed by one thread per map, but
// only over the (normally zero) empties: an earlier index sees the divided value, a later one the
// raw sum, exactly as the in-place loop would.
__global__ void __launch_bounds__(1024) seg_lowres_kernel(const SegState st, const int* __restrict__ labels,
                                                          const int* __restrict__ slicSums,
                                                          const unsigned* __restrict__ dcounts,
                                                          const float* __restrict__ raw, float* __restrict__ low,
                                                          unsigned* __restrict__ counts, float* __restrict__ unary,
                                                          SegModelData* __restrict__ md, float* __restrict__ depthRangeOut,
                                                          const uint8_t* __restrict__ rgb, float* __restrict__ T2,
                                                          float* __restrict__ f6) {
  __shared__ int sEmpty[kMaxMaps];
  __shared__ float sMin[32], sMax[32];
  __shared__ float sRange;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = st.N, M = 1 + 2 * st.numModels, L = st.numLabels;
  const float MAX_DEPTH = 100;
  if (tid < M) sEmpty[tid] = 0;
  __syncthreads();
  for (int q = tid; q < M * N; q += blockDim.x) {
    const int mi = q / N, idx = q - mi * N;
    const int cnt = mi == 0 ? (int)dcounts[idx] : slicSums[idx * 6 + 5];
    if (cnt != 0)
      low[q] = raw[q] / (float)cnt;
    else
      atomicAdd(&sEmpty[mi], 1);
  }
  for (int i = tid; i < N; i += blockDim.x) counts[i] = (unsigned)slicSums[i * 6 + 5];
  __syncthreads();
  if (tid < M && sEmpty[tid] > 0) {
    const float* r = raw + tid * N;
    float* o = low + tid * N;
    for (int index = 0; index < N; index++) {
      const int own = tid == 0 ? (int)dcounts[index] : slicSums[index * 6 + 5];
      if (own != 0) continue;
      const int readIndex = resample_empty_index(st, labels, index);
      const int cnt = slicSums[readIndex * 6 + 5];  // thresholded variant: TOTAL count of the other (quirk)
      // what the in-place loop would find at readIndex: already divided only if it came earlier
      const float src = readIndex < index ? o[readIndex] : r[readIndex];
      o[index] = src / (float)cnt;
    }
  }
  __syncthreads();
  // depth range (Segmentation.cpp:166-178): min / max are order free
  float mn = 3.402823466e+38f, mxv = 0.f;
  for (int i = tid; i < N; i += blockDim.x) {
    const float d = low[i];
    if (d > MAX_DEPTH || d < 0 || !isfinite(d)) continue;
    if (mxv < d) mxv = d;
    if (mn > d) mn = d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mxv = fmaxf(mxv, __shfl_xor_sync(0xffffffffu, mxv, o));
  }
  if (lane == 0) {
    sMin[warp] = mn;
    sMax[warp] = mxv;
  }
  if (tid < L) {
    SegModelData z = {};
    z.top = 65535;
    z.left = 65535;
    z.id = st.modelIds[tid];
    md[tid] = z;
  }
  __syncthreads();
  if (tid == 0) {
    float a = sMin[0], b = sMax[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      a = fminf(a, sMin[w]);
      b = fmaxf(b, sMax[w]);
    }
    sRange = b - a;
    *depthRangeOut = b - a;
  }
  // average confidence (:195-207): index-order sum, one warp per model, NaN / inf texels -> 0
  if (warp < st.numModels) {
    float* conf = low + (2 + 2 * warp) * N;
    float avg = 0.f;
    for (int base = 0; base < N; base += 32) {
      const int j = base + lane;
      const float v = j < N ? conf[j] : 0.f;
      const bool fin = j < N && isfinite(v);
      if (j < N && !fin) conf[j] = 0.f;
      ordered_add(avg, v, __ballot_sync(0xffffffffu, fin));
    }
    if (lane == 0) md[warp].avgConfidence = avg / (float)N;
  }
  __syncthreads();
  const float depthRange = sRange;
  for (int k = tid; k < N; k += blockDim.x) {  // unaries (:237-298, :458-460), node major [k][label]
    float icp[SegLimits::kMaxModels], conf[SegLimits::kMaxModels], u[SegLimits::kMaxModels + 1];
#pragma unroll
    for (int i = 0; i < SegLimits::kMaxModels; ++i)
      if (i < st.numModels) {
        icp[i] = low[(1 + 2 * i) * N + k];
        conf[i] = low[(2 + 2 * i) * N + k];
      }
    const float lowDepth = low[k];
    if (conf[0] < 0.3f) icp[0] = depthRange * 0.01f;
#pragma unroll
    for (int i = 1; i < SegLimits::kMaxModels; i++)
      if (i < st.numModels && conf[i] <= 0.4f) icp[i] = depthRange * st.prm.unaryKError;
    float lowestError = icp[0] / depthRange;
#pragma unroll
    for (int i = 0; i < SegLimits::kMaxModels; i++)
      if (i < st.numModels) {
        const float error = icp[i] / depthRange;
        if (error < lowestError) lowestError = error;
        u[i] = st.prm.unaryWeightError * error;
        low[(1 + 2 * i) * N + k] = icp[i];  // getError() returns a reference: the map is modified in place
      }
#pragma unroll
    for (int l = 0; l <= SegLimits::kMaxModels; ++l)
      if (l < L) {
        float v = u[l < st.numModels ? l : 0];
        if (l == st.numModels) {  // the "new" label
          const float un = st.prm.unaryThresholdNew - st.prm.unaryWeightError * lowestError;
          v = un > 0.01f ? un : 0.01f;
        }
        unary[k * L + l] = v <= 1e-5f ? 1e-5f : v;
      }
    // CRF features (:436-452).  The smoothness kernel only depends on the grid offset: its values
    // go to a table T2[|dy|][|dx|] = exp(-((dx/2)^2 + (dy/2)^2)/2), all operands exact.
    const int i = k % st.mx, j = k / st.mx;
    {
      const float a = (float)i / 2.0f, b = (float)j / 2.0f;
      float d2 = 0.f;
      d2 += a * a;
      d2 += b * b;
      T2[k] = det_expf(-0.5f * d2);
    }
    // node record: 6 appearance features + the grid position (as int bits), 32 bytes
    f6[k * 8 + 0] = (float)i * st.prm.scaleFeaturesPos;
    f6[k * 8 + 1] = (float)j * st.prm.scaleFeaturesPos;
    // quirk: the FULL-RES rgb buffer is read with the LOW-RES index (Segmentation.cpp:445-447)
    f6[k * 8 + 2] = (float)rgb[k * 3 + 0] * st.prm.scaleFeaturesRGB;
    f6[k * 8 + 3] = (float)rgb[k * 3 + 1] * st.prm.scaleFeaturesRGB;
    f6[k * 8 + 4] = (float)rgb[k * 3 + 2] * st.prm.scaleFeaturesRGB;
    const float fd = lowDepth * st.prm.scaleFeaturesDepth;
    f6[k * 8 + 5] = fd < 100.0f ? fd : 100.0f;
    f6[k * 8 + 6] = __int_as_float(i);
    f6[k * 8 + 7] = __int_as_float(j);
  }
}

// ------------------------------------------------------------------------------- dense CRF
// Exact Gaussian kernels over the N super-pixels: k2(i,j) is a table lookup by grid offset,
// k6(i,j) = exp(-|f6_i - f6_j|^2 / 2).  Both N x N matrices are built once per frame (11.5 MB at
// 640x480, L2 resident) and streamed by the ten mean-field iterations, which then cost ~20
// instructions per pair instead of ~75.  Four warps per node; the sum over j follows the oracle's
// frozen order: 128 partial sums (partial k owns j = k, k+128, ...), an xor
// butterfly inside each group of 32, then (s0 + s1) + (s2 + s3).
struct CrfNode {  // 32-byte record per super-pixel: appearance features + grid position
  float f[6];
  int x, y;
};
static_assert(sizeof(CrfNode) == 32, "CrfNode layout");
__device__ __forceinline__ float warp_butterfly_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ CrfNode load_node(const CrfNode* __restrict__ nodes, int j) {
  const float4* p = reinterpret_cast<const float4*>(nodes + j);
  const float4 a = __ldg(p), b = __ldg(p + 1);
  CrfNode n;
  n.f[0] = a.x;
  n.f[1] = a.y;
  n.f[2] = a.z;
  n.f[3] = a.w;
  n.f[4] = b.x;
  n.f[5] = b.y;
  n.x = __float_as_int(b.z);
  n.y = __float_as_int(b.w);
  return n;
}
__device__ __forceinline__ void crf_pair(const CrfNode& r, const CrfNode& q, const float* __restrict__ T2, int mx,
                                         float& k2, float& k6) {
  float d2 = 0.f;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const float d = r.f[k] - q.f[k];
    d2 += d * d;
  }
  k6 = det_expf(-0.5f * d2);
  k2 = __ldg(T2 + abs(r.y - q.y) * mx + abs(r.x - q.x));
}
constexpr int kCrfRowsPerCta = 2;  // 8 warps: 2 nodes x 4 warps
__global__ void __launch_bounds__(256) crf_norm_kernel(const CrfNode* __restrict__ nodes, const float* __restrict__ T2,
                                                       int N, int mx, float* __restrict__ K2, float* __restrict__ K6,
                                                       float* __restrict__ n2, float* __restrict__ n6) {
  __shared__ float sh[kCrfRowsPerCta][4][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, row = warp >> 2, part = warp & 3;
  const int i = blockIdx.x * kCrfRowsPerCta + row;
  if (i < N) {
    const CrfNode r = load_node(nodes, i);
    float rs2 = 0.f, rs6 = 0.f;
#pragma unroll 2
    for (int j = part * 32 + lane; j < N; j += 128) {
      const CrfNode q = load_node(nodes, j);
      float k2, k6;
      crf_pair(r, q, T2, mx, k2, k6);
      K2[(size_t)i * N + j] = k2;
      K6[(size_t)i * N + j] = k6;
      rs2 += k2;
      rs6 += k6;
    }
    rs2 = warp_butterfly_sum(rs2);
    rs6 = warp_butterfly_sum(rs6);
    if (lane == 0) {
      sh[row][part][0] = rs2;
      sh[row][part][1] = rs6;
    }
  }
  __syncthreads();
  if (i < N && part == 0 && lane == 0) {
    const float t2 = (sh[row][0][0] + sh[row][1][0]) + (sh[row][2][0] + sh[row][3][0]);
    const float t6 = (sh[row][0][1] + sh[row][1][1]) + (sh[row][2][1] + sh[row][3][1]);
    n2[i] = 1.0f / sqrtf(t2 + 1e-20f);
    n6[i] = 1.0f / sqrtf(t6 + 1e-20f);
  }
}
// softmax of one node (DenseCRF::expAndNormalize) + the normalised copies the next apply consumes
template <int LP>
__device__ __forceinline__ void crf_softmax_store(const float (&t)[LP], int L, int i, float n2i, float n6i,
                                                  float* __restrict__ Q, float* __restrict__ nq2,
                                                  float* __restrict__ nq6) {
  float mxv = t[0];
#pragma unroll
  for (int l = 1; l < LP; ++l)
    if (l < L) mxv = t[l] > mxv ? t[l] : mxv;
  float v[LP], sum = 0.f;
#pragma unroll
  for (int l = 0; l < LP; ++l)
    if (l < L) {
      v[l] = det_expf(t[l] - mxv);
      sum += v[l];
    }
#pragma unroll
  for (int l = 0; l < LP; ++l)
    if (l < L) {
      const float q = v[l] / sum;
      Q[i * L + l] = q;
      nq2[i * L + l] = n2i * q;
      nq6[i * L + l] = n6i * q;
    }
}
template <int LP>
__global__ void __launch_bounds__(256) crf_init_kernel(const float* __restrict__ unary, int N, int L,
                                                       const float* __restrict__ n2, const float* __restrict__ n6,
                                                       float* __restrict__ Q, float* __restrict__ nq2,
                                                       float* __restrict__ nq6) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float t[LP];
#pragma unroll
  for (int l = 0; l < LP; ++l) t[l] = l < L ? -unary[i * L + l] : 0.f;
  crf_softmax_store<LP>(t, L, i, n2[i], n6[i], Q, nq2, nq6);
}
// one mean-field iteration: t1 = -unary - (-w2 * n2_i * sum_j K2_ij nq2_jl) - (-w6 * n6_i * sum_j K6_ij nq6_jl)
template <int LP>
__global__ void __launch_bounds__(256) crf_iter_kernel(const float* __restrict__ unary, const float* __restrict__ K2,
                                                       const float* __restrict__ K6, const float* __restrict__ n2,
                                                       const float* __restrict__ n6, const float* __restrict__ nq2in,
                                                       const float* __restrict__ nq6in, int N, int L, float w2, float w6,
                                                       float* __restrict__ Q, float* __restrict__ nq2out,
                                                       float* __restrict__ nq6out) {
  __shared__ float sh[kCrfRowsPerCta][4][2 * LP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, row = warp >> 2, part = warp & 3;
  const int i = blockIdx.x * kCrfRowsPerCta + row;
  if (i < N) {
    const float* __restrict__ k2r = K2 + (size_t)i * N;
    const float* __restrict__ k6r = K6 + (size_t)i * N;
    float a2[LP], a6[LP];
#pragma unroll
    for (int l = 0; l < LP; ++l) a2[l] = a6[l] = 0.f;
#pragma unroll 4
    for (int j = part * 32 + lane; j < N; j += 128) {
      const float k2 = __ldg(k2r + j), k6 = __ldg(k6r + j);
#pragma unroll
      for (int l = 0; l < LP; ++l)
        if (l < L) {
          a2[l] += k2 * __ldg(nq2in + j * L + l);
          a6[l] += k6 * __ldg(nq6in + j * L + l);
        }
    }
#pragma unroll
    for (int l = 0; l < LP; ++l)
      if (l < L) {  // warp-uniform branch: the shuffles inside are executed by all lanes
        const float s2 = warp_butterfly_sum(a2[l]), s6 = warp_butterfly_sum(a6[l]);
        if (lane == 0) {
          sh[row][part][2 * l] = s2;
          sh[row][part][2 * l + 1] = s6;
        }
      }
  }
  __syncthreads();
  if (i < N && part == 0 && lane == 0) {
    const float n2i = n2[i], n6i = n6[i];
    float t[LP];
#pragma unroll
    for (int l = 0; l < LP; ++l) {
      t[l] = 0.f;
      if (l < L) {
        const float s2 = (sh[row][0][2 * l] + sh[row][1][2 * l]) + (sh[row][2][2 * l] + sh[row][3][2 * l]);
        const float s6 = (sh[row][0][2 * l + 1] + sh[row][1][2 * l + 1]) + (sh[row][2][2 * l + 1] + sh[row][3][2 * l + 1]);
        float tt = -unary[i * L + l];
        tt -= -w2 * (n2i * s2);
        tt -= -w6 * (n6i * s6);
        t[l] = tt;
      }
    }
    crf_softmax_store<LP>(t, L, i, n2i, n6i, Q, nq2out, nq6out);
  }
}

// ------------------------------------------------------- argmax, components, post-processing
struct Comp {
  unsigned char label;
  int top, right, bottom, left, size;
};
__device__ int find_root(const int* roots, int i) {
  while (i != roots[i]) i = roots[i];
  return i;
}
// index-order accumulation over the super-pixels carrying `id`, warp uniform (see ordered_add)
template <class F>
__device__ __forceinline__ void for_each_labelled(const uint8_t* smap, const float* sdepth, int N, int id, int lane,
                                                  F&& f) {
  for (int base = 0; base < N; base += 32) {
    const int i = base + lane;
    const bool m = i < N && smap[i] == id;
    const float d = i < N ? sdepth[i] : 0.f;
    const unsigned mask = __ballot_sync(0xffffffffu, m);
    if (mask == 0) continue;
#pragma unroll
    for (int b = 0; b < 32; ++b) {
      const float t = __shfl_sync(0xffffffffu, d, b);
      if ((mask >> b) & 1u) f(t);
    }
  }
}
// One CTA; all working sets live in shared memory.  Thread 0 replays the sequential parts of the
// reference (two-pass union-find labelling, the list surgery on labelToComponents) so component ids
// and tie breaks are the reference's; argmax, relabelling and the per-label statistics are parallel.
__global__ void __launch_bounds__(1024) seg_post_kernel(const SegState st, const float* __restrict__ Q,
                                                        const float* __restrict__ lowDepth, uint8_t* __restrict__ map,
                                                        SegModelData* __restrict__ md, SegResultHeader* __restrict__ hdr) {
  extern __shared__ unsigned char smem_raw[];
  const int N = st.N, L = st.numLabels, rows = st.my, cols = st.mx;
  int* comp = (int*)smem_raw;
  int* roots = comp + N;
  int* mapping = roots + N;
  int* next = mapping + N;
  float* sdepth = (float*)(next + N);
  Comp* cc = (Comp*)(sdepth + N);
  int* head = (int*)(cc + N);
  int* tail = head + 256;
  uint8_t* smap = (uint8_t*)(tail + 256);
  __shared__ SegModelData smd[SegLimits::kMaxModels + 1];
  __shared__ int sWarpTot[32];
  __shared__ int sNcc;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < N; i += blockDim.x) {  // maxCoeff: first maximum (Segmentation.cpp:477-480)
    int best = 0;
    for (int l = 1; l < L; ++l)
      if (Q[i * L + l] > Q[i * L + best]) best = l;
    smap[i] = st.modelIds[best];
    sdepth[i] = lowDepth[i];
  }
  if (tid < 256) head[tid] = tail[tid] = -1;
  if (tid < L) smd[tid] = md[tid];
  for (int i = tid; i < N; i += blockDim.x) roots[i] = i;
  __syncthreads();
  // ---- ConnectedLabels.hpp:50-172 (two-pass union-find, 4-connectivity) in parallel.  The reference
  // numbers components by their smallest provisional id, i.e. by their FIRST pixel in raster order;
  // min-index propagation with pointer jumping finds exactly that pixel for every component, and a
  // prefix count over those pixels reproduces the numbering.
  for (;;) {
    int changed = 0;
    for (int i = tid; i < N; i += blockDim.x) {
      const uint8_t v = smap[i];
      const int x = i % cols, y = i / cols, l = roots[i];
      int m = l;
      if (x > 0 && smap[i - 1] == v) m = min(m, roots[i - 1]);
      if (x + 1 < cols && smap[i + 1] == v) m = min(m, roots[i + 1]);
      if (y > 0 && smap[i - cols] == v) m = min(m, roots[i - cols]);
      if (y + 1 < rows && smap[i + cols] == v) m = min(m, roots[i + cols]);
      m = min(m, roots[m]);
      if (m < l) {
        roots[i] = m;  // monotone, always a pixel of the same component: races only delay convergence
        changed = 1;
      }
    }
    if (!__syncthreads_or(changed)) break;
  }
  // rank of every component's first pixel = component id
  {
    const int per = (N + (int)blockDim.x - 1) / (int)blockDim.x, beg = tid * per, end = min(beg + per, N);
    int mine = 0;
    for (int i = beg; i < end; ++i) mine += roots[i] == i;
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) sWarpTot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = lane < (int)(blockDim.x >> 5) ? sWarpTot[lane] : 0, wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      sWarpTot[lane] = wi - w;
      if (lane == 31) sNcc = wi;
    }
    __syncthreads();
    int rank = sWarpTot[warp] + incl - mine;
    for (int i = beg; i < end; ++i)
      if (roots[i] == i) {
        mapping[i] = rank;
        Comp z;
        z.top = 2147483647;
        z.left = 2147483647;
        z.right = z.bottom = z.size = 0;
        z.label = smap[i];
        cc[rank] = z;
        rank++;
      }
    __syncthreads();
    for (int i = tid; i < N; i += blockDim.x) {
      const int c = mapping[roots[i]], x = i % cols, y = i / cols;
      comp[i] = c;
      atomicAdd(&cc[c].size, 1);
      atomicMin(&cc[c].top, y);
      atomicMax(&cc[c].bottom, y);
      atomicMin(&cc[c].left, x);
      atomicMax(&cc[c].right, x);
    }
    __syncthreads();
  }
  if (tid == 0) {
    const int ncc = sNcc;
    // ---- labelToComponents (ConnectedLabels.hpp:40-48) as per-label linked lists in component
    // order, then the list surgery of Segmentation.cpp:496-563
    for (int c = 0; c < ncc; ++c) {
      const int lab = cc[c].label;
      next[c] = -1;
      if (head[lab] < 0)
        head[lab] = c;
      else
        next[tail[lab]] = c;
      tail[lab] = c;
    }
    // the labels present are a subset of the L model ids: the first entry of the std::map is the
    // smallest id present
    int firstLabel = 256;
    for (int m = 0; m < L; ++m)
      if (head[smd[m].id] >= 0 && (int)smd[m].id < firstLabel) firstLabel = (int)smd[m].id;
    for (int m = 0; m < L; ++m) {  // onlyKeepLargest: earlier component wins ties
      const int lab = (int)smd[m].id;
      if (lab == firstLabel || head[lab] < 0) continue;
      int cur = head[lab];
      for (int c2 = next[cur]; c2 >= 0;) {
        const int nx = next[c2];
        if (cc[cur].size < cc[c2].size) {
          cc[cur].label = 255;
          cur = c2;
        } else
          cc[c2].label = 255;
        c2 = nx;
      }
      head[lab] = cur;
      next[cur] = -1;
    }
    if (st.allowNew) {
      const int minSize = (int)((float)N * st.prm.minRelSizeNew), maxSize = (int)((float)N * st.prm.maxRelSizeNew);
      for (int c = head[st.nextModelID]; c >= 0; c = next[c])
        if (cc[c].size < minSize || cc[c].size > maxSize) cc[c].label = 255;
    }
    for (int m = 0; m < L; ++m) {  // bounding boxes in unsigned shorts (:533-547)
      SegModelData* d = &smd[m];
      for (int c = head[d->id]; c >= 0; c = next[c]) {
        const Comp* s = &cc[c];
        if (s->left < (int)d->left) d->left = (unsigned short)s->left;
        if (s->top < (int)d->top) d->top = (unsigned short)s->top;
        if (s->right > (int)d->right) d->right = (unsigned short)s->right;
        if (s->bottom > (int)d->bottom) d->bottom = (unsigned short)s->bottom;
      }
      int px = (int)(d->left * kSp + kSp * 0.5), py = (int)(d->top * kSp + kSp * 0.5);
      d->left = (unsigned short)px;
      d->top = (unsigned short)py;
      px = (int)(d->right * kSp + kSp * 0.5);
      py = (int)(d->bottom * kSp + kSp * 0.5);
      d->right = (unsigned short)px;
      d->bottom = (unsigned short)py;
    }
    const unsigned borderSize = 20;
    for (int m = 0; m < L; ++m) {  // objects hugging the image border are dropped (:549-563)
      SegModelData* d = &smd[m];
      if (d->id == 0) continue;
      if ((d->top < borderSize && d->bottom < borderSize) || (d->left < borderSize && d->right < borderSize) ||
          (d->top > (unsigned)st.H - borderSize && d->bottom > (unsigned)st.H - borderSize) ||
          (d->left > (unsigned)st.W - borderSize && d->right > (unsigned)st.W - borderSize)) {
        for (int c = head[d->id]; c >= 0; c = next[c]) cc[c].label = 255;
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < N; i += blockDim.x) {
    const uint8_t v = cc[comp[i]].label;
    smap[i] = v;
    map[i] = v;
  }
  __syncthreads();
  // ---- depth statistics (:570-621): one warp per label, index-order sums
  if (warp < L) {
    SegModelData* d = &smd[warp];
    const int id = (int)d->id;
    float sumDepth = 0.f, sumDev = 0.f;
    unsigned cnt = 0;
    for_each_labelled(smap, sdepth, N, id, lane, [&](float v) {
      sumDepth += v;
      cnt++;
    });
    const unsigned superPixels = cnt;
    const float mean0 = cnt ? sumDepth / (float)cnt : 0;
    for_each_labelled(smap, sdepth, N, id, lane, [&](float v) { sumDev += fabsf(mean0 - v); });
    const float std0 = cnt ? sumDev / (float)cnt : 0;
    if (warp != 0) {  // outlier rejection for object models only (idx != 0)
      for_each_labelled(smap, sdepth, N, id, lane, [&](float v) {
        if ((double)v > 1.1 * (double)std0 + (double)mean0) {
          sumDepth -= v;
          sumDev -= fabsf(mean0 - v);
          cnt--;
        }
      });
    }
    if (lane == 0) {
      d->depthMean = cnt ? sumDepth / (float)cnt : 0;
      d->depthStd = cnt ? sumDev / (float)cnt : 0;
      d->superPixelCount = superPixels;
    }
  }
  __syncthreads();
  if (tid < L) md[tid] = smd[tid];
  if (tid == 0) {
    hdr->numModelData = L;
    hdr->hasNewLabel = 0;
    if (st.allowNew) {
      if (smd[st.numModels].superPixelCount > 0)
        hdr->hasNewLabel = 1;
      else
        hdr->numModelData = st.numModels;
    }
  }
}
__global__ void seg_upsample_kernel(const uint8_t* __restrict__ map, const int* __restrict__ labels, int n,
                                    uint8_t* __restrict__ full) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) full[i] = map[labels[i]];  // Slic::upsample (Slic.h:133-147)
}

inline unsigned cdiv(unsigned a, unsigned b) { return (a + b - 1) / b; }

template <class T>
bool dalloc(T** p, size_t n) {
  return cudaMalloc((void**)p, n * sizeof(T)) == cudaSuccess && cudaMemset(*p, 0, n * sizeof(T)) == cudaSuccess;
}

size_t post_smem_bytes(int N) { return (size_t)N * (5 * 4 + sizeof(Comp) + 1) + 2 * 256 * 4 + 16; }

}  // namespace

#define RET_IF(e)                       \
  do {                                  \
    cudaError_t e__ = (e);              \
    if (e__ != cudaSuccess) return e__; \
  } while (0)

void seg_default_params(SegParams* p) {
  // GUI defaults (GUI/Tools/GUI.h:212-227) that MainController pushes every frame (:463-473)
  p->crfIterations = 10;
  p->scaleFeaturesRGB = 1.0f / 10.0f;
  p->scaleFeaturesDepth = 1.0f / 0.9f;
  p->scaleFeaturesPos = 1.0f / 1.8f;
  p->weightAppearance = 7;
  p->weightSmoothness = 2;
  p->unaryThresholdNew = 5.5f;
  p->unaryKError = 0.0375f;
  p->unaryWeightError = 75.0f;
  p->maxRelSizeNew = 0.4f;
  p->minRelSizeNew = 0.015f;
}

Segmentation::Segmentation(int W_, int H_) : W(W_), H(H_) {
  mx = W / kSp;
  my = H / kSp;
  N = mx * my;
  const int Lmax = SegLimits::kMaxModels + 1;
  bool good = (W % kSp == 0) && (H % kSp == 0) && post_smem_bytes(N) <= 227 * 1024;
  good = good && dalloc(&labels, (size_t)W * H) && dalloc(&centers, (size_t)2 * N * 5) &&
         dalloc(&slicSums, (size_t)6 * N * 6) && dalloc(&counts, N) && dalloc(&dcounts, N) &&
         dalloc(&sums, (size_t)kMaxMaps * N) && dalloc(&low, (size_t)kMaxMaps * N) &&
         dalloc(&unary, (size_t)N * Lmax) && dalloc(&T2, (size_t)N) && dalloc(&f6, (size_t)N * 8) &&
         dalloc(&K2, (size_t)N * N) && dalloc(&K6, (size_t)N * N) && dalloc(&n2, N) && dalloc(&n6, N) &&
         dalloc(&Q, (size_t)N * Lmax) && dalloc(&nq2, (size_t)2 * N * Lmax) && dalloc(&nq6, (size_t)2 * N * Lmax) &&
         dalloc(&lowMap, N) && dalloc(&md, Lmax) && dalloc(&hdr, 1) && dalloc(&depthRange, 1);
  good = good && cudaMallocHost(&h_out, sizeof(SegResultHeader) + Lmax * sizeof(SegModelData)) == cudaSuccess;
  good = good && cudaFuncSetAttribute(seg_post_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)post_smem_bytes(N)) == cudaSuccess;
  ok_ = good;
}

Segmentation::~Segmentation() {
  void* ptrs[] = {labels, centers, slicSums, counts, dcounts, sums, low, unary, T2, f6,  K2,
                  K6,     n2,      n6,       Q,      nq2,     nq6,  lowMap, md, hdr, depthRange};
  for (void* p : ptrs) cudaFree(p);
  cudaFreeHost(h_out);
  for (int i = 0; i < kGraphSlots; ++i) {
    if (graphExec_[i]) cudaGraphExecDestroy((cudaGraphExec_t)graphExec_[i]);
    free(graphKey_[i]);
  }
}

cudaError_t Segmentation::slic(const uint8_t* rgb, cudaStream_t s) {
  float max_xy_dist = 1.0f / (1.4242f * (float)kSp);
  float max_color_dist = 5.0f / (1.7321f * 128);
  max_color_dist *= max_color_dist;
  max_xy_dist *= max_xy_dist;
  RET_IF(cudaMemsetAsync(slicSums, 0, sizeof(int) * 6 * N * 6, s));
  const dim3 b(kSp, kSp / 2), g(mx, my);
  for (int it = 0; it <= 5; ++it) {  // no_iters = 5 (Slic.cpp:39): 6 assignments, 5 centre updates
    const int* prev = it ? slicSums + (size_t)(it - 1) * N * 6 : nullptr;
    slic_iter_kernel<<<g, b, 0, s>>>(rgb, W, H, mx, my, it, prev, centers + (size_t)((it + 1) & 1) * N * 5,
                                     centers + (size_t)(it & 1) * N * 5, slicSums + (size_t)it * N * 6, 0.6f,
                                     max_xy_dist, max_color_dist, labels);
  }
  return cudaGetLastError();
}

namespace {
template <int LP>
void launch_crf(const Segmentation& g, int L, const SegParams& prm, cudaStream_t s) {
  const int N = g.N;
  const size_t half = (size_t)N * (SegLimits::kMaxModels + 1);
  crf_init_kernel<LP><<<cdiv(N, 256), 256, 0, s>>>(g.unary, N, L, g.n2, g.n6, g.Q, g.nq2, g.nq6);
  for (int it = 0; it < prm.crfIterations; ++it) {
    const size_t in = (size_t)(it & 1) * half, out = (size_t)((it + 1) & 1) * half;
    crf_iter_kernel<LP><<<cdiv(N, kCrfRowsPerCta), 256, 0, s>>>(g.unary, g.K2, g.K6, g.n2, g.n6, g.nq2 + in, g.nq6 + in, N, L,
                                                   prm.weightSmoothness, prm.weightAppearance, g.Q, g.nq2 + out,
                                                   g.nq6 + out);
  }
}
}  // namespace

// Everything one performSegmentationCRF enqueues: 1 memset, 6 + 5 + crfIterations kernels, 2 copies.
cudaError_t Segmentation::enqueue(const uint8_t* rgb, const float* depth, int numModels, const unsigned char* modelIds,
                                  const float* const* icpError, const float* const* vertConf4,
                                  unsigned char nextModelID, bool allowNew, const SegParams& prm, uint8_t* fullSeg,
                                  cudaStream_t s) {
  if (slicAheadOf != rgb) RET_IF(slic(rgb, s));
  SegState st;
  st.W = W;
  st.H = H;
  st.mx = mx;
  st.my = my;
  st.N = N;
  st.numModels = numModels;
  st.allowNew = allowNew ? 1 : 0;
  st.numLabels = numModels + st.allowNew;
  for (int m = 0; m <= SegLimits::kMaxModels; ++m) st.modelIds[m] = 0;
  for (int m = 0; m < numModels; ++m) st.modelIds[m] = modelIds[m];
  st.modelIds[numModels] = nextModelID;
  st.nextModelID = nextModelID;
  st.prm = prm;
  const int L = st.numLabels;
  DsArgs a;
  a.nmaps = 1 + 2 * numModels;
  a.m[0] = DsMap{depth, 1, 0, 0.02f};
  for (int m = 0; m < numModels; ++m) {
    a.m[1 + 2 * m] = DsMap{icpError[m], 1, 0, -1.f};
    a.m[2 + 2 * m] = DsMap{vertConf4[m], 4, 3, -1.f};
  }
  for (int m = a.nmaps; m < kMaxMaps; ++m) a.m[m] = a.m[0];
  const int* finalSums = slicSums + (size_t)5 * N * 6;
  spixel_sum_kernel<<<dim3(cdiv(N, 8), a.nmaps), 256, 0, s>>>(a, labels, W, H, mx, my, sums, dcounts);
  seg_lowres_kernel<<<1, 1024, 0, s>>>(st, labels, finalSums, dcounts, sums, low, counts, unary, md, depthRange, rgb,
                                       T2, f6);
  crf_norm_kernel<<<cdiv(N, kCrfRowsPerCta), 256, 0, s>>>((const CrfNode*)f6, T2, N, mx, K2, K6, n2, n6);
  if (L <= 2)
    launch_crf<2>(*this, L, prm, s);
  else if (L <= 4)
    launch_crf<4>(*this, L, prm, s);
  else if (L <= 8)
    launch_crf<8>(*this, L, prm, s);
  else
    launch_crf<16>(*this, L, prm, s);
  seg_post_kernel<<<1, 1024, post_smem_bytes(N), s>>>(st, Q, low, lowMap, md, hdr);
  seg_upsample_kernel<<<cdiv(W * H, 256), 256, 0, s>>>(lowMap, labels, W * H, fullSeg);
  RET_IF(cudaGetLastError());
  SegResultHeader* hh = (SegResultHeader*)h_out;
  SegModelData* hm = (SegModelData*)((char*)h_out + sizeof(SegResultHeader));
  RET_IF(cudaMemcpyAsync(hh, hdr, sizeof(SegResultHeader), cudaMemcpyDeviceToHost, s));
  RET_IF(cudaMemcpyAsync(hm, md, sizeof(SegModelData) * L, cudaMemcpyDeviceToHost, s));
  return cudaSuccess;
}

namespace {
struct GraphKey {  // every argument baked into the captured launches
  const void *rgb, *depth, *fullSeg;
  const void* maps[2 * SegLimits::kMaxModels];
  unsigned char ids[SegLimits::kMaxModels + 1];
  int numModels, allowNew, slicAhead;
  SegParams prm;
};
bool same_key(const GraphKey& a, const GraphKey& b) {
  if (a.rgb != b.rgb || a.depth != b.depth || a.fullSeg != b.fullSeg || a.numModels != b.numModels ||
      a.allowNew != b.allowNew || a.slicAhead != b.slicAhead || memcmp(&a.prm, &b.prm, sizeof(SegParams)) != 0)
    return false;
  for (int m = 0; m < a.numModels; ++m)
    if (a.maps[2 * m] != b.maps[2 * m] || a.maps[2 * m + 1] != b.maps[2 * m + 1] || a.ids[m] != b.ids[m]) return false;
  return a.ids[a.numModels] == b.ids[b.numModels];
}
}  // namespace

cudaError_t Segmentation::performSegmentationCRF(const uint8_t* rgb, const float* depth, int numModels,
                                                 const unsigned char* modelIds, const float* const* icpError,
                                                 const float* const* vertConf4, unsigned char nextModelID,
                                                 bool allowNew, const SegParams& prm, uint8_t* fullSeg,
                                                 SegModelData* md_host, int* md_count, bool* hasNew, cudaStream_t s) {
  if (numModels < 1 || numModels > SegLimits::kMaxModels) return cudaErrorInvalidValue;
  // The launch sequence only changes when a model is spawned / lost or a parameter moves: replay it as
  // a CUDA graph (the legacy default stream cannot be captured -> plain launches there).
  bool launched = false;
  if (useGraph && s != nullptr && s != cudaStreamLegacy && s != cudaStreamPerThread) {
    GraphKey k;
    memset(&k, 0, sizeof(k));
    k.rgb = rgb;
    k.depth = depth;
    k.fullSeg = fullSeg;
    k.numModels = numModels;
    k.allowNew = allowNew ? 1 : 0;
    k.slicAhead = slicAheadOf == rgb ? 1 : 0;
    k.prm = prm;
    for (int m = 0; m < numModels; ++m) {
      k.maps[2 * m] = icpError[m];
      k.maps[2 * m + 1] = vertConf4[m];
      k.ids[m] = modelIds[m];
    }
    k.ids[numModels] = nextModelID;
    // a few cached graphs: the frame buffers of a context alternate (double-buffered upload)
    int slot = -1;
    for (int i = 0; i < kGraphSlots; ++i)
      if (graphExec_[i] && graphKey_[i] && same_key(*(GraphKey*)graphKey_[i], k)) slot = i;
    if (slot < 0) {
      slot = graphNext_;
      graphNext_ = (graphNext_ + 1) % kGraphSlots;
      if (graphExec_[slot]) cudaGraphExecDestroy((cudaGraphExec_t)graphExec_[slot]);
      graphExec_[slot] = nullptr;
      cudaGraph_t graph = nullptr;
      RET_IF(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
      const cudaError_t ce =
          enqueue(rgb, depth, numModels, modelIds, icpError, vertConf4, nextModelID, allowNew, prm, fullSeg, s);
      const cudaError_t ee = cudaStreamEndCapture(s, &graph);
      if (ce != cudaSuccess) return ce;
      RET_IF(ee);
      cudaGraphExec_t exec = nullptr;
      RET_IF(cudaGraphInstantiate(&exec, graph, 0));
      cudaGraphDestroy(graph);
      graphExec_[slot] = exec;
      if (!graphKey_[slot]) graphKey_[slot] = malloc(sizeof(GraphKey));
      *(GraphKey*)graphKey_[slot] = k;
    }
    RET_IF(cudaGraphLaunch((cudaGraphExec_t)graphExec_[slot], s));
    launched = true;
  }
  if (!launched)
    RET_IF(enqueue(rgb, depth, numModels, modelIds, icpError, vertConf4, nextModelID, allowNew, prm, fullSeg, s));
  slicAheadOf = nullptr;
  RET_IF(cudaStreamSynchronize(s));
  const SegResultHeader* hh = (const SegResultHeader*)h_out;
  const SegModelData* hm = (const SegModelData*)((const char*)h_out + sizeof(SegResultHeader));
  if (md_count) *md_count = hh->numModelData;
  if (hasNew) *hasNew = hh->hasNewLabel != 0;
  if (md_host) memcpy(md_host, hm, sizeof(SegModelData) * hh->numModelData);
  launches = 6 + 1 + 1 + 1 + 1 + prm.crfIterations + 2;
  return cudaSuccess;
}

}  // namespace cfb
