// cfb_common.cuh -- shared device/host helpers for the cofusion_b200 CUDA module (sm_100a).
//
// Design notes (DESIGN.md has the long form):
//  * every map is planar f32 with an explicit pitch in BYTES, the layout contract of the reference's
//    DeviceArray2D (Core/Cuda/containers/kernel_containers.hpp:60-93, reduce.cu:287-289);
//  * reductions never use a second launch: warp "transpose" reduction (31 SHFL for 32 values instead
//    of 32x5), per-block partials, and the LAST block to finish (atomic ticket) folds the partials in
//    a fixed order -> deterministic sums, no float atomics, no host round trip.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cfb {

struct Mat33 {  // row-major, same memory image as the reference mat33 (Core/Cuda/types.cuh:61-73)
  float m[9];
};
struct Vec3 {
  float x, y, z;
};
struct Intr {  // CameraModel (types.cuh:83-99)
  float fx, fy, cx, cy;
  __host__ __device__ Intr level(int l) const {
    int div = 1 << l;
    return Intr{fx / div, fy / div, cx / div, cy / div};
  }
};

// DataTerm (types.cuh:75-81): 16 B, valid at offset 12
struct DataTerm {
  short2 zero;
  short2 one;
  float diff;
  bool valid;
};
static_assert(sizeof(DataTerm) == 16, "DataTerm must match the reference layout");

__host__ __device__ __forceinline__ float qnan() {
#ifdef __CUDA_ARCH__
  return __int_as_float(0x7fffffff);
#else
  union {
    unsigned u;
    float f;
  } c;
  c.u = 0x7fffffffu;
  return c.f;
#endif
}

__device__ __forceinline__ float3 operator-(float3 a, float3 b) {
  return make_float3(a.x - b.x, a.y - b.y, a.z - b.z);
}
__device__ __forceinline__ float3 operator+(float3 a, float3 b) {
  return make_float3(a.x + b.x, a.y + b.y, a.z + b.z);
}
__device__ __forceinline__ float3 cross(float3 a, float3 b) {
  return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float norm(float3 a) { return sqrtf(dot(a, a)); }
__device__ __forceinline__ float3 normalized(float3 a) {
  float rn = rsqrtf(dot(a, a));
  return make_float3(a.x * rn, a.y * rn, a.z * rn);
}
__device__ __forceinline__ float3 mul(const Mat33& m, float3 a) {
  return make_float3(m.m[0] * a.x + m.m[1] * a.y + m.m[2] * a.z,
                     m.m[3] * a.x + m.m[4] * a.y + m.m[5] * a.z,
                     m.m[6] * a.x + m.m[7] * a.y + m.m[8] * a.z);
}

// ---- programmatic dependent launch (PDL).  A frame is a chain of ~25 small kernels on one stream; between two
// of them the GPU idles for the few microseconds the next launch takes to start.  Every kernel of the chain begins
// with pdl_prologue(): it lets the NEXT kernel of the stream start launching right away (its CTAs become resident
// and park at their own griddepcontrol.wait) and then waits until the PREVIOUS kernel has completed and its writes
// are visible.  Data dependencies are untouched -- nothing is read or written before the wait -- only the launch
// latency moves off the critical path.  Launched without the attribute (or on older parts) both are no-ops.
__device__ __forceinline__ void pdl_prologue() {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 900
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
bool pdl_enabled();  // false when CFB_NO_PDL is set (cabi.cu) or switched off for the calling thread's frame
// Early-launched kernels park on the SMs until their predecessor ends.  On ONE stream that is free; with several
// streams in flight (a frame with object models, segmentation) the parked CTAs take the slots another stream's
// kernel could run in -- measured: 899 -> 711 frames/s on the 4-object scene.  CoFusion switches it per frame.
void pdl_set(bool on);
#define CFB_PDL(e)                      \
  do {                                  \
    cudaError_t e__ = (e);              \
    if (e__ != cudaSuccess) return e__; \
  } while (0)
template <class... P, class... A>
inline cudaError_t launch_pdl(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}

template <class T>
__device__ __forceinline__ T* row_ptr(T* base, size_t pitch_bytes, int y) {
  return (T*)((char*)base + (size_t)y * pitch_bytes);
}
template <class T>
__device__ __forceinline__ const T* row_ptr(const T* base, size_t pitch_bytes, int y) {
  return (const T*)((const char*)base + (size_t)y * pitch_bytes);
}

// ---------------------------------------------------------------------------------------------
// Warp transpose-reduction: every lane holds v[0..31]; on return lane L holds sum over lanes of v[L].
// 16+8+4+2+1 = 31 shuffles (vs 160 for 32 independent butterfly reductions).
template <int HALF>
__device__ __forceinline__ void wtr_step(float (&v)[32], unsigned lane) {
  const bool upper = (lane & HALF) != 0;
#pragma unroll
  for (int i = 0; i < HALF; ++i) {
    float keep = upper ? v[i + HALF] : v[i];
    float send = upper ? v[i] : v[i + HALF];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, HALF);
  }
}
__device__ __forceinline__ float warp_transpose_reduce32(float (&v)[32]) {
  const unsigned lane = threadIdx.x & 31;
  wtr_step<16>(v, lane);
  wtr_step<8>(v, lane);
  wtr_step<4>(v, lane);
  wtr_step<2>(v, lane);
  wtr_step<1>(v, lane);
  return v[0];  // lane L: total of index L
}

// Block-level: returns (in warp 0, lane L) the block total of index L.  smem: [nwarps][32] floats.
__device__ __forceinline__ float block_reduce32(float (&v)[32], float* smem) {
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float t = warp_transpose_reduce32(v);
  smem[warp * 32 + lane] = t;
  __syncthreads();
  float s = 0.f;
  if (warp == 0)
    for (unsigned w = 0; w < nw; ++w) s += smem[w * 32 + lane];
  return s;
}

// Grid-level fixed-order finalisation.  Every block writes its 32 partials to partials[block][32];
// the last block to arrive sums rows in a fixed order and returns true (in all its threads) with
// the grand total of index L available via out32[L] (shared memory, 32 floats).
// `ticket` must be zero before the launch and is reset to zero by the last block.
__device__ __forceinline__ bool grid_finalize32(float block_total, float* partials, unsigned* ticket,
                                                float* smem /* >= nwarps*32 */, float* out32) {
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __shared__ bool is_last;
  if (warp == 0) {
    partials[blockIdx.x * 32 + lane] = block_total;
    __threadfence();
    if (lane == 0) {
      unsigned t = atomicAdd(ticket, 1u);
      is_last = (t == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (!is_last) return false;
  __threadfence();
  float s = 0.f;
  {  // independent loads, 4 in flight per warp; fixed summation order
    unsigned b = warp;
    for (; b + 3 * nw < gridDim.x; b += 4 * nw) {
      float p0 = __ldcg(&partials[b * 32 + lane]), p1 = __ldcg(&partials[(b + nw) * 32 + lane]);
      float p2 = __ldcg(&partials[(b + 2 * nw) * 32 + lane]), p3 = __ldcg(&partials[(b + 3 * nw) * 32 + lane]);
      s += p0;
      s += p1;
      s += p2;
      s += p3;
    }
    for (; b < gridDim.x; b += nw) s += __ldcg(&partials[b * 32 + lane]);
  }
  __syncthreads();  // smem reuse
  smem[warp * 32 + lane] = s;
  __syncthreads();
  if (warp == 0) {
    float tot = 0.f;
    for (unsigned w = 0; w < nw; ++w) tot += smem[w * 32 + lane];
    out32[lane] = tot;
    if (lane == 0) *ticket = 0;
  }
  __syncthreads();
  return true;
}

}  // namespace cfb

#define CFB_CUDA_OK(expr)                                                         \
  do {                                                                            \
    cudaError_t e__ = (expr);                                                     \
    if (e__ != cudaSuccess) return cfb::set_error(e__, #expr, __FILE__, __LINE__); \
  } while (0)

namespace cfb {
int set_error(cudaError_t e, const char* what, const char* file, int line);
int set_error_msg(int code, const char* msg);
}  // namespace cfb
