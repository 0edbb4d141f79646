// tracker_kernels.cuh -- declarations of the tracker reduction kernels' launch wrappers.
#pragma once
#include "cfb_common.cuh"

namespace cfb {

constexpr int kMaxBlocks = 1024;

// Device scratch shared by all reduction steps (the `sum`/`out` DeviceArrays of the reference's
// free functions, Core/Cuda/cudafuncs.cuh:64-82).  Must be zero-initialised once.
struct StepScratch {
  unsigned ticket;           // last-block election, self-resetting
  unsigned pad0[31];
  int rgb_count;             // computeRgbResidual outputs (integer atomics -> order independent)
  int rgb_sigma;
  int pad1[30];
  float result[32];          // packed sums of the last step (29 SE3 / 11 SO3 used)
  float partials[kMaxBlocks * 32];  // ONE row of 32 partial sums per block: grid_for() caps every grid at kMaxBlocks
};

// Pose block read by the ICP kernel from device memory (so the device-resident GN loop can update
// it without a host round trip).
struct IcpPose {
  Mat33 Rcurr;
  float tcurr[3];
  Mat33 Rprev_inv;
  float tprev[3];
};
struct RgbWarp {
  Mat33 krkinv;
  float kt[3];
};

struct PlanarMap {  // 3 planes of `rows` rows, pitch in bytes
  const float* p;
  size_t pitch;
};

struct IcpArgs {
  PlanarMap vmap_curr, nmap_curr, vmap_g_prev, nmap_g_prev;
  Intr intr;
  float distThres, angleThres;
  int cols, rows;
  float* error_map;  // optional, pitch error_pitch
  size_t error_pitch;
};

struct RgbResidualArgs {
  float minScale, maxDepthDelta;
  const short *dIdx, *dIdy;
  size_t grad_pitch;
  const float *lastDepth, *nextDepth;
  size_t depth_pitch;
  const unsigned char *lastImage, *nextImage;
  size_t img_pitch;
  DataTerm* corres;  // unpitched cols*rows (reduce.cu:862 indexes data[k])
  int cols, rows;
};

struct RgbStepArgs {
  const DataTerm* corres;
  const float* cloud;  // AoS float3, pitch cloud_pitch bytes
  size_t cloud_pitch;
  const short *dIdx, *dIdy;
  size_t grad_pitch;
  float fx, fy, sobelScale;
  int cols, rows;
};

struct So3Args {
  const unsigned char *lastImage, *nextImage;
  size_t img_pitch;
  Mat33 imageBasis, kinv, krlr;
  int cols, rows;
};

int num_sms();
// Each launches ONE kernel on `stream`; results land in scratch->result (device).
cudaError_t launch_icp_step(const IcpArgs& a, const IcpPose* d_pose, StepScratch* scratch,
                            cudaStream_t stream);
cudaError_t launch_rgb_residual(const RgbResidualArgs& a, const RgbWarp* d_warp, StepScratch* scratch,
                                cudaStream_t stream);
// sigma < -1.5 means "derive sigma from scratch->rgb_count/rgb_sigma on the device"
// (RGBDOdometry.cpp:373-374), used by the device-resident loop.
cudaError_t launch_rgb_step(const RgbStepArgs& a, float sigma, StepScratch* scratch,
                            cudaStream_t stream);
cudaError_t launch_so3_step(const So3Args& a, StepScratch* scratch, cudaStream_t stream);

}  // namespace cfb
