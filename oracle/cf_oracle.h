/*
 * cf_oracle.h -- CPU oracle for the Co-Fusion per-frame hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * algorithm (martinruenz/co-fusion @ 11b9fef), written to be the checker for
 * the CUDA product under cofusion_b200/.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may link or call it.
 * The product never routes through this code.
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   - tracker kernels (image prep, icp/rgb/so3 steps): pinned against the
 *     reference's own Core/Cuda/{reduce,cudafuncs}.cu compiled unmodified for
 *     sm_100a (oracle/_ref, built by oracle/build_ref.sh) -- fixtures under
 *     tests/golden/ are generated from that build on a B200.
 *   - GL predict/fuse/clean and the CRF segmentation: PARITY UNPINNED.  The
 *     reference has no tests/golden vectors and no GL
 *     context / gSLICr / densecrf exist in this image, so this restatement
 *     itself defines the expected result.
 *
 * Conventions: all images row-major, unpitched.  "Planar" maps are 3 planes of
 * H rows each ([k*H + y][x]), exactly the reference layout with pitch == W*4
 * (Core/Cuda/reduce.cu:287-289).  float4 images are AoS (x,y,z,w).
 */
#ifndef CF_ORACLE_H_
#define CF_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Core/Cuda/types.cuh:75-81 -- 16 bytes, `valid` at offset 12. */
typedef struct {
  int16_t zero_x, zero_y;
  int16_t one_x, one_y;
  float diff;
  uint8_t valid;
  uint8_t pad_[3];
} OrcDataTerm;

/* ---------------- image preparation (Core/Cuda/cudafuncs.cu, GLSL) -------- */
void orc_bilateral_filter(const float* depth, int W, int H, float maxD, float* out);
void orc_pyr_down_gauss_f(const float* src, int sw, int sh, float* dst);
void orc_pyr_down_uchar_gauss(const uint8_t* src, int sw, int sh, uint8_t* dst);
void orc_create_vmap(const float* depth, int W, int H, float fx, float fy, float cx, float cy,
                     float cutoff, float* vmap);
void orc_create_nmap(const float* vmap, int W, int H, float* nmap);
void orc_copy_maps(const float* v4, const float* n4, int W, int H, float* vmap, float* nmap);
void orc_resize_map(const float* in, int sw, int sh, int normalize, float* out);
void orc_transform_maps(const float* vsrc, const float* nsrc, int W, int H, const float R[9],
                        const float t[3], float* vdst, float* ndst);
void orc_vertices_to_depth(const float* v4, int W, int H, float cutoff, float* depth);
void orc_rgb_to_intensity(const uint8_t* rgb, int channels, int W, int H, uint8_t* grey);
void orc_derivative_images(const uint8_t* img, int W, int H, int16_t* dx, int16_t* dy);
void orc_project_to_point_cloud(const float* depth, int W, int H, float fx, float fy, float cx,
                                float cy, float* cloud3);

/* ---------------- tracker steps (Core/Cuda/reduce.cu) --------------------- */
void orc_icp_step(const float Rcurr[9], const float tcurr[3], const float* vmap_curr,
                  const float* nmap_curr, const float Rprev_inv[9], const float tprev[3], float fx,
                  float fy, float cx, float cy, const float* vmap_g_prev, const float* nmap_g_prev,
                  float distThres, float angleThres, int W, int H, float A[36], float b[6],
                  float residual[2], float* error_map /* may be NULL */);
void orc_rgb_residual(float minScale, const int16_t* dIdx, const int16_t* dIdy,
                      const float* lastDepth, const float* nextDepth, const uint8_t* lastImage,
                      const uint8_t* nextImage, OrcDataTerm* corres, float maxDepthDelta,
                      const float kt[3], const float krkinv[9], int W, int H, int* sigmaSum,
                      int* count);
void orc_rgb_step(const OrcDataTerm* corres, float sigma, const float* cloud3, float fx, float fy,
                  const int16_t* dIdx, const int16_t* dIdy, float sobelScale, int W, int H,
                  float A[36], float b[6]);
void orc_so3_step(const uint8_t* lastImage, const uint8_t* nextImage, const float imageBasis[9],
                  const float kinv[9], const float krlr[9], int W, int H, float A[9], float b[3],
                  float residual[2]);

/* ---------------- RGBDOdometry restatement (Core/Utils/RGBDOdometry.cpp) -- */
typedef struct OrcOdometry OrcOdometry;

typedef struct {
  float lastICPError, lastICPCount, lastRGBError, lastRGBCount, lastSO3Error, lastSO3Count;
  double lastA[36];
  double lastb[6];
  int so3_iterations;
} OrcTrackStats;

OrcOdometry* orc_odom_create(int W, int H, float cx, float cy, float fx, float fy, float distThresh,
                             float angleThresh);
void orc_odom_destroy(OrcOdometry* o);
/* initICPModel + initRGBModel: model prediction (AoS float4 vertex/normal in the camera frame of
 * `pose`, RGBA8 or RGB8 image).  Pose is row-major 4x4 camera->world. */
void orc_odom_init_model(OrcOdometry* o, const float* v4, const float* n4, const uint8_t* img,
                         int img_channels, const float pose[16]);
/* initICP(depth pyramid) + initRGB(rgb): `depth_filtered` is level 0; the oracle builds the
 * pyramid with orc_pyr_down_gauss_f as Model::generateCUDATextures does. */
void orc_odom_init_frame(OrcOdometry* o, const float* depth_filtered, const uint8_t* rgb,
                         int img_channels, float depthCutoff);
void orc_odom_init_first_rgb(OrcOdometry* o, const uint8_t* rgb, int img_channels);
/* getIncrementalTransformation: trans[3], rot[9] (row-major) in/out. error_map may be NULL. */
void orc_odom_track(OrcOdometry* o, float trans[3], float rot[9], int rgbOnly, float icpWeight,
                    int pyramid, int fastOdom, int so3, float* icp_error_map, OrcTrackStats* stats);
/* Same loop with the four reduction steps supplied by a backend (the compiled reference CUDA
 * kernels in oracle/_ref use this to be driven through the reference's own call sequence). */
typedef struct OrcStepBackend {
  void* user;
  void (*begin)(void* user, OrcOdometry* o);
  void (*so3_step)(void* user, OrcOdometry* o, int level, const float imageBasis[9],
                   const float kinv[9], const float krlr[9], float A[9], float b[3],
                   float residual[2]);
  void (*rgb_residual)(void* user, OrcOdometry* o, int level, float minScale, const float kt[3],
                       const float krkinv[9], int* sigma, int* count);
  void (*icp_step)(void* user, OrcOdometry* o, int level, const float Rcurr[9],
                   const float tcurr[3], const float Rprev_inv[9], const float tprev[3],
                   float A[36], float b[6], float residual[2], float* error_map);
  void (*rgb_step)(void* user, OrcOdometry* o, int level, float sigma, float A[36], float b[6]);
  void (*end)(void* user, OrcOdometry* o);
} OrcStepBackend;
void orc_odom_track_ex(OrcOdometry* o, float trans[3], float rot[9], int rgbOnly, float icpWeight,
                       int pyramid, int fastOdom, int so3, float* icp_error_map,
                       OrcTrackStats* stats, const OrcStepBackend* backend);
void orc_odom_dims(const OrcOdometry* o, int* W, int* H, float intr[4]);
/* Views into the oracle's pyramids, for fixture generation / unit parity. which:
 * 0 vmap_curr 1 nmap_curr 2 vmap_g_prev 3 nmap_g_prev 4 lastDepth 5 nextDepth 6 lastImage
 * 7 nextImage 8 dIdx 9 dIdy 10 lastNextImage 11 cloud */
const void* orc_odom_view(OrcOdometry* o, int which, int level);

#ifdef __cplusplus
}
#endif
#endif
