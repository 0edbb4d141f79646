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

/* ---------------------------------------------------------------- surfel map (oracle/surfel.c)
 * Restatement of the OpenGL predict / fuse / clean stage (Core/Model, Core/Shaders). PARITY UNPINNED
 * (no GL context, no golden vectors in the reference); frozen GL semantics F1-F6 in surfel.c. */
typedef struct { /* Core/Shaders/Vertex.cpp:21-43 -- 48 bytes */
  float pos[4]; /* xyz, confidence */
  float col[4]; /* colour as 24-bit int in a float, unused, init time, last-update time */
  float nrm[4]; /* normal xyz, radius */
} OrcSurfel;
typedef struct OrcSurfelMap OrcSurfelMap;

OrcSurfelMap* orc_map_create(int W, int H, float fx, float fy, float cx, float cy, unsigned max_surfels);
void orc_map_destroy(OrcSurfelMap* m);
unsigned orc_map_count(const OrcSurfelMap* m);
const OrcSurfel* orc_map_surfels(const OrcSurfelMap* m);
unsigned orc_map_unstable_count(const OrcSurfelMap* m);
const OrcSurfel* orc_map_unstable(const OrcSurfelMap* m);
void orc_map_set_surfels(OrcSurfelMap* m, const OrcSurfel* s, unsigned n);
/* which: 0 index(u32) 1 vertConf 2 colorTime 3 normRad (float4) 4 image(RGBA8) 5 splat vertexConf
 * 6 splat normalRad (float4) 7 splat time(u16) 8 fill image 9 fill vertex 10 fill normal */
const void* orc_map_view(const OrcSurfelMap* m, int which);
void orc_pose_inverse(const float T[16], float Ti[16]);
/* Model::initialise via vertex_feedback + init_unstable (Model.cpp:227-272, CoFusion.cpp:161-169) */
void orc_map_initialise(OrcSurfelMap* m, const uint8_t* rgb, const float* depthRaw, const float* depthFiltered,
                        int time, float maxDepth);
/* ModelProjection::predictIndices (ModelProjection.cpp:105-157) */
void orc_map_predict_indices(OrcSurfelMap* m, const float pose[16], int time, float maxDepth, int timeDelta);
/* Model::computeFusionWeight (Model.cpp:391-406) */
float orc_fusion_weight(const float pose[16], const float lastPose[16], float weightMultiplier);
/* Model::fuse (Model.cpp:408-563); maxDepth = min(depthCutoff, model maxDepth) */
void orc_map_fuse(OrcSurfelMap* m, const float pose[16], int time, const uint8_t* rgb, const uint8_t* mask,
                  const float* depthRaw, const float* depthFiltered, float maxDepth, float weighting,
                  unsigned maskID);
/* Model::clean (Model.cpp:565-697) */
void orc_map_clean(OrcSurfelMap* m, const float pose[16], int time, float confThreshold, int timeDelta,
                   const float* depthFiltered, const uint8_t* mask, unsigned maskID, float outlierCoeff);
/* ModelProjection::combinedPredict (ModelProjection.cpp:192-273) */
void orc_map_combined_predict(OrcSurfelMap* m, const float pose[16], float maxDepth, float confThreshold, int time,
                              int maxTime, int timeDelta);
/* Model::performFillIn (Model.cpp:901-909) */
void orc_map_fill_in(OrcSurfelMap* m, const uint8_t* rgb, const float* depthFiltered, int passthrough_geom,
                     int passthrough_rgb);
/* CoFusion::requiresFillIn (CoFusion.cpp:547-565) */
int orc_map_requires_fill_in(const OrcSurfelMap* m, float ratio);

/* ---------------------------------------------------------------- segmentation (oracle/segment.c)
 * Restatement of Segmentation::performSegmentationCRF + Slic + ConnectedLabels (Core/Segmentation).
 * PARITY UNPINNED: gSLICr and densecrf are un-vendored and unpinned; see the header of segment.c for
 * the published algorithms restated and the frozen choice (exact Gaussian kernels). */
typedef struct { /* Segmentation.h:123-142 with the GUI defaults of GUI/Tools/GUI.h:212-227 */
  int crfIterations;
  float scaleFeaturesRGB, scaleFeaturesDepth, scaleFeaturesPos;
  float weightAppearance, weightSmoothness;
  float unaryThresholdNew, unaryKError, unaryWeightError;
  float maxRelSizeNew, minRelSizeNew;
} OrcSegParams;
typedef struct { /* SegmentationResult::ModelData (Segmentation.h:41-67) */
  unsigned id;
  unsigned superPixelCount;
  float avgConfidence, depthMean, depthStd;
  unsigned short top, right, bottom, left;
} OrcModelData;
void orc_seg_default_params(OrcSegParams* p);
/* second witness (lattice.c): 1 = evaluate the CRF kernel products on a permutohedral lattice like densecrf,
 * 0 (default) = exactly.  Affects the following orc_segment_crf calls. */
void orc_segment_set_kernel_mode(int mode);
typedef struct OrcLattice OrcLattice;
OrcLattice* orc_lattice_create(const float* feat, int d, int N);
void orc_lattice_destroy(OrcLattice* L);
void orc_lattice_compute(const OrcLattice* L, const float* in, int vs, float* out);
void orc_lattice_norm(const OrcLattice* L, float* norm);
void orc_lattice_apply(const OrcLattice* L, const float* norm, const float* Q, int Lbl, float* out);
/* gSLICr restatement: labels[H*W] in [0, ceil(W/s)*ceil(H/s)) */
void orc_slic(const uint8_t* rgb, int W, int H, int spixel_size, int no_iters, float coh_weight, int* labels);
/* performSegmentationCRF.  icpError[m]: HxW f32 (Model::icpError), vertConf4[m]: HxW float4 splat
 * vertex+confidence (only .w is read, Segmentation.cpp:187).  md must hold numModels+1 entries.
 * Returns the number of valid md entries; fullSeg (HxW u8) receives model ids / 255.
 * Optional outputs (may be NULL): slicLabels (H*W int), unary (N*numLabels, node major), lowMap (N). */
int orc_segment_crf(const uint8_t* rgb, const float* depth, int W, int H, int numModels, const unsigned char* modelIds,
                    const float* const* icpError, const float* const* vertConf4, unsigned char nextModelID,
                    int allowNew, const OrcSegParams* prm, uint8_t* fullSeg, OrcModelData* md, int* hasNewLabel,
                    int* slicLabelsOut, float* unaryOut, uint8_t* lowMapOut);

#ifdef __cplusplus
}
#endif
#endif
