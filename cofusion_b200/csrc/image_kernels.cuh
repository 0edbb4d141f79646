// image_kernels.cuh -- launch wrappers of the image preparation kernels (see image_kernels.cu).
// All pointers are device pointers, pitches are in bytes, `s` is the stream the launch goes to.
#pragma once
#include "cfb_common.cuh"

namespace cfb {

// frame ingest: raw u16 depth -> metric f32 (x scale), optional first/third channel swap of the 8-bit image
// (either input may be null: that output is left untouched)
cudaError_t launch_ingest(const uint8_t* img, const uint16_t* depth_u16, float scale, int flip, uint8_t* rgb, float* depth,
                          int n, cudaStream_t s);
// both pyramid levels (sw/2 x sh/2, sw/4 x sh/4) of up to four unpitched images in ONE launch; is_u8[k]: 0 = f32 with NaN
// as invalid (pyrDownGaussF), 1 = u8 with 0 as invalid (pyrDownUcharGauss).  Bit-identical to two launch_pyr_down_* calls.
cudaError_t launch_pyramid2(int njobs, const void* const* src, void* const* l1, void* const* l2, const int* is_u8, int sw, int sh,
                            cudaStream_t s);
cudaError_t launch_bilateral(const float* depth, size_t dpitch, int W, int H, float maxD, float* out,
                             size_t opitch, cudaStream_t s);
cudaError_t launch_pyr_down_gauss_f(const float* src, size_t spitch, int sw, int sh, float* dst,
                                    size_t dpitch, cudaStream_t s);
cudaError_t launch_pyr_down_uchar(const unsigned char* src, size_t spitch, int sw, int sh,
                                  unsigned char* dst, size_t dpitch, cudaStream_t s);
cudaError_t launch_create_vmap(const float* depth, size_t dpitch, int W, int H, Intr k, float cutoff,
                               float* vmap, size_t vpitch, cudaStream_t s);
cudaError_t launch_create_nmap(const float* vmap, size_t vpitch, int W, int H, float* nmap, size_t npitch,
                               cudaStream_t s);
cudaError_t launch_copy_maps(const float* v4, const float* n4, int W, int H, float* vmap, size_t vpitch,
                             float* nmap, size_t npitch, cudaStream_t s);
cudaError_t launch_resize_map(const float* in, size_t ipitch, int sw, int sh, bool normalize, float* out,
                              size_t opitch, cudaStream_t s);
cudaError_t launch_transform_maps(const float* vsrc, size_t vspitch, const float* nsrc, size_t nspitch, int W,
                                  int H, const Mat33& R, const float t[3], float* vdst, size_t vdpitch,
                                  float* ndst, size_t ndpitch, cudaStream_t s);
cudaError_t launch_vertices_to_depth(const float* v4, int W, int H, float cutoff, float* dst, size_t dpitch,
                                     cudaStream_t s);
cudaError_t launch_rgb_to_intensity(const unsigned char* rgb, size_t pitch, int channels, int W, int H,
                                    unsigned char* dst, size_t dpitch, cudaStream_t s);
cudaError_t launch_derivative_images(const unsigned char* src, size_t spitch, int W, int H, short* dx, short* dy,
                                     size_t gpitch, cudaStream_t s);
cudaError_t launch_project_to_point_cloud(const float* depth, size_t dpitch, int W, int H, Intr k, float* cloud,
                                          size_t cpitch, cudaStream_t s);

// ---- fused builders for Model::performTracking (same arithmetic, fewer launches; unpitched buffers)
// Model::initICP's choice between the splat prediction and the fill-in images (Model.cpp:350-367,
// CoFusion::requiresFillIn), taken where the maps are READ: *sel != 0 -> the alternative vertex / normal maps;
// *sel != 0 or img_always -> the alternative image.  All null: no choice to make.
struct PredAlt {
  const float *v4 = nullptr, *n4 = nullptr;
  const unsigned char* img = nullptr;
  const unsigned* sel = nullptr;
  int img_always = 0;
};
cudaError_t launch_model_pyramid(const float* v4, const float* n4, int W, int H, const Mat33& R, const float t[3],
                                 float cutoffRGB, float* const v[3], float* const n[3], float* depth0,
                                 cudaStream_t s, const float* pose34_dev = nullptr, const PredAlt* alt = nullptr);
cudaError_t launch_frame_maps(const float* const depth[3], int W, int H, Intr K, float cutoff, float* const v[3],
                              float* const n[3], cudaStream_t s, const unsigned char* imgA = nullptr, int chA = 0,
                              unsigned char* greyA = nullptr, const unsigned char* imgB = nullptr, int chB = 0,
                              unsigned char* greyB = nullptr, const PredAlt* altA = nullptr);
cudaError_t launch_intensity2(const unsigned char* a, int cha, unsigned char* da, const unsigned char* b, int chb,
                              unsigned char* db, int n, cudaStream_t s);
cudaError_t launch_pyr_down_uchar2(const unsigned char* sa, unsigned char* da, const unsigned char* sb,
                                   unsigned char* db, int sw, int sh, cudaStream_t s);

}  // namespace cfb
