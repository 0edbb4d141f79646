/*
 * ref_driver.cu -- thin extern "C" driver around the REFERENCE's own CUDA sources
 * (/root/reference/Core/Cuda/{reduce.cu,cudafuncs.cu,containers/device_memory.cpp}), compiled
 * where they lie by oracle/build_ref.sh into oracle/_ref/libcfref.so.
 *
 * TEST INFRASTRUCTURE ONLY (see cf_oracle.h).  It exists so that (a) the CPU restatement in
 * oracle/tracker.c and the product kernels can both be checked against the reference kernels
 * themselves on a B200, and (b) "the reference CUDA tracker on the same box" can be timed through
 * the reference's own call sequence (RGBDOdometry.cpp:217-477), unknown-GPU launch defaults
 * (GPUConfig.h:50-58).  Nothing here is reference source: it only calls the functions declared in
 * Core/Cuda/cudafuncs.cuh:64-193 with host buffers uploaded into the reference's DeviceArray2D.
 */
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <vector>

#include "cudafuncs.cuh"  // reference header (include path set by build_ref.sh)
#include "cf_oracle.h"

namespace {
template <class T>
void up(DeviceArray2D<T>& d, const void* h, int rows, int cols) {
  d.create(rows, cols);
  d.upload(h, (size_t)cols * sizeof(T), rows, cols);
}
template <class T>
void down(const DeviceArray2D<T>& d, void* h) {
  d.download(h, (size_t)d.cols() * sizeof(T));
}
mat33 m33(const float* r) {
  mat33 m;
  memcpy(m.data, r, sizeof(float) * 9);
  return m;
}
struct Launch {
  int icpT = 128, icpB = 112, rgbT = 128, rgbB = 112, resT = 256, resB = 336, so3T = 160, so3B = 64;
};  // GPUConfig.h:50-58 (defaults for an unknown GPU)
const Launch kL;
}  // namespace

extern "C" {

int ref_device_ok() {
  int n = 0;
  return cudaGetDeviceCount(&n) == cudaSuccess && n > 0;
}

void ref_create_vmap(const float* depth, int W, int H, float fx, float fy, float cx, float cy,
                     float cutoff, float* vmap) {
  DeviceArray2D<float> d, v;
  DeviceArray2D<unsigned char> m;
  std::vector<unsigned char> zero((size_t)W * H, 0);
  up(d, depth, H, W);
  up(m, zero.data(), H, W);
  // pre-fill the output with NaN so the y/z planes of invalid pixels are defined
  std::vector<float> nanbuf((size_t)3 * W * H);
  unsigned q = 0x7fffffffu;
  for (auto& f : nanbuf) memcpy(&f, &q, 4);
  up(v, nanbuf.data(), 3 * H, W);
  createVMap(CameraModel(fx, fy, cx, cy), d, m, v, cutoff, 0);
  cudaDeviceSynchronize();
  down(v, vmap);
}

void ref_create_nmap(const float* vmap, int W, int H, float* nmap) {
  DeviceArray2D<float> v, n;
  up(v, vmap, 3 * H, W);
  std::vector<float> nanbuf((size_t)3 * W * H);
  unsigned q = 0x7fffffffu;
  for (auto& f : nanbuf) memcpy(&f, &q, 4);
  up(n, nanbuf.data(), 3 * H, W);
  createNMap(v, n);
  cudaDeviceSynchronize();
  down(n, nmap);
}

void ref_copy_maps(const float* v4, const float* n4, int W, int H, float* vmap, float* nmap) {
  DeviceArray<float> vs, ns;
  vs.upload(v4, (size_t)W * H * 4);
  ns.upload(n4, (size_t)W * H * 4);
  DeviceArray2D<float> vd(3 * H, W), nd(3 * H, W);
  copyMaps(vs, ns, vd, nd);
  cudaDeviceSynchronize();
  down(vd, vmap);
  down(nd, nmap);
}

void ref_resize_map(const float* in, int sw, int sh, int normalize, float* out) {
  DeviceArray2D<float> i, o;
  up(i, in, 3 * sh, sw);
  std::vector<float> nanbuf((size_t)3 * (sw / 2) * (sh / 2));
  unsigned q = 0x7fffffffu;
  for (auto& f : nanbuf) memcpy(&f, &q, 4);
  up(o, nanbuf.data(), 3 * (sh / 2), sw / 2);
  if (normalize)
    resizeNMap(i, o);
  else
    resizeVMap(i, o);
  down(o, out);
}

void ref_transform_maps(const float* vsrc, const float* nsrc, int W, int H, const float* R,
                        const float* t, float* vdst, float* ndst) {
  DeviceArray2D<float> v, n;
  up(v, vsrc, 3 * H, W);
  up(n, nsrc, 3 * H, W);
  std::vector<float> nanbuf((size_t)3 * W * H);
  unsigned q = 0x7fffffffu;
  for (auto& f : nanbuf) memcpy(&f, &q, 4);
  DeviceArray2D<float> vd, nd;
  up(vd, nanbuf.data(), 3 * H, W);
  up(nd, nanbuf.data(), 3 * H, W);
  float3 tv = {t[0], t[1], t[2]};
  tranformMaps(v, n, m33(R), tv, vd, nd);
  cudaDeviceSynchronize();
  down(vd, vdst);
  down(nd, ndst);
}

void ref_pyr_down_gauss_f(const float* src, int sw, int sh, float* dst) {
  DeviceArray2D<float> s, d;
  up(s, src, sh, sw);
  pyrDownGaussF(s, d);
  cudaDeviceSynchronize();
  down(d, dst);
}

void ref_pyr_down_uchar_gauss(const unsigned char* src, int sw, int sh, unsigned char* dst) {
  DeviceArray2D<unsigned char> s, d;
  up(s, src, sh, sw);
  pyrDownUcharGauss(s, d);
  cudaDeviceSynchronize();
  down(d, dst);
}

void ref_vertices_to_depth(const float* v4, int W, int H, float cutoff, float* depth) {
  DeviceArray<float> vs;
  vs.upload(v4, (size_t)W * H * 4);
  DeviceArray2D<float> d(H, W);
  verticesToDepth(vs, d, cutoff);
  cudaDeviceSynchronize();
  down(d, depth);
}

void ref_derivative_images(const unsigned char* img, int W, int H, short* dx, short* dy) {
  DeviceArray2D<unsigned char> s;
  up(s, img, H, W);
  DeviceArray2D<short> gx(H, W), gy(H, W);
  computeDerivativeImages(s, gx, gy);
  down(gx, dx);
  down(gy, dy);
}

void ref_project_to_point_cloud(const float* depth, int W, int H, float fx, float fy, float cx,
                                float cy, float* cloud3) {
  DeviceArray2D<float> d;
  up(d, depth, H, W);
  DeviceArray2D<float3> c(H, W);
  CameraModel intr(fx, fy, cx, cy);
  projectToPointCloud(d, c, intr, 0);
  down(c, cloud3);
}

void ref_icp_step(const float* Rcurr, const float* tcurr, const float* vmap_curr,
                  const float* nmap_curr, const float* Rprev_inv, const float* tprev, float fx,
                  float fy, float cx, float cy, const float* vmap_g_prev, const float* nmap_g_prev,
                  float distThres, float angleThres, int W, int H, float* A, float* b,
                  float* residual) {
  DeviceArray2D<float> vc, nc, vp, np;
  up(vc, vmap_curr, 3 * H, W);
  up(nc, nmap_curr, 3 * H, W);
  up(vp, vmap_g_prev, 3 * H, W);
  up(np, nmap_g_prev, 3 * H, W);
  DeviceArray<JtJJtrSE3> sum, out;
  sum.create(MAX_THREADS);
  out.create(1);
  float3 tc = {tcurr[0], tcurr[1], tcurr[2]}, tp = {tprev[0], tprev[1], tprev[2]};
  icpStep(m33(Rcurr), tc, vc, nc, m33(Rprev_inv), tp, CameraModel(fx, fy, cx, cy), vp, np,
          distThres, angleThres, sum, out, A, b, residual, kL.icpT, kL.icpB, 0);
}

void ref_rgb_residual(float minScale, const short* dIdx, const short* dIdy, const float* lastDepth,
                      const float* nextDepth, const unsigned char* lastImage,
                      const unsigned char* nextImage, void* corres, float maxDepthDelta,
                      const float* kt, const float* krkinv, int W, int H, int* sigmaSum,
                      int* count) {
  DeviceArray2D<short> gx, gy;
  DeviceArray2D<float> ld, nd;
  DeviceArray2D<unsigned char> li, ni, lm, nm;
  up(gx, dIdx, H, W);
  up(gy, dIdy, H, W);
  up(ld, lastDepth, H, W);
  up(nd, nextDepth, H, W);
  up(li, lastImage, H, W);
  up(ni, nextImage, H, W);
  lm.create(H, W);
  nm.create(H, W);
  // corresImg is indexed as data[k] with k = y*cols+x (reduce.cu:862): it must be unpitched.
  DataTerm* cbuf;
  cudaMalloc(&cbuf, sizeof(DataTerm) * W * H);
  cudaMemset(cbuf, 0, sizeof(DataTerm) * W * H);
  DeviceArray2D<DataTerm> ci(H, W, cbuf, sizeof(DataTerm) * W);
  DeviceArray<int2> sumRes;
  sumRes.create(MAX_THREADS);
  float3 ktv = {kt[0], kt[1], kt[2]};
  computeRgbResidual(minScale, gx, gy, ld, nd, li, ni, lm, nm, ci, sumRes, maxDepthDelta, ktv,
                     m33(krkinv), *sigmaSum, *count, kL.resT, kL.resB, 0, 0);
  cudaMemcpy(corres, cbuf, sizeof(DataTerm) * W * H, cudaMemcpyDeviceToHost);
  cudaFree(cbuf);
}

void ref_rgb_step(const void* corres, float sigma, const float* cloud3, float fx, float fy,
                  const short* dIdx, const short* dIdy, float sobelScale, int W, int H, float* A,
                  float* b) {
  DataTerm* cbuf;
  cudaMalloc(&cbuf, sizeof(DataTerm) * W * H);
  cudaMemcpy(cbuf, corres, sizeof(DataTerm) * W * H, cudaMemcpyHostToDevice);
  DeviceArray2D<DataTerm> ci(H, W, cbuf, sizeof(DataTerm) * W);
  DeviceArray2D<float3> cl;
  up(cl, cloud3, H, W);
  DeviceArray2D<short> gx, gy;
  up(gx, dIdx, H, W);
  up(gy, dIdy, H, W);
  DeviceArray<JtJJtrSE3> sum, out;
  sum.create(MAX_THREADS);
  out.create(1);
  rgbStep(ci, sigma, cl, fx, fy, gx, gy, sobelScale, sum, out, A, b, kL.rgbT, kL.rgbB);
  cudaFree(cbuf);
}

void ref_so3_step(const unsigned char* lastImage, const unsigned char* nextImage,
                  const float* imageBasis, const float* kinv, const float* krlr, int W, int H,
                  float* A, float* b, float* residual) {
  DeviceArray2D<unsigned char> li, ni;
  up(li, lastImage, H, W);
  up(ni, nextImage, H, W);
  DeviceArray<JtJJtrSO3> sum, out;
  sum.create(MAX_THREADS);
  out.create(1);
  so3Step(li, ni, m33(imageBasis), m33(kinv), m33(krlr), sum, out, A, b, residual, kL.so3T,
          kL.so3B);
}

/* ---------------------------------------------------------------- full tracker through the
 * reference kernels: step backend for orc_odom_track_ex (host GN loop = oracle/tracker.c). */
struct RefBackend {
  DeviceArray2D<float> vc[3], nc[3], vp[3], np[3], ld[3], nd[3];
  DeviceArray2D<unsigned char> li[3], ni[3], lni[3], lm[3], nm[3];
  DeviceArray2D<short> gx[3], gy[3];
  DeviceArray2D<float3> cl[3];
  DataTerm* cbuf[3] = {0, 0, 0};
  DeviceArray2D<DataTerm> ci[3];
  DeviceArray<JtJJtrSE3> sumSE3, outSE3;
  DeviceArray<JtJJtrSO3> sumSO3, outSO3;
  DeviceArray<int2> sumRes;
  double step_ms = 0;  // wall time spent inside the reference *Step calls (they sync internally)
  int steps = 0;
  int W = 0, H = 0;
  float intr[4];
  float distThres, angleThres, sobelScale, maxDepthDelta;
};

static void rb_begin(void* u, OrcOdometry* o) {
  RefBackend* r = (RefBackend*)u;
  orc_odom_dims(o, &r->W, &r->H, r->intr);
  for (int i = 0; i < 3; ++i) {
    int w = r->W >> i, h = r->H >> i;
    up(r->vc[i], orc_odom_view(o, 0, i), 3 * h, w);
    up(r->nc[i], orc_odom_view(o, 1, i), 3 * h, w);
    up(r->vp[i], orc_odom_view(o, 2, i), 3 * h, w);
    up(r->np[i], orc_odom_view(o, 3, i), 3 * h, w);
    up(r->ld[i], orc_odom_view(o, 4, i), h, w);
    up(r->nd[i], orc_odom_view(o, 5, i), h, w);
    up(r->li[i], orc_odom_view(o, 6, i), h, w);
    up(r->ni[i], orc_odom_view(o, 7, i), h, w);
    up(r->gx[i], orc_odom_view(o, 8, i), h, w);
    up(r->gy[i], orc_odom_view(o, 9, i), h, w);
    up(r->lni[i], orc_odom_view(o, 10, i), h, w);
    up(r->cl[i], orc_odom_view(o, 11, i), h, w);
    r->lm[i].create(h, w);
    r->nm[i].create(h, w);
    if (!r->cbuf[i]) cudaMalloc(&r->cbuf[i], sizeof(DataTerm) * w * h);
    r->ci[i] = DeviceArray2D<DataTerm>(h, w, r->cbuf[i], sizeof(DataTerm) * w);
  }
  if (r->sumSE3.size() == 0) {
    r->sumSE3.create(MAX_THREADS);
    r->outSE3.create(1);
    r->sumSO3.create(MAX_THREADS);
    r->outSO3.create(1);
    r->sumRes.create(MAX_THREADS);
  }
  cudaDeviceSynchronize();
}

struct Tic {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  double ms() const {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
};

static void rb_so3(void* u, OrcOdometry*, int L, const float* ib, const float* kinv,
                   const float* krlr, float* A, float* b, float* res) {
  RefBackend* r = (RefBackend*)u;
  Tic t;
  so3Step(r->lni[L], r->ni[L], m33(ib), m33(kinv), m33(krlr), r->sumSO3, r->outSO3, A, b, res,
          kL.so3T, kL.so3B);
  r->step_ms += t.ms();
  r->steps++;
}
static void rb_res(void* u, OrcOdometry*, int i, float minScale, const float* kt,
                   const float* krkinv, int* sigma, int* count) {
  RefBackend* r = (RefBackend*)u;
  float3 ktv = {kt[0], kt[1], kt[2]};
  Tic t;
  computeRgbResidual(minScale, r->gx[i], r->gy[i], r->ld[i], r->nd[i], r->li[i], r->ni[i], r->lm[i],
                     r->nm[i], r->ci[i], r->sumRes, 0.07f, ktv, m33(krkinv), *sigma, *count,
                     kL.resT, kL.resB, 0, 0);
  r->step_ms += t.ms();
  r->steps++;
}
static void rb_icp(void* u, OrcOdometry*, int i, const float* Rc, const float* tc,
                   const float* Rpi, const float* tp, float* A, float* b, float* res,
                   float* error_map) {
  RefBackend* r = (RefBackend*)u;
  (void)error_map;  // the error surface needs a GL texture in the reference; not witnessed here
  int div = 1 << i;
  float3 tcv = {tc[0], tc[1], tc[2]}, tpv = {tp[0], tp[1], tp[2]};
  Tic t;
  icpStep(m33(Rc), tcv, r->vc[i], r->nc[i], m33(Rpi), tpv,
          CameraModel(r->intr[0] / div, r->intr[1] / div, r->intr[2] / div, r->intr[3] / div),
          r->vp[i], r->np[i], r->distThres, r->angleThres, r->sumSE3, r->outSE3, A, b, res, kL.icpT,
          kL.icpB, 0);
  r->step_ms += t.ms();
  r->steps++;
}
static void rb_rgb(void* u, OrcOdometry*, int i, float sigma, float* A, float* b) {
  RefBackend* r = (RefBackend*)u;
  int div = 1 << i;
  Tic t;
  rgbStep(r->ci[i], sigma, r->cl[i], r->intr[0] / div, r->intr[1] / div, r->gx[i], r->gy[i],
          0.125f, r->sumSE3, r->outSE3, A, b, kL.rgbT, kL.rgbB);
  r->step_ms += t.ms();
  r->steps++;
}

/* Runs oracle/tracker.c's GN loop with the reference kernels as the step backend.  step_ms
 * returns the wall time spent inside the reference *Step host wrappers (each ends with
 * cudaDeviceSynchronize + D2H, so wall time is what the reference app pays). */
void ref_odom_track(OrcOdometry* o, float* trans, float* rot, int rgbOnly, float icpWeight,
                    int pyramid, int fastOdom, int so3, float distThres, float angleThres,
                    OrcTrackStats* st, double* step_ms, int* steps) {
  static RefBackend* rb = nullptr;
  if (!rb) rb = new RefBackend();
  rb->distThres = distThres;
  rb->angleThres = angleThres;
  rb->step_ms = 0;
  rb->steps = 0;
  OrcStepBackend be = {rb, rb_begin, rb_so3, rb_res, rb_icp, rb_rgb, 0};
  orc_odom_track_ex(o, trans, rot, rgbOnly, icpWeight, pyramid, fastOdom, so3, nullptr, st, &be);
  if (step_ms) *step_ms = rb->step_ms;
  if (steps) *steps = rb->steps;
}

}  // extern "C"
