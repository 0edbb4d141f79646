// cofusion_binding.cpp -- the reference-side binding of INTEGRATION.md section 1 as a compiled C++ translation
// unit (TEST INFRASTRUCTURE): a stand-in for Core/CoFusion.{h,cpp} whose processFrame / spawnObjectModel /
// savePly / exportPoses forward to libcofusion_b200.so through include/cofusion_b200.h.  Eigen / OpenCV / Pangolin
// are not in this image, so FrameData and Matrix4f are minimal stand-ins with the reference's member names
// (Core/FrameData.h:25-50); everything else is the code a maintainer would paste into Core/CoFusion.cpp.
//
//   g++ -std=c++14 -I include tests/binding/cofusion_binding.cpp -L cofusion_b200 -lcofusion_b200 -o binding_test
//   ./binding_test            exit 0: ran 4 frames on cuda:0 (or: verified the loud failure without a CUDA device)
#include <cofusion_b200.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace ref_standin {
struct Mat {  // cv::Mat: what the binding touches
  std::vector<unsigned char> buf;
  unsigned char* data = nullptr;
  size_t n = 0;
  size_t total() const { return n; }
};
struct FrameData {  // Core/FrameData.h:25-50
  int64_t timestamp = 0;
  Mat mask, rgb, depth;
};
struct Matrix4f {  // Eigen::Matrix4f, row-major here
  float m[16];
};
}  // namespace ref_standin
using namespace ref_standin;

class CoFusion {  // Core/CoFusion.h:47-128, the members the path needs
 public:
  CoFusion(int w, int h, float fx, float fy, float cx, float cy, int timeDelta, float depthCut, float icpThresh, bool fastOdom,
           bool so3, bool frameToFrameRGB, float initConfidenceGlobal, float initConfidenceObject, unsigned maxVertices,
           bool enableMultipleModels_)
      : enableMultipleModels(enableMultipleModels_) {
    cfb_cofusion_params p;
    cfb_cofusion_default_params(&p);
    p.timeDelta = timeDelta;
    p.depthCutoff = depthCut;
    p.icpWeight = icpThresh;
    p.fastOdom = fastOdom;
    p.so3 = so3;
    p.frameToFrameRGB = frameToFrameRGB;
    p.confGlobalInit = initConfidenceGlobal;
    p.confObjectInit = initConfidenceObject;
    p.maxSurfels = maxVertices;
    p.enableMultipleModels = enableMultipleModels ? 1 : 0;
    if (cfb_cofusion_create(/*device*/ 0, w, h, fx, fy, cx, cy, &p, &cfb) != 0) throw std::runtime_error(cfb_last_error());
  }
  ~CoFusion() { cfb_cofusion_destroy(cfb); }

  // bool CoFusion::processFrame(const FrameData&, const Eigen::Matrix4f* inPose, const float weightMultiplier, const bool bootstrap)
  bool processFrame(const FrameData& frame, const Matrix4f* inPose = nullptr, const float weightMultiplier = 1.f,
                    const bool bootstrap = false) {
    cfb_frame f;
    std::memset(&f, 0, sizeof(f));
    f.rgb = frame.rgb.data;
    f.depth = reinterpret_cast<const float*>(frame.depth.data);
    f.depth_scale = 0.001f;
    f.mask = (!enableMultipleModels && frame.mask.total()) ? frame.mask.data : nullptr;
    f.timestamp = frame.timestamp;
    if (cfb_cofusion_process_frame_ex(cfb, &f, inPose ? inPose->m : nullptr, weightMultiplier, bootstrap ? 1 : 0) != 0)
      throw std::runtime_error(cfb_last_error());
    return false;
  }
  Matrix4f getPose() {  // globalModel->getPose()
    Matrix4f T;
    if (cfb_model_get_pose(cfb_cofusion_model(cfb, 0), T.m) != 0) throw std::runtime_error(cfb_last_error());
    return T;
  }
  unsigned lastCount() {
    unsigned n = 0;
    if (cfb_model_last_count(cfb_cofusion_model(cfb, 0), &n) != 0) throw std::runtime_error(cfb_last_error());
    return n;
  }
  void savePly(const std::string& exportDir) {
    if (cfb_cofusion_save_ply(cfb, exportDir.c_str()) != 0) throw std::runtime_error(cfb_last_error());
  }
  void exportPoses(const std::string& exportDir) {
    if (cfb_cofusion_export_poses(cfb, exportDir.c_str()) != 0) throw std::runtime_error(cfb_last_error());
  }
  void enablePoseLogging() { cfb_cofusion_enable_pose_logging(cfb, 1); }

 private:
  cfb_cofusion* cfb = nullptr;
  bool enableMultipleModels;
};

// a textured fronto-parallel wall at 2 m seen by a camera that slides 2 mm per frame
static void synth_frame(int W, int H, float fx, float fy, float cx, float cy, int t, FrameData& fr) {
  fr.rgb.buf.resize((size_t)W * H * 3);
  fr.depth.buf.resize((size_t)W * H * 4);
  fr.rgb.data = fr.rgb.buf.data();
  fr.depth.data = fr.depth.buf.data();
  fr.rgb.n = fr.depth.n = (size_t)W * H;
  fr.timestamp = 33 * t;
  float* d = reinterpret_cast<float*>(fr.depth.data);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const float z = 2.0f + 0.3f * std::sin(0.01f * x) * std::cos(0.013f * y);
      const float X = (x - cx) * z / fx + 0.002f * t, Y = (y - cy) * z / fy;
      d[y * W + x] = z;
      const int check = ((int)std::floor(X * 8.f) + (int)std::floor(Y * 8.f)) & 1;
      float g = 128.f + 60.f * std::sin(8.f * X) * std::cos(6.f * Y) + (check ? 40.f : -40.f);
      g = g < 1.f ? 1.f : g > 255.f ? 255.f : g;
      unsigned char* px = fr.rgb.data + 3 * (y * W + x);
      px[0] = px[1] = px[2] = (unsigned char)g;
    }
}

int main(int argc, char** argv) {
  const int W = 320, H = 240;
  const float fx = 264.f, fy = 264.f, cx = 160.f, cy = 120.f;
  if (cfb_device_count() <= 0) {
    // no CUDA device: the module must fail loudly (there is no CPU fallback)
    try {
      CoFusion f(W, H, fx, fy, cx, cy, 200, 5.f, 10.f, false, true, false, 10.f, 0.01f, 1u << 18, false);
    } catch (const std::exception& e) {
      std::printf("binding: no CUDA device -> %s\n", e.what());
      return std::strstr(e.what(), "CUDA device") ? 0 : 2;
    }
    return 3;
  }
  CoFusion f(W, H, fx, fy, cx, cy, 200, 5.f, 10.f, false, true, false, 10.f, 0.01f, 1u << 18, false);
  f.enablePoseLogging();
  FrameData fr;
  for (int t = 0; t < 4; ++t) {
    synth_frame(W, H, fx, fy, cx, cy, t, fr);
    f.processFrame(fr);
  }
  const Matrix4f T = f.getPose();
  const unsigned n = f.lastCount();
  std::printf("binding: 4 frames, %u surfels, t = (%.5f %.5f %.5f)\n", n, T.m[3], T.m[7], T.m[11]);
  bool ok = n > (unsigned)(W * H) / 2 && std::isfinite(T.m[3]) && std::fabs(T.m[3]) < 0.05f && std::fabs(T.m[15] - 1.f) < 1e-6f;
  // pose supplied by the caller: no tracking, the pose is taken as is (CoFusion.cpp:343-345)
  Matrix4f P = T;
  P.m[3] += 0.001f;
  synth_frame(W, H, fx, fy, cx, cy, 4, fr);
  f.processFrame(fr, &P);
  const Matrix4f T2 = f.getPose();
  ok = ok && std::memcmp(T2.m, P.m, sizeof(P.m)) == 0;
  if (argc > 1) {
    f.savePly(argv[1]);
    f.exportPoses(argv[1]);
  }
  std::printf("binding: %s\n", ok ? "ok" : "FAILED");
  return ok ? 0 : 1;
}
