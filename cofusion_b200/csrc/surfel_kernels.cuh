// surfel_kernels.cuh -- the surfel-map stage (reference: OpenGL shaders under Core/Shaders driven by
// Core/Model/Model.cpp and ModelProjection.cpp) as plain CUDA over raw device memory.
#pragma once
#include <stdint.h>

#include "cfb_common.cuh"

namespace cfb {

struct Surfel {  // Core/Shaders/Vertex.cpp:21-43 -- 48 B AoS
  float4 pos;    // xyz, confidence
  float4 col;    // colour (24-bit int as float), unused, init time, last time
  float4 nrm;    // normal, radius
};
static_assert(sizeof(Surfel) == 48, "surfel layout");

struct SurfelGeom {
  int W, H;
  float fx, fy, cx, cy;
};

struct Pose34 {  // first 3 rows of a row-major 4x4
  float m[12];
};

// Per-model pose block kept on the device: written by the tracker kernel's epilogue (or uploaded by the
// host after a pose override), read by every surfel-stage kernel -- no host round trip inside a frame.
struct PoseDev {
  Pose34 pose;       // model pose (camera -> model frame)
  Pose34 inv;        // its inverse
  Pose34 last;       // the pose before the last tracking step
  float tr[12];      // pose in the tracker's layout: t[3], R[9] row-major
  float weightBase;  // Model::computeFusionWeight with multiplier 1 (pose vs last)
  float pad[3];
};
// A pose argument: by value (host-driven seams, tests) or read from a PoseDev at kernel run time.
struct PoseRef {
  Pose34 v;
  const Pose34* dev;
  PoseRef(const Pose34& p) : v(p), dev(nullptr) {}
  explicit PoseRef(const Pose34* d) : v(), dev(d) {}
};
struct WeightRef {  // fusion weight by value, or weightBase of a PoseDev times a multiplier
  float v;
  const float* dev;
  float mult;
  WeightRef(float w) : v(w), dev(nullptr), mult(1.f) {}
  WeightRef(const float* d, float m) : v(0.f), dev(d), mult(m) {}
};

// Device-side counters of a map (kept on the device so no stage needs a host round trip).
struct MapCounters {
  unsigned count;          // live surfels in the current source buffer
  unsigned unstableCount;  // candidates produced by the last fuse
  unsigned scanTotal;      // scratch: result of the last flag scan
  unsigned fillInRequired; // CoFusion::requiresFillIn of the last prediction
  unsigned cleanTick;      // `time` of the clean() that produced `count` (lets the host bound the count without a sync)
  unsigned cleanTicket;    // scratch: blocks of the last clean_scatter that have finished (the last one closes the pass)
  unsigned fillSamples, fillTicket;  // scratch of the requiresFillIn count inside fill_in_kernel
};

struct IndexMaps {  // ModelProjection sparse targets (ModelProjection.cpp:72-76)
  uint32_t* index;
  float4 *vertConf, *colorTime, *normRad;
};
struct SplatMaps {  // combinedPredict targets (ModelProjection.cpp:90-94)
  uchar4* image;
  float4 *vertexConf, *normalRad;
  uint16_t* time;
};
struct FillMaps {  // FillIn targets (FillIn.cpp:21-23)
  uchar4* image;
  float4 *vertex, *normal;
};

struct ScanHostState {  // host-side bookkeeping of the single-pass scan (no reset launch between scans)
  unsigned epoch = 0;       // bumped per scan: tile status words of earlier scans never match
  unsigned ticketBase = 0;  // tickets handed out by all earlier scans
};
struct ScanScratch {
  uint8_t* flags;       // one byte per item
  uint32_t* ranks;      // exclusive prefix of flags
  uint32_t* blockSums;  // tile status words of the single-pass scan (u64 per 2048-item tile) + the ticket counter
  size_t capacity;      // items
  ScanHostState* host;  // owned by the Model
};

// a18: Model::initialise (Model.cpp:227-272) from the first frame
cudaError_t launch_surfel_initialise(const SurfelGeom& g, const uint8_t* rgb, const float* depthRaw,
                                     const float* depthFiltered, int time, float maxDepth, Surfel* dst,
                                     unsigned capacity, Surfel* stagingRaw, Surfel* stagingFil, ScanScratch sc,
                                     MapCounters* counters, cudaStream_t s);
// a13: ModelProjection::predictIndices
cudaError_t launch_predict_indices(const SurfelGeom& g, const Surfel* surfels, unsigned count_ub,
                                   const MapCounters* counters, const PoseRef& t_inv, int time, float maxDepth,
                                   int timeDelta, unsigned long long* keys, IndexMaps out, cudaStream_t s);
// a16: Model::fuse (data association + update), in place on `surfels`
cudaError_t launch_fuse(const SurfelGeom& g, Surfel* surfels, unsigned count_ub, MapCounters* counters,
                        const PoseRef& pose, int time, const uint8_t* rgb, const uint8_t* mask, const float* depthRaw,
                        const float* depthFiltered, float maxDepth, const WeightRef& weighting, unsigned maskID, IndexMaps idx,
                        uint32_t* winner, Surfel* candStaging, uint32_t* candBest, Surfel* unstable, ScanScratch sc,
                        cudaStream_t s);
// a17: Model::clean (stable compaction of old surfels then candidates into dst)
cudaError_t launch_clean(const SurfelGeom& g, Surfel* src, Surfel* unstable, Surfel* dst, unsigned count_ub,
                         unsigned cand_ub, unsigned capacity, MapCounters* counters, const PoseRef& t_inv, int time,
                         float confThreshold, int timeDelta, const float* depthFiltered, const uint8_t* mask,
                         unsigned maskID, float outlierCoeff, IndexMaps idx, ScanScratch sc, cudaStream_t s);
// a14: ModelProjection::combinedPredict
cudaError_t launch_combined_predict(const SurfelGeom& g, const Surfel* surfels, unsigned count_ub,
                                    MapCounters* counters, const PoseRef& t_inv, float maxDepth, float confThreshold,
                                    int time, int maxTime, int timeDelta, unsigned long long* keys, SplatMaps out,
                                    cudaStream_t s);
// a15: Model::performFillIn + CoFusion::requiresFillIn
cudaError_t launch_fill_in(const SurfelGeom& g, SplatMaps splat, const uint8_t* rgb, const float* depthFiltered,
                           int passthrough_geom, int passthrough_rgb, FillMaps out, MapCounters* counters,
                           float ratio, cudaStream_t s);

}  // namespace cfb
