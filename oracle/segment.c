/*
 * segment.c -- CPU oracle (TEST INFRASTRUCTURE, see cf_oracle.h) for the motion segmentation
 * (Core/Segmentation/Segmentation.cpp:124-706, Slic.h:48-209, Slic.cpp:23-112,
 * ConnectedLabels.hpp:50-172).
 *
 * PARITY UNPINNED.  The two libraries that do the heavy lifting are NOT under /root/reference:
 *   - gSLICr (carlren/gSLICr, `git clone --depth=1` HEAD, Scripts/install.sh:85 -- unpinned):
 *     call sites Slic.cpp:33-46 (settings: spixel_size 16, coh_weight 0.6, 5 iterations, RGB colour
 *     space, GIVEN_SIZE, no connectivity enforcement) and :72-75.  Restated from the published
 *     algorithm (Ren, Prisacariu, Reid, "gSLICr: SLIC superpixels at over 250Hz", 2015): centres on a
 *     regular grid, 6 x [assign each pixel to the nearest of the 3x3 neighbouring centres under
 *     d = sqrt(dc^2 * Nc + w * dxy^2 * Nxy)], 5 x centre update.
 *   - densecrf (martinruenz/densecrf HEAD, Scripts/install.sh:84 -- unpinned): call sites
 *     Segmentation.cpp:221, :436-437, :452, :462-471.  Published algorithm: Kraehenbuehl & Koltun,
 *     "Efficient inference in fully connected CRFs with Gaussian edge potentials" (2011): mean-field
 *     updates Q <- softmax(-unary - sum_k compat_k * K_k Q) with Gaussian kernels K_k; the library
 *     evaluates K Q approximately on a permutohedral lattice.  FROZEN CHOICE: this restatement
 *     evaluates the kernel product EXACTLY, k(i,j) = exp(-|f_i - f_j|^2 / 2) including j = i, with the
 *     library's symmetric normalisation n_i = 1/sqrt(sum_j k(i,j) + 1e-20), out = n .* K (n .* Q);
 *     the float sums over j use the fixed order documented at butterfly32() below.
 *     With 1200 super-pixels that is a 1200 x 1200 kernel -- cheaper than building a lattice and
 *     free of its (order dependent) approximation error.  Labels are therefore bit-exact only
 *     against this restatement, not against a densecrf build (hard part 6).
 * Everything that IS in the tree is restated line by line, quirks included:
 *   - CRF appearance features read the FULL-RES rgb buffer with a LOW-RES index (Segmentation.cpp:445-447);
 *   - Slic::mapToHigh(index) divides by spixelY instead of spixelX (Slic.h:197);
 *   - downsampleThresholded divides empty super-pixels by the wrong count (Slic.h:117-122) and both
 *     downsample variants resolve empties in place, in index order (Slic.h:74-83);
 *   - the bounding boxes live in unsigned shorts and wrap (Segmentation.h:58-61, .cpp:541-546).
 * exp() = orc_expf (detmath.h).  Compile with -ffp-contract=off.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "cf_oracle.h"
#include "detmath.h"

/* ---------------------------------------------------------------- SLIC (gSLICr restatement) */
void orc_slic(const uint8_t* rgb, int W, int H, int spixel_size, int no_iters, float coh_weight, int* labels) {
  const int mx = (W + spixel_size - 1) / spixel_size, my = (H + spixel_size - 1) / spixel_size, ns = mx * my;
  float max_xy_dist = 1.0f / (1.4242f * (float)spixel_size);
  float max_color_dist = 5.0f / (1.7321f * 128);
  max_color_dist *= max_color_dist;
  max_xy_dist *= max_xy_dist;
  float* cx = (float*)malloc(sizeof(float) * ns * 5); /* x, y, c0, c1, c2 */
  for (int y = 0; y < my; ++y)
    for (int x = 0; x < mx; ++x) {
      int ix = x * spixel_size + spixel_size / 2, iy = y * spixel_size + spixel_size / 2;
      ix = ix >= W ? (x * spixel_size + W) / 2 : ix;
      iy = iy >= H ? (y * spixel_size + H) / 2 : iy;
      float* c = &cx[(y * mx + x) * 5];
      c[0] = (float)ix;
      c[1] = (float)iy;
      for (int k = 0; k < 3; ++k) c[2 + k] = (float)rgb[(iy * W + ix) * 3 + k];
    }
  double* acc = (double*)malloc(sizeof(double) * ns * 6);
  for (int it = 0; it <= no_iters; ++it) {
    if (it > 0) { /* Update_Cluster_Center: sums of small integers, exact in f32 in any order */
      memset(acc, 0, sizeof(double) * ns * 6);
      for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
          double* a = &acc[labels[y * W + x] * 6];
          a[0] += x;
          a[1] += y;
          for (int k = 0; k < 3; ++k) a[2 + k] += rgb[(y * W + x) * 3 + k];
          a[5] += 1;
        }
      for (int s = 0; s < ns; ++s)
        if (acc[s * 6 + 5] != 0) {
          float n = (float)acc[s * 6 + 5];
          for (int k = 0; k < 5; ++k) cx[s * 5 + k] = (float)acc[s * 6 + k] / n;
        }
    }
    /* Find_Center_Association */
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        const int ctr_x = x / spixel_size, ctr_y = y / spixel_size;
        int minidx = -1;
        float dist = 999999.9999f;
        const uint8_t* p = &rgb[(y * W + x) * 3];
        for (int i = -1; i <= 1; i++)
          for (int j = -1; j <= 1; j++) {
            int cxc = ctr_x + j, cyc = ctr_y + i;
            if (cxc >= 0 && cyc >= 0 && cxc < mx && cyc < my) {
              const float* c = &cx[(cyc * mx + cxc) * 5];
              float dcolor = ((float)p[0] - c[2]) * ((float)p[0] - c[2]) + ((float)p[1] - c[3]) * ((float)p[1] - c[3]) +
                             ((float)p[2] - c[4]) * ((float)p[2] - c[4]);
              float dxy = ((float)x - c[0]) * ((float)x - c[0]) + ((float)y - c[1]) * ((float)y - c[1]);
              float cdist = sqrtf(dcolor * max_color_dist + coh_weight * dxy * max_xy_dist);
              if (cdist < dist) {
                dist = cdist;
                minidx = cyc * mx + cxc;
              }
            }
          }
        if (minidx >= 0) labels[y * W + x] = minidx;
      }
  }
  free(cx);
  free(acc);
}

/* ---------------------------------------------------------------- Slic.h helpers */
typedef struct {
  int W, H, spixelSize, spixelX, spixelY, spixelNum;
  const int* labels;
  unsigned* counts;
} SlicCtx;

static int map_to_high_index(const SlicCtx* s, unsigned index, int* ox, int* oy) {
  /* Slic.h:193-197: mapToHigh(index % spixelX, index / spixelY) -- quirk kept */
  int x = (int)(index % (unsigned)s->spixelX), y = (int)(index / (unsigned)s->spixelY);
  *ox = (int)(x * s->spixelSize + s->spixelSize * 0.5);
  *oy = (int)(y * s->spixelSize + s->spixelSize * 0.5);
  return 0;
}
static int resample_empty_index(const SlicCtx* s, unsigned index) { /* Slic.h:199-209 */
  int cx, cy;
  map_to_high_index(s, index, &cx, &cy);
  if (cy >= s->H) cy = s->H - 1;
  if (cx >= s->W) cx = s->W - 1;
  return s->labels[cx + cy * s->W];
}
/* Slic.h:48-84 downsample<float>(image, channel) */
static void slic_downsample_f(const SlicCtx* s, const float* img, int channels, int channel, float* res) {
  memset(res, 0, sizeof(float) * s->spixelNum);
  for (int i = 0; i < s->W * s->H; ++i) res[s->labels[i]] += img[(size_t)i * channels + channel];
  for (int index = 0; index < s->spixelNum; index++) {
    int cnt = (int)s->counts[index], readIndex = index;
    if (cnt == 0) {
      readIndex = resample_empty_index(s, index);
      cnt = (int)s->counts[readIndex];
    }
    res[index] = res[readIndex] / (float)cnt;
  }
}
/* Slic.h:88-126 downsampleThresholded<float> */
static void slic_downsample_thresholded(const SlicCtx* s, const float* img, float minThreshold, float* res) {
  memset(res, 0, sizeof(float) * s->spixelNum);
  unsigned* dc = (unsigned*)calloc(s->spixelNum, sizeof(unsigned));
  for (int i = 0; i < s->W * s->H; ++i)
    if (img[i] > minThreshold) {
      res[s->labels[i]] += img[i];
      dc[s->labels[i]]++;
    }
  for (int index = 0; index < s->spixelNum; index++) {
    int cnt = (int)dc[index], readIndex = index;
    if (cnt == 0) {
      readIndex = resample_empty_index(s, index);
      cnt = (int)s->counts[readIndex]; /* quirk: total count of the other super-pixel */
    }
    res[index] = res[readIndex] / (float)cnt;
  }
  free(dc);
}

/* ---------------------------------------------------------------- ConnectedLabels.hpp:50-172 */
typedef struct {
  unsigned char label;
  int top, right, bottom, left, size;
} Comp;

static int find_root(const int* roots, int i) {
  while (i != roots[i]) i = roots[i];
  return i;
}
static int connected_labels(const uint8_t* in, int rows, int cols, int* comp, Comp* stats /* rows*cols */) {
  int* roots = (int*)malloc(sizeof(int) * rows * cols);
  int nroots = 0;
#define NEWC() (roots[nroots] = nroots, nroots++)
  comp[0] = NEWC();
  for (int c = 1; c < cols; c++) comp[c] = (in[c] == in[c - 1]) ? comp[c - 1] : NEWC();
  for (int r = 1; r < rows; r++) {
    const uint8_t *row = in + r * cols, *last = in + (r - 1) * cols;
    int *cr = comp + r * cols, *lr = comp + (r - 1) * cols;
    cr[0] = (row[0] == last[0]) ? lr[0] : NEWC();
    for (int c = 1; c < cols; c++) {
      if (row[c] == row[c - 1]) {
        int cLeft = cr[c - 1], cTop = lr[c];
        if (row[c] == last[c] && cLeft != cTop) {
          int r1 = find_root(roots, cTop), r2 = find_root(roots, cLeft);
          if (r1 < r2) {
            roots[r2] = r1;
            cr[c] = r1;
          } else {
            roots[r1] = r2;
            cr[c] = r2;
          }
        } else
          cr[c] = cLeft;
      } else if (row[c] == last[c])
        cr[c] = lr[c];
      else
        cr[c] = NEWC();
    }
  }
#undef NEWC
  int* mapping = (int*)malloc(sizeof(int) * nroots);
  int rootCnt = 0;
  for (int id = 0; id < nroots; id++) {
    int root = find_root(roots, id);
    if (root == id)
      mapping[root] = rootCnt++;
    else
      roots[id] = root;
  }
  for (int id = 0; id < nroots; id++) roots[id] = mapping[roots[id]];
  for (int i = 0; i < rootCnt; ++i) {
    stats[i].top = 2147483647;
    stats[i].left = 2147483647;
    stats[i].right = stats[i].bottom = stats[i].size = 0;
    stats[i].label = 0;
  }
  for (int y = 0; y < rows; y++)
    for (int x = 0; x < cols; x++) {
      int c = roots[comp[y * cols + x]];
      comp[y * cols + x] = c;
      Comp* d = &stats[c];
      d->size++;
      d->label = in[y * cols + x];
      if (y < d->top) d->top = y;
      if (y > d->bottom) d->bottom = y;
      if (x < d->left) d->left = x;
      if (x > d->right) d->right = x;
    }
  free(roots);
  free(mapping);
  return rootCnt;
}

/* ---------------------------------------------------------------- dense CRF (exact kernels) */
ORC_FMA_CLONES static void exp_and_normalize(float* out, const float* in, int L, int N) { /* DenseCRF::expAndNormalize */
  for (int i = 0; i < N; ++i) {
    float mx = in[i * L];
    for (int l = 1; l < L; ++l) mx = in[i * L + l] > mx ? in[i * L + l] : mx;
    float sum = 0;
    for (int l = 0; l < L; ++l) {
      float v = orc_expf(in[i * L + l] - mx);
      out[i * L + l] = v;
      sum += v;
    }
    for (int l = 0; l < L; ++l) out[i * L + l] = out[i * L + l] / sum;
  }
}
/* FROZEN SUMMATION ORDER of the kernel products (the lattice of the library has its own, unrelated
 * order): 128 partial sums, partial k owning j = k, k+128, ... in ascending j; each group of 32
 * partials is combined by an xor butterfly (distance 16, 8, 4, 2, 1) and the four group sums as
 * (s0 + s1) + (s2 + s3).  It is the order four 32-wide warps produce naturally. */
static float butterfly32(float* v) {
  for (int o = 16; o > 0; o >>= 1) {
    float nv[32];
    for (int k = 0; k < 32; ++k) nv[k] = v[k] + v[k ^ o];
    memcpy(v, nv, sizeof(nv));
  }
  return v[0];
}
static float combine128(float* part) {
  float s0 = butterfly32(part), s1 = butterfly32(part + 32), s2 = butterfly32(part + 64), s3 = butterfly32(part + 96);
  return (s0 + s1) + (s2 + s3);
}
/* K (N x N, row-major), norm (N): out = norm .* (K (norm .* Q)); Q, out are N x L (node major) */
static void kernel_apply(const float* K, const float* norm, const float* Q, int N, int L, float* out) {
  float* nq = (float*)malloc(sizeof(float) * N * L);
  for (int i = 0; i < N; ++i)
    for (int l = 0; l < L; ++l) nq[i * L + l] = norm[i] * Q[i * L + l];
  for (int i = 0; i < N; ++i)
    for (int l = 0; l < L; ++l) {
      float part[128];
      memset(part, 0, sizeof(part));
      for (int j = 0; j < N; ++j) part[j & 127] += K[(size_t)i * N + j] * nq[j * L + l];
      out[i * L + l] = norm[i] * combine128(part);
    }
  free(nq);
}
ORC_FMA_CLONES static void build_kernel(const float* feat, int D, int N, float* K, float* norm) {
  for (int i = 0; i < N; ++i) {
    float part[128];
    memset(part, 0, sizeof(part));
    for (int j = 0; j < N; ++j) {
      float d2 = 0;
      for (int k = 0; k < D; ++k) {
        float d = feat[i * D + k] - feat[j * D + k];
        d2 += d * d;
      }
      float v = orc_expf(-0.5f * d2);
      K[(size_t)i * N + j] = v;
      part[j & 127] += v;
    }
    norm[i] = 1.0f / sqrtf(combine128(part) + 1e-20f);
  }
}

/* kernel realisation of the next orc_segment_crf calls: 0 = exact products (the oracle of record),
 * 1 = permutohedral lattice (lattice.c, the densecrf realisation -- second witness only) */
static int g_crf_kernel_mode = 0;
void orc_segment_set_kernel_mode(int mode) { g_crf_kernel_mode = mode; }

/* ---------------------------------------------------------------- Segmentation::performSegmentationCRF */
void orc_seg_default_params(OrcSegParams* p) {
  /* GUI defaults (GUI/Tools/GUI.h:212-227) which override the class defaults each frame
   * (GUI/MainController.cpp:463-473) */
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

int orc_segment_crf(const uint8_t* rgb, const float* depth, int W, int H, int numModels, const unsigned char* modelIds,
                    const float* const* icpError, const float* const* vertConf4, unsigned char nextModelID,
                    int allowNew, const OrcSegParams* prm, uint8_t* fullSeg, OrcModelData* md, int* hasNewLabel,
                    int* slicLabelsOut, float* unaryOut, uint8_t* lowMapOut) {
  const float MAX_DEPTH = 100;
  const int spixelSize = 16;
  SlicCtx s;
  s.W = W;
  s.H = H;
  s.spixelSize = spixelSize;
  s.spixelX = W / spixelSize;
  s.spixelY = H / spixelSize;
  s.spixelNum = s.spixelX * s.spixelY;
  const int N = s.spixelNum, lowW = s.spixelX, lowH = s.spixelY;
  const int numLabels = allowNew ? numModels + 1 : numModels;
  int* labels = (int*)malloc(sizeof(int) * W * H);
  /* Slic::setInputImage swaps red and blue (Slic.cpp:52-60); the SLIC distance is symmetric in the
   * channels, so the restatement feeds the buffer as is */
  orc_slic(rgb, W, H, spixelSize, 5, 0.6f, labels);
  s.labels = labels;
  s.counts = (unsigned*)calloc(N, sizeof(unsigned));
  for (int i = 0; i < W * H; ++i) s.counts[labels[i]]++;
  if (slicLabelsOut) memcpy(slicLabelsOut, labels, sizeof(int) * W * H);

  /* Slic::downsample() (Slic.cpp:82-112): integer sums, integer division */
  uint8_t* lowRGB = (uint8_t*)calloc(N * 3, 1);
  {
    int* sums = (int*)calloc(N * 3, sizeof(int));
    for (int i = 0; i < W * H; ++i)
      for (int k = 0; k < 3; ++k) sums[labels[i] * 3 + k] += rgb[i * 3 + k];
    for (int index = 0; index < N; index++) {
      int cnt = (int)s.counts[index], readIndex = index;
      if (cnt == 0) {
        readIndex = resample_empty_index(&s, index);
        cnt = (int)s.counts[readIndex];
      }
      for (int k = 0; k < 3; ++k) lowRGB[index * 3 + k] = (uint8_t)(sums[readIndex * 3 + k] / cnt);
    }
    free(sums);
  }
  float* lowDepth = (float*)malloc(sizeof(float) * N);
  slic_downsample_thresholded(&s, depth, 0.02f, lowDepth);

  float depthMin = 3.402823466e+38f, depthMax = 0;
  for (int i = 0; i < N; i++) {
    float d = lowDepth[i];
    if (d > MAX_DEPTH || d < 0 || !isfinite(d)) continue;
    if (depthMax < d) depthMax = d;
    if (depthMin > d) depthMin = d;
  }
  const float depthRange = depthMax - depthMin;

  float** lowICP = (float**)malloc(sizeof(float*) * numModels);
  float** lowConf = (float**)malloc(sizeof(float*) * numModels);
  unsigned char modelIdToIndex[256];
  memset(modelIdToIndex, 0, sizeof(modelIdToIndex));
  for (int m = 0; m < numLabels; ++m) {
    memset(&md[m], 0, sizeof(OrcModelData));
    md[m].top = 65535;
    md[m].left = 65535;
  }
  for (int m = 0; m < numModels; ++m) {
    lowICP[m] = (float*)malloc(sizeof(float) * N);
    lowConf[m] = (float*)malloc(sizeof(float) * N);
    slic_downsample_f(&s, icpError[m], 1, 0, lowICP[m]);
    slic_downsample_f(&s, vertConf4[m], 4, 3, lowConf[m]);
    md[m].id = modelIds[m];
    modelIdToIndex[modelIds[m]] = (unsigned char)m;
    float avg = 0;
    for (int j = 0; j < N; j++) {
      float* c = &lowConf[m][j];
      if (!isfinite(*c)) {
        *c = 0;
        continue;
      }
      avg += *c;
    }
    md[m].avgConfidence = avg / (float)N;
  }
  if (allowNew) {
    modelIdToIndex[nextModelID] = (unsigned char)numModels;
    md[numModels].id = nextModelID;
  }

  /* unaries (Segmentation.cpp:237-298); unary is numLabels x N, stored node major here */
  float* unary = (float*)calloc((size_t)N * numLabels, sizeof(float));
  for (int k = 0; k < N; k++) {
    if (lowConf[0][k] < 0.3f) lowICP[0][k] = depthRange * 0.01f;
    for (int i = 1; i < numModels; i++)
      if (lowConf[i][k] <= 0.4f) lowICP[i][k] = depthRange * prm->unaryKError;
    float lowestError = lowICP[0][k] / depthRange;
    for (int i = 0; i < numModels; i++) { /* multimap with equal keys iterates in insertion order */
      float error = lowICP[i][k] / depthRange;
      if (error < lowestError) lowestError = error;
      unary[k * numLabels + i] = prm->unaryWeightError * error;
    }
    if (allowNew) {
      float u = prm->unaryThresholdNew - prm->unaryWeightError * lowestError;
      unary[k * numLabels + numModels] = u > 0.01f ? u : 0.01f;
    }
  }
  for (size_t i = 0; i < (size_t)N * numLabels; ++i)
    if (unary[i] <= 1e-5f) unary[i] = 1e-5f; /* :458-460 */
  if (unaryOut) memcpy(unaryOut, unary, sizeof(float) * N * numLabels);

  /* pairwise kernels: 2-D Gaussian (sx = sy = 2, :437) and the 6-D appearance kernel (:439-452) */
  float* f2 = (float*)malloc(sizeof(float) * N * 2);
  float* f6 = (float*)malloc(sizeof(float) * N * 6);
  for (int j = 0; j < lowH; j++)
    for (int i = 0; i < lowW; i++) {
      int index = j * lowW + i;
      f2[index * 2 + 0] = (float)i / 2.0f;
      f2[index * 2 + 1] = (float)j / 2.0f;
      f6[index * 6 + 0] = (float)i * prm->scaleFeaturesPos;
      f6[index * 6 + 1] = (float)j * prm->scaleFeaturesPos;
      f6[index * 6 + 2] = (float)rgb[index * 3 + 0] * prm->scaleFeaturesRGB; /* quirk: full-res buffer, low-res index */
      f6[index * 6 + 3] = (float)rgb[index * 3 + 1] * prm->scaleFeaturesRGB;
      f6[index * 6 + 4] = (float)rgb[index * 3 + 2] * prm->scaleFeaturesRGB;
      float fd = lowDepth[index] * prm->scaleFeaturesDepth;
      f6[index * 6 + 5] = fd < 100.0f ? fd : 100.0f;
    }
  const int lattice = g_crf_kernel_mode == 1;
  float* K2 = (float*)malloc(sizeof(float) * (lattice ? 1 : (size_t)N * N));
  float* K6 = (float*)malloc(sizeof(float) * (lattice ? 1 : (size_t)N * N));
  float* n2 = (float*)malloc(sizeof(float) * N);
  float* n6 = (float*)malloc(sizeof(float) * N);
  OrcLattice *L2 = NULL, *L6 = NULL;
  if (lattice) {
    L2 = orc_lattice_create(f2, 2, N);
    L6 = orc_lattice_create(f6, 6, N);
    orc_lattice_norm(L2, n2);
    orc_lattice_norm(L6, n6);
  } else {
    build_kernel(f2, 2, N, K2, n2);
    build_kernel(f6, 6, N, K6, n6);
  }

  /* mean field (:455-471) */
  const int L = numLabels;
  float* Q = (float*)malloc(sizeof(float) * N * L);
  float* t1 = (float*)malloc(sizeof(float) * N * L);
  float* t2 = (float*)malloc(sizeof(float) * N * L);
  for (int i = 0; i < N * L; ++i) t1[i] = -unary[i];
  exp_and_normalize(Q, t1, L, N);
  for (int it = 0; it < prm->crfIterations; it++) {
    for (int i = 0; i < N * L; ++i) t1[i] = -unary[i];
    if (lattice)
      orc_lattice_apply(L2, n2, Q, L, t2);
    else
      kernel_apply(K2, n2, Q, N, L, t2); /* Potts: tmp2 = -w * filtered ; tmp1 -= tmp2 */
    for (int i = 0; i < N * L; ++i) t1[i] -= -prm->weightSmoothness * t2[i];
    if (lattice)
      orc_lattice_apply(L6, n6, Q, L, t2);
    else
      kernel_apply(K6, n6, Q, N, L, t2);
    for (int i = 0; i < N * L; ++i) t1[i] -= -prm->weightAppearance * t2[i];
    exp_and_normalize(Q, t1, L, N);
  }
  uint8_t* map = (uint8_t*)malloc(N);
  for (int i = 0; i < N; i++) { /* maxCoeff: first maximum */
    int best = 0;
    for (int l = 1; l < L; ++l)
      if (Q[i * L + l] > Q[i * L + best]) best = l;
    map[i] = (uint8_t)md[best].id;
  }

  /* connected components + post-processing (:484-649) */
  int* comp = (int*)malloc(sizeof(int) * N);
  Comp* cc = (Comp*)malloc(sizeof(Comp) * N);
  const int ncc = connected_labels(map, lowH, lowW, comp, cc);
  /* labelToComponents (ConnectedLabels.hpp:40-48): std::map<int, std::list<int>> keyed by the CRF
   * label, lists in component order, built BEFORE any relabelling.  Simulated with explicit lists. */
  int* lists = (int*)malloc(sizeof(int) * 256 * (size_t)(ncc > 0 ? ncc : 1));
  int listN[256];
  memset(listN, 0, sizeof(listN));
  for (int c = 0; c < ncc; ++c) lists[cc[c].label * ncc + listN[cc[c].label]++] = c;
  int firstLabel = -1;
  for (int lab = 0; lab < 256 && firstLabel < 0; ++lab)
    if (listN[lab]) firstLabel = lab;
  /* onlyKeepLargest (:496-517): every label except the first entry of the map keeps only its
   * largest component (the earlier one on ties); the others become 255 and leave the list */
  for (int lab = 0; lab < 256; ++lab) {
    if (lab == firstLabel || listN[lab] == 0) continue;
    int cur = lists[lab * ncc];
    for (int q = 1; q < listN[lab]; ++q) {
      int c2 = lists[lab * ncc + q];
      if (cc[cur].size < cc[c2].size) {
        cc[cur].label = 255;
        cur = c2;
      } else
        cc[c2].label = 255;
    }
    lists[lab * ncc] = cur;
    listN[lab] = 1;
  }
  if (allowNew) { /* :521-530 */
    const int minSize = (int)((float)N * prm->minRelSizeNew), maxSize = (int)((float)N * prm->maxRelSizeNew);
    for (int q = 0; q < listN[nextModelID]; ++q) {
      int c = lists[nextModelID * ncc + q];
      if (cc[c].size < minSize || cc[c].size > maxSize) cc[c].label = 255;
    }
  }
  for (int m = 0; m < numLabels; ++m) { /* bounding boxes (:533-547) */
    OrcModelData* d = &md[m];
    for (int q = 0; q < listN[d->id]; ++q) {
      const Comp* st = &cc[lists[d->id * ncc + q]];
      if (st->left < (int)d->left) d->left = (unsigned short)st->left;
      if (st->top < (int)d->top) d->top = (unsigned short)st->top;
      if (st->right > (int)d->right) d->right = (unsigned short)st->right;
      if (st->bottom > (int)d->bottom) d->bottom = (unsigned short)st->bottom;
    }
    /* Slic::mapToHigh(x, y) (Slic.h:189-191) stored back into unsigned shorts */
    int px = (int)(d->left * spixelSize + spixelSize * 0.5), py = (int)(d->top * spixelSize + spixelSize * 0.5);
    d->left = (unsigned short)px;
    d->top = (unsigned short)py;
    px = (int)(d->right * spixelSize + spixelSize * 0.5);
    py = (int)(d->bottom * spixelSize + spixelSize * 0.5);
    d->right = (unsigned short)px;
    d->bottom = (unsigned short)py;
  }
  const unsigned borderSize = 20;
  for (int m = 0; m < numLabels; ++m) { /* :549-563 */
    OrcModelData* d = &md[m];
    if (d->id == 0) continue;
    if ((d->top < borderSize && d->bottom < borderSize) || (d->left < borderSize && d->right < borderSize) ||
        (d->top > (unsigned)H - borderSize && d->bottom > (unsigned)H - borderSize) ||
        (d->left > (unsigned)W - borderSize && d->right > (unsigned)W - borderSize)) {
      for (int q = 0; q < listN[d->id]; ++q) cc[lists[d->id * ncc + q]].label = 255;
    }
  }
  free(lists);
  for (int i = 0; i < N; i++) map[i] = cc[comp[i]].label;

  { /* depth statistics (:570-621) */
    float* sumsDepth = (float*)calloc(numLabels, sizeof(float));
    float* sumsDev = (float*)calloc(numLabels, sizeof(float));
    unsigned* cnts = (unsigned*)calloc(numLabels, sizeof(unsigned));
    for (int i = 0; i < N; i++) {
      if (map[i] == 255) continue;
      int idx = modelIdToIndex[map[i]];
      sumsDepth[idx] += lowDepth[i];
      cnts[idx]++;
    }
    for (int m = 0; m < numLabels; ++m) md[m].depthMean = cnts[m] ? sumsDepth[m] / (float)cnts[m] : 0;
    for (int i = 0; i < N; i++) {
      if (map[i] == 255) continue;
      int idx = modelIdToIndex[map[i]];
      sumsDev[idx] += fabsf(md[idx].depthMean - lowDepth[i]);
    }
    for (int m = 0; m < numLabels; ++m) md[m].depthStd = cnts[m] ? sumsDev[m] / (float)cnts[m] : 0;
    for (int i = 0; i < N; i++) {
      if (map[i] == 255) continue;
      int idx = modelIdToIndex[map[i]];
      if (idx != 0) {
        float d = lowDepth[i];
        if ((double)d > 1.1 * (double)md[idx].depthStd + (double)md[idx].depthMean) {
          sumsDepth[idx] -= d;
          sumsDev[idx] -= fabsf(md[idx].depthMean - d);
          cnts[idx]--;
        }
      }
    }
    for (int m = 0; m < numLabels; ++m) {
      md[m].depthMean = cnts[m] ? sumsDepth[m] / (float)cnts[m] : 0;
      md[m].depthStd = cnts[m] ? sumsDev[m] / (float)cnts[m] : 0;
    }
    free(sumsDepth);
    free(sumsDev);
    free(cnts);
  }
  for (int k = 0; k < N; k++) {
    if (map[k] == 255) continue;
    md[modelIdToIndex[map[k]]].superPixelCount++;
  }
  int outModels = numLabels;
  *hasNewLabel = 0;
  if (allowNew) {
    if (md[numModels].superPixelCount > 0)
      *hasNewLabel = 1;
    else
      outModels = numModels;
  }
  if (lowMapOut) memcpy(lowMapOut, map, N);
  for (int i = 0; i < W * H; ++i) fullSeg[i] = map[labels[i]]; /* Slic::upsample */

  for (int m = 0; m < numModels; ++m) {
    free(lowICP[m]);
    free(lowConf[m]);
  }
  free(lowICP);
  free(lowConf);
  free(labels);
  free(s.counts);
  free(lowRGB);
  free(lowDepth);
  free(unary);
  free(f2);
  free(f6);
  orc_lattice_destroy(L2);
  orc_lattice_destroy(L6);
  free(K2);
  free(K6);
  free(n2);
  free(n6);
  free(Q);
  free(t1);
  free(t2);
  free(map);
  free(comp);
  free(cc);
  return outModels;
}
