/*
 * lattice.c -- SECOND WITNESS for the dense-CRF message passing (TEST INFRASTRUCTURE, see cf_oracle.h).
 *
 * The reference hands its mean-field inference to martinruenz/densecrf (Scripts/install.sh:84, unpinned,
 * not under /root/reference; call sites Core/Segmentation/Segmentation.cpp:221, :436-437, :452, :462-471).
 * That library evaluates the Gaussian kernel products K Q approximately on a permutohedral lattice
 * (Adams, Baek, Davis: "Fast high-dimensional filtering using the permutohedral lattice", 2010, as
 * implemented in Kraehenbuehl & Koltun's densecrf `permutohedral.cpp`): embed every feature vector in the
 * (d+1)-dimensional hyperplane lattice, splat its values to the d+1 vertices of the enclosing simplex with
 * barycentric weights, blur along the d+1 lattice axes with the [1/2 1 1/2] stencil, slice back, scale by
 * alpha = 1 / (1 + 2^-d).  The oracle of record (segment.c) and the CUDA kernels evaluate the same
 * products EXACTLY over the 1200 super-pixels.  This file restates the lattice so that the difference
 * between the two realisations can be MEASURED (tests/test_oracle_segment.py::test_lattice_witness,
 * DESIGN.md section 2) instead of argued about.  NORMALIZE_SYMMETRIC as at the call sites:
 *     norm = 1 / sqrt(lattice(1) + 1e-20),   out = norm .* lattice(norm .* in).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "cf_oracle.h"

typedef struct {
  int d, cap, size; /* key length, table capacity (power of two), entries */
  short* keys;      /* size * d */
  int* table;       /* cap, -1 = empty */
} LHash;

static unsigned lhash_of(const short* k, int d) {
  unsigned r = 0;
  for (int i = 0; i < d; ++i) {
    r += (unsigned)(unsigned short)k[i];
    r *= 1664525u;
  }
  return r;
}
static void lhash_init(LHash* h, int d, int expected) {
  h->d = d;
  h->cap = 1;
  while (h->cap < 2 * expected) h->cap <<= 1;
  h->size = 0;
  h->keys = (short*)malloc(sizeof(short) * (size_t)expected * d);
  h->table = (int*)malloc(sizeof(int) * h->cap);
  for (int i = 0; i < h->cap; ++i) h->table[i] = -1;
}
static int lhash_find(LHash* h, const short* k, int create) {
  unsigned pos = lhash_of(k, h->d) & (unsigned)(h->cap - 1);
  for (;;) {
    const int e = h->table[pos];
    if (e == -1) {
      if (!create) return -1;
      memcpy(h->keys + (size_t)h->size * h->d, k, sizeof(short) * h->d);
      h->table[pos] = h->size;
      return h->size++;
    }
    if (!memcmp(h->keys + (size_t)e * h->d, k, sizeof(short) * h->d)) return e;
    pos = (pos + 1) & (unsigned)(h->cap - 1);
  }
}
static void lhash_free(LHash* h) {
  free(h->keys);
  free(h->table);
}

struct OrcLattice {
  int N, d, M;
  int* offset;      /* N * (d+1) */
  float* bary;      /* N * (d+1) */
  int* n1;          /* (d+1) * M */
  int* n2;
};

OrcLattice* orc_lattice_create(const float* feat, int d, int N) {
  OrcLattice* L = (OrcLattice*)calloc(1, sizeof(OrcLattice));
  L->N = N;
  L->d = d;
  L->offset = (int*)malloc(sizeof(int) * (size_t)N * (d + 1));
  L->bary = (float*)malloc(sizeof(float) * (size_t)N * (d + 1));
  LHash h;
  lhash_init(&h, d, N * (d + 1));
  float* scale = (float*)malloc(sizeof(float) * d);
  float* elevated = (float*)malloc(sizeof(float) * (d + 1));
  float* rem0 = (float*)malloc(sizeof(float) * (d + 1));
  float* bary = (float*)malloc(sizeof(float) * (d + 2));
  short* rank = (short*)malloc(sizeof(short) * (d + 1));
  short* canonical = (short*)malloc(sizeof(short) * (d + 1) * (d + 1));
  short* key = (short*)malloc(sizeof(short) * (d + 1));
  for (int i = 0; i <= d; ++i) {
    for (int j = 0; j <= d - i; ++j) canonical[i * (d + 1) + j] = (short)i;
    for (int j = d - i + 1; j <= d; ++j) canonical[i * (d + 1) + j] = (short)(i - (d + 1));
  }
  const float inv_std_dev = sqrtf(2.0f / 3.0f) * (float)(d + 1);
  for (int i = 0; i < d; ++i) scale[i] = 1.0f / sqrtf((float)((i + 2) * (i + 1))) * inv_std_dev;
  for (int k = 0; k < N; ++k) {
    const float* f = feat + (size_t)k * d;
    float sm = 0;
    for (int j = d; j > 0; --j) { /* elevate into the hyperplane sum = 0 */
      const float cf = f[j - 1] * scale[j - 1];
      elevated[j] = sm - (float)j * cf;
      sm += cf;
    }
    elevated[0] = sm;
    const float down_factor = 1.0f / (float)(d + 1), up_factor = (float)(d + 1);
    float sum = 0;
    for (int i = 0; i <= d; ++i) { /* nearest remainder-0 lattice point */
      const float v = down_factor * elevated[i];
      const float up = ceilf(v) * up_factor, down = floorf(v) * up_factor;
      rem0[i] = (up - elevated[i] < elevated[i] - down) ? up : down;
      sum += rem0[i] * down_factor;
    }
    const int isum = (int)sum;
    for (int i = 0; i <= d; ++i) rank[i] = 0;
    for (int i = 0; i < d; ++i) {
      const float di = elevated[i] - rem0[i];
      for (int j = i + 1; j <= d; ++j)
        if (di < elevated[j] - rem0[j])
          rank[i]++;
        else
          rank[j]++;
    }
    for (int i = 0; i <= d; ++i) { /* the rounded point may lie off the plane: walk back */
      rank[i] = (short)(rank[i] + isum);
      if (rank[i] < 0) {
        rank[i] = (short)(rank[i] + d + 1);
        rem0[i] += (float)(d + 1);
      } else if (rank[i] > d) {
        rank[i] = (short)(rank[i] - (d + 1));
        rem0[i] -= (float)(d + 1);
      }
    }
    for (int i = 0; i <= d + 1; ++i) bary[i] = 0;
    for (int i = 0; i <= d; ++i) {
      const float v = (elevated[i] - rem0[i]) * down_factor;
      bary[d - rank[i]] += v;
      bary[d - rank[i] + 1] -= v;
    }
    bary[0] += 1.0f + bary[d + 1];
    for (int r = 0; r <= d; ++r) {
      for (int i = 0; i < d; ++i) key[i] = (short)(rem0[i] + (float)canonical[r * (d + 1) + rank[i]]);
      L->offset[(size_t)k * (d + 1) + r] = lhash_find(&h, key, 1);
      L->bary[(size_t)k * (d + 1) + r] = bary[r];
    }
  }
  const int M = h.size;
  L->M = M;
  L->n1 = (int*)malloc(sizeof(int) * (size_t)(d + 1) * M);
  L->n2 = (int*)malloc(sizeof(int) * (size_t)(d + 1) * M);
  short* a = (short*)malloc(sizeof(short) * (d + 1));
  short* b = (short*)malloc(sizeof(short) * (d + 1));
  for (int j = 0; j <= d; ++j)
    for (int i = 0; i < M; ++i) {
      const short* k0 = h.keys + (size_t)i * d;
      for (int k = 0; k < d; ++k) {
        a[k] = (short)(k0[k] - 1);
        b[k] = (short)(k0[k] + 1);
      }
      if (j < d) {
        a[j] = (short)(k0[j] + d);
        b[j] = (short)(k0[j] - d);
      }
      L->n1[(size_t)j * M + i] = lhash_find(&h, a, 0);
      L->n2[(size_t)j * M + i] = lhash_find(&h, b, 0);
    }
  free(a);
  free(b);
  free(scale);
  free(elevated);
  free(rem0);
  free(bary);
  free(rank);
  free(canonical);
  free(key);
  lhash_free(&h);
  return L;
}

void orc_lattice_destroy(OrcLattice* L) {
  if (!L) return;
  free(L->offset);
  free(L->bary);
  free(L->n1);
  free(L->n2);
  free(L);
}

/* out (N x vs) = slice(blur(splat(in (N x vs)))) * alpha */
void orc_lattice_compute(const OrcLattice* L, const float* in, int vs, float* out) {
  const int d = L->d, M = L->M, N = L->N;
  float* val = (float*)calloc((size_t)(M + 2) * vs, sizeof(float));
  float* nv = (float*)calloc((size_t)(M + 2) * vs, sizeof(float));
  for (int i = 0; i < N; ++i)
    for (int j = 0; j <= d; ++j) {
      const int o = L->offset[(size_t)i * (d + 1) + j] + 1;
      const float w = L->bary[(size_t)i * (d + 1) + j];
      for (int k = 0; k < vs; ++k) val[(size_t)o * vs + k] += w * in[(size_t)i * vs + k];
    }
  for (int j = 0; j <= d; ++j) {
    for (int i = 0; i < M; ++i) {
      const float* oc = val + (size_t)(i + 1) * vs;
      float* nc = nv + (size_t)(i + 1) * vs;
      const float* a = val + (size_t)(L->n1[(size_t)j * M + i] + 1) * vs; /* index -1 -> slot 0, which stays zero */
      const float* b = val + (size_t)(L->n2[(size_t)j * M + i] + 1) * vs;
      for (int k = 0; k < vs; ++k) nc[k] = oc[k] + 0.5f * (a[k] + b[k]);
    }
    float* t = val;
    val = nv;
    nv = t;
  }
  const float alpha = 1.0f / (1.0f + powf(2.0f, -(float)d));
  for (int i = 0; i < N; ++i) {
    for (int k = 0; k < vs; ++k) out[(size_t)i * vs + k] = 0;
    for (int j = 0; j <= d; ++j) {
      const int o = L->offset[(size_t)i * (d + 1) + j] + 1;
      const float w = L->bary[(size_t)i * (d + 1) + j];
      for (int k = 0; k < vs; ++k) out[(size_t)i * vs + k] += w * val[(size_t)o * vs + k] * alpha;
    }
  }
  free(val);
  free(nv);
}

/* DenseKernel::initLattice with NORMALIZE_SYMMETRIC: norm (N) */
void orc_lattice_norm(const OrcLattice* L, float* norm) {
  float* ones = (float*)malloc(sizeof(float) * L->N);
  for (int i = 0; i < L->N; ++i) ones[i] = 1.0f;
  orc_lattice_compute(L, ones, 1, norm);
  for (int i = 0; i < L->N; ++i) norm[i] = 1.0f / sqrtf(norm[i] + 1e-20f);
  free(ones);
}

/* DenseKernel::filter with NORMALIZE_SYMMETRIC: out = norm .* lattice(norm .* Q); Q, out are N x Lbl */
void orc_lattice_apply(const OrcLattice* L, const float* norm, const float* Q, int Lbl, float* out) {
  float* nq = (float*)malloc(sizeof(float) * (size_t)L->N * Lbl);
  for (int i = 0; i < L->N; ++i)
    for (int l = 0; l < Lbl; ++l) nq[(size_t)i * Lbl + l] = norm[i] * Q[(size_t)i * Lbl + l];
  orc_lattice_compute(L, nq, Lbl, out);
  for (int i = 0; i < L->N; ++i)
    for (int l = 0; l < Lbl; ++l) out[(size_t)i * Lbl + l] *= norm[i];
  free(nq);
}
