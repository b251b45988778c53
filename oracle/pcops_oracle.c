/*
 * pcops_oracle.c -- CPU oracle for the PointNet++ geometry-op hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (libpcops.so) never links, imports or calls anything in oracle/.
 *
 * Every function restates one reference kernel op-for-op in plain C, citing the
 * reference file:line it follows (paths relative to
 * /root/reference/pointnet2_tensorflow/tf_ops/ unless noted).  Arithmetic contract
 * (BASELINE.json north_star): fp32, un-fused multiply/add, evaluated left to right
 * exactly as the reference expressions are written; build with -ffp-contract=off.
 *
 * Parity pinning (see oracle/README.md and tests/test_oracle_pins.py):
 *   ball query / group / group-grad  -> pinned against grouping/test/query_ball_point.cpp
 *                                       compiled unmodified into oracle/_ref/libref_cpu.so
 *   selection sort                   -> pinned against grouping/test/selection_sort.cpp
 *                                       (its own known-answer b=2,n=4,m=2,k=3) in _ref
 *   three_nn / interpolate / grad    -> pinned against interpolation_3d/tf_interpolate.cpp
 *                                       compiled unmodified (TF headers stubbed) in _ref
 *   FPS / gather / gather-grad / prob_sample -> pinned on the GPU box against sampling/tf_sampling_g.cu
 *                                       compiled unmodified (--fmad=false) in _ref
 *   kNN distances, FP weights, attention -> TensorFlow ops in the reference; TF is absent,
 *                                       so these three are "parity unpinned" (source-pinned only).
 *
 * The *_omp entry points run the same loops with OpenMP over the batch dimension; they
 * exist only for the cpu_baseline timing in bench.py.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * a1  farthest point sampling -- sampling/tf_sampling_g.cu:105-170
 *
 * The reference kernel runs 512 threads per scene.  Thread t visits k = t, t+512, ... in
 * ascending order with a strict '>' against best=-1/besti=0 (:125-126,:146-149); the tree
 * (:153-163) lets the right slot win only if strictly larger.  The emulation below keeps
 * the 512 lanes and the tree literally, so ties resolve exactly as on the GPU
 * (smallest k mod 512, then smallest k).
 * ------------------------------------------------------------------------------------------ */
static void fps_one(int n, int m, const float *dataset, float *temp, int *idxs) {
  enum { BS = 512 };
  float dists[BS];
  int dists_i[BS];
  if (m <= 0) return;                                   /* :106-107 */
  int old = 0;
  idxs[0] = old;                                        /* :114-116 */
  for (int j = 0; j < n; ++j) temp[j] = 1e38f;          /* :117-119 */
  for (int j = 1; j < m; ++j) {                         /* :123 */
    float x1 = dataset[old * 3 + 0];                    /* :127-129 */
    float y1 = dataset[old * 3 + 1];
    float z1 = dataset[old * 3 + 2];
    for (int t = 0; t < BS; ++t) {
      int besti = 0;                                    /* :124 */
      float best = -1.0f;                               /* :125 */
      for (int k = t; k < n; k += BS) {                 /* :130 */
        float td = temp[k];
        float x2 = dataset[k * 3 + 0], y2 = dataset[k * 3 + 1], z2 = dataset[k * 3 + 2];
        float dx = x2 - x1, dy = y2 - y1, dz = z2 - z1;
        float d = (dx * dx + dy * dy) + dz * dz;        /* :142, un-fused, left to right */
        float d2 = fminf(d, td);                        /* :143 CUDA min(float,float) == fminf */
        if (d2 != td) temp[k] = d2;                     /* :144-145 */
        if (d2 > best) { best = d2; besti = k; }        /* :146-149 */
      }
      dists[t] = best;                                  /* :151-152 */
      dists_i[t] = besti;
    }
    for (int u = 0; (1 << u) < BS; ++u) {               /* :153 */
      for (int t = 0; t < (BS >> (u + 1)); ++t) {       /* :155 */
        int i1 = (t * 2) << u;
        int i2 = (t * 2 + 1) << u;
        if (dists[i1] < dists[i2]) {                    /* :158 */
          dists[i1] = dists[i2];
          dists_i[i1] = dists_i[i2];
        }
      }
    }
    old = dists_i[0];                                   /* :165 */
    idxs[j] = old;                                      /* :166-167 */
  }
}

ORC_API void orc_fps(int b, int n, int m, const float *xyz, int *out) {
  float *temp = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < b; ++i) fps_one(n, m, xyz + (size_t)i * n * 3, temp, out + (size_t)i * m);
  free(temp);
}

ORC_API void orc_fps_omp(int b, int n, int m, const float *xyz, int *out) {
#pragma omp parallel
  {
    float *temp = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
#pragma omp for schedule(dynamic, 1)
    for (int i = 0; i < b; ++i) fps_one(n, m, xyz + (size_t)i * n * 3, temp, out + (size_t)i * m);
    free(temp);
  }
}

/* ------------------------------------------------------------------------------------------
 * a0  prob_sample -- sampling/tf_sampling_g.cu:7-104 (cumsumKernel + binarysearchKernel), launcher :197-200.
 * Registered next to FPS in the sampling library, called by no model; restated because it is part of the
 * op surface.  The cumulative sum is NOT a left-to-right sum: per block of 8192 values the kernel forms
 *   - inside each group of four: p1 = v1, p2 = v2 + v1, p3 = v3 + p2, p4 = (v4 + v3) + p2       (:19-32)
 *     (a trailing partial group is summed serially from 0, :33-42)
 *   - an in-place scan of the group totals: up-sweep  G[((2k+2)<<u)-1] += G[((2k+1)<<u)-1]       (:47-56)
 *                                           down-sweep G[((2k+3)<<u)-1] += G[((2k+2)<<u)-1]      (:58-67)
 *   - element = in-group prefix + total of the groups before it (:69-77), + running sum of earlier blocks (:79-81)
 *   - the running sum is carried across blocks with a compensation term                           (:82-85)
 * and the fp32 roundings follow that tree.  The sample for uniform r is the smallest index whose
 * cumulative value is >= r * total, found by the power-of-two descent of :92-101.
 * ------------------------------------------------------------------------------------------ */
static void cumsum_row(int n, const float *inp, float *out) {
  enum { BLOCK = 8192 };
  static __thread float p[BLOCK], G[BLOCK / 4];
  float runningsum = 0.f, runningsum2 = 0.f;
  for (int j = 0; j < n; j += BLOCK) {
    const int cnt = n - j < BLOCK ? n - j : BLOCK;
    const int n24 = (cnt + 3) & ~3, n2 = n24 >> 2;
    for (int k = 0; k < cnt; k += 4) {
      if (k + 3 < cnt) {
        float v1 = inp[j + k], v2 = inp[j + k + 1], v3 = inp[j + k + 2], v4 = inp[j + k + 3];
        v2 += v1;
        v4 += v3;
        v3 += v2;
        v4 += v2;
        p[k] = v1; p[k + 1] = v2; p[k + 2] = v3; p[k + 3] = v4;
        G[k >> 2] = v4;
      } else {
        float v = 0.f;
        for (int k2 = k; k2 < cnt; ++k2) { v += inp[j + k2]; p[k2] = v; }
        for (int k2 = cnt; k2 < n24; ++k2) p[k2] = v;
        G[k >> 2] = v;
      }
    }
    int u = 0;
    for (; (2 << u) <= n2; ++u)
      for (int k = 0; k < (n2 >> (u + 1)); ++k) G[(((k << 1) + 2) << u) - 1] += G[(((k << 1) + 1) << u) - 1];
    for (--u; u >= 0; --u)
      for (int k = 0; k < ((n2 - (1 << u)) >> (u + 1)); ++k) G[(((k << 1) + 3) << u) - 1] += G[(((k << 1) + 2) << u) - 1];
    for (int k = 4; k < n24; ++k) p[k] += G[(k >> 2) - 1];
    for (int k = 0; k < cnt; ++k) out[j + k] = p[k] + runningsum;
    const float t = G[n2 - 1] + runningsum2;
    const float r2 = runningsum + t;
    runningsum2 = t - (r2 - runningsum);
    runningsum = r2;
  }
}

ORC_API void orc_cumsum(int b, int n, const float *inp, float *out) {
  for (int i = 0; i < b; ++i) cumsum_row(n, inp + (size_t)i * n, out + (size_t)i * n);
}

ORC_API void orc_prob_sample(int b, int n, int m, const float *inp_p, const float *inp_r, float *temp, int *out) {
  int base = 1;
  while (base < n) base <<= 1;
  for (int i = 0; i < b; ++i) {
    float *cdf = temp + (size_t)i * n;
    cumsum_row(n, inp_p + (size_t)i * n, cdf);
    for (int j = 0; j < m; ++j) {
      const float q = inp_r[(size_t)i * m + j] * cdf[n - 1];
      int r = n - 1;
      for (int k = base; k >= 1; k >>= 1)
        if (r >= k && cdf[r - k] >= q) r -= k;
      out[(size_t)i * m + j] = r;
    }
  }
}

/* a2  gather_point -- sampling/tf_sampling_g.cu:172-181 */
ORC_API void orc_gather_point(int b, int n, int m, const float *inp, const int *idx, float *out) {
  for (int i = 0; i < b; ++i)
    for (int j = 0; j < m; ++j) {
      int a = idx[(size_t)i * m + j];
      for (int l = 0; l < 3; ++l) out[((size_t)i * m + j) * 3 + l] = inp[((size_t)i * n + a) * 3 + l];
    }
}

/* a3  gather_point_grad -- sampling/tf_sampling_g.cu:183-192 after the zero-fill at
 * sampling/tf_sampling.cpp:174.  The reference uses float atomicAdd (order undefined); the
 * contract here is the serial order j = 0..m-1, the same order the reference's CPU programs
 * use for their scatter-adds (grouping/test/query_ball_point.cpp:70-84). */
ORC_API void orc_gather_point_grad(int b, int n, int m, const float *out_g, const int *idx, float *inp_g) {
  memset(inp_g, 0, sizeof(float) * (size_t)b * n * 3);
  for (int i = 0; i < b; ++i)
    for (int j = 0; j < m; ++j) {
      int a = idx[(size_t)i * m + j];
      for (int l = 0; l < 3; ++l) inp_g[((size_t)i * n + a) * 3 + l] += out_g[((size_t)i * m + j) * 3 + l];
    }
}

/* ------------------------------------------------------------------------------------------
 * a4  query_ball_point -- grouping/tf_grouping_g.cu:3-36 (== grouping/test/query_ball_point.cpp:19-47
 * plus pts_cnt from tf_grouping_g.cu:34).  Rows of empty balls are zero (the CPU program
 * pre-zeroes idx at query_ball_point.cpp:94; the GPU op leaves them uninitialised).
 * ------------------------------------------------------------------------------------------ */
static void ball_one(int n, int m, float radius, int nsample, const float *xyz1, const float *xyz2, int *idx,
                     int *pts_cnt) {
  for (int j = 0; j < m; ++j) {
    int cnt = 0;
    for (int l = 0; l < nsample; ++l) idx[(size_t)j * nsample + l] = 0;
    float x2 = xyz2[j * 3 + 0], y2 = xyz2[j * 3 + 1], z2 = xyz2[j * 3 + 2];
    for (int k = 0; k < n; ++k) {
      if (cnt == nsample) break;                                             /* :16-17 */
      float x1 = xyz1[k * 3 + 0], y1 = xyz1[k * 3 + 1], z1 = xyz1[k * 3 + 2];
      float dx = x2 - x1, dy = y2 - y1, dz = z2 - z1;
      float d = fmaxf(sqrtf((dx * dx + dy * dy) + dz * dz), 1e-20f);         /* :24 */
      if (d < radius) {                                                       /* :25 */
        if (cnt == 0)
          for (int l = 0; l < nsample; ++l) idx[(size_t)j * nsample + l] = k; /* :26-29 */
        idx[(size_t)j * nsample + cnt] = k;                                   /* :30 */
        cnt += 1;
      }
    }
    pts_cnt[j] = cnt;                                                         /* :34 */
  }
}

ORC_API void orc_query_ball(int b, int n, int m, float radius, int nsample, const float *xyz1, const float *xyz2,
                            int *idx, int *pts_cnt) {
  for (int i = 0; i < b; ++i)
    ball_one(n, m, radius, nsample, xyz1 + (size_t)i * n * 3, xyz2 + (size_t)i * m * 3,
             idx + (size_t)i * m * nsample, pts_cnt + (size_t)i * m);
}

ORC_API void orc_query_ball_omp(int b, int n, int m, float radius, int nsample, const float *xyz1,
                                const float *xyz2, int *idx, int *pts_cnt) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int i = 0; i < b; ++i)
    ball_one(n, m, radius, nsample, xyz1 + (size_t)i * n * 3, xyz2 + (size_t)i * m * 3,
             idx + (size_t)i * m * nsample, pts_cnt + (size_t)i * m);
}

/* a5  group_point -- grouping/tf_grouping_g.cu:40-57 == grouping/test/query_ball_point.cpp:52-66 */
static void group_one(int n, int c, int m, int nsample, const float *points, const int *idx, float *out) {
  (void)n;
  for (int j = 0; j < m; ++j)
    for (int k = 0; k < nsample; ++k) {
      int ii = idx[(size_t)j * nsample + k];
      memcpy(out + ((size_t)j * nsample + k) * c, points + (size_t)ii * c, sizeof(float) * (size_t)c);
    }
}

ORC_API void orc_group_point(int b, int n, int c, int m, int nsample, const float *points, const int *idx,
                             float *out) {
  for (int i = 0; i < b; ++i)
    group_one(n, c, m, nsample, points + (size_t)i * n * c, idx + (size_t)i * m * nsample,
              out + (size_t)i * m * nsample * c);
}

ORC_API void orc_group_point_omp(int b, int n, int c, int m, int nsample, const float *points, const int *idx,
                                 float *out) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int i = 0; i < b; ++i)
    group_one(n, c, m, nsample, points + (size_t)i * n * c, idx + (size_t)i * m * nsample,
              out + (size_t)i * m * nsample * c);
}

/* a6  group_point_grad -- grouping/tf_grouping_g.cu:61-78 after the zero-fill at
 * grouping/tf_grouping.cpp:204; serial (j,k) order of grouping/test/query_ball_point.cpp:70-84. */
static void group_grad_one(int n, int c, int m, int nsample, const float *grad_out, const int *idx,
                           float *grad_points) {
  memset(grad_points, 0, sizeof(float) * (size_t)n * c);
  for (int j = 0; j < m; ++j)
    for (int k = 0; k < nsample; ++k) {
      int ii = idx[(size_t)j * nsample + k];
      const float *g = grad_out + ((size_t)j * nsample + k) * c;
      float *dst = grad_points + (size_t)ii * c;
      for (int l = 0; l < c; ++l) dst[l] += g[l];
    }
}

ORC_API void orc_group_point_grad(int b, int n, int c, int m, int nsample, const float *grad_out, const int *idx,
                                  float *grad_points) {
  for (int i = 0; i < b; ++i)
    group_grad_one(n, c, m, nsample, grad_out + (size_t)i * m * nsample * c, idx + (size_t)i * m * nsample,
                   grad_points + (size_t)i * n * c);
}

ORC_API void orc_group_point_grad_omp(int b, int n, int c, int m, int nsample, const float *grad_out,
                                      const int *idx, float *grad_points) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int i = 0; i < b; ++i)
    group_grad_one(n, c, m, nsample, grad_out + (size_t)i * m * nsample * c, idx + (size_t)i * m * nsample,
                   grad_points + (size_t)i * n * c);
}

/* ------------------------------------------------------------------------------------------
 * a7  selection sort -- grouping/tf_grouping_g.cu:83-123 == grouping/test/selection_sort.cpp:20-63
 * Full-length (b,m,n) outputs; the first k columns are the k smallest, in the order the
 * swap sequence leaves them (unstable among equal values).
 * ------------------------------------------------------------------------------------------ */
static void selsort_row(int n, int k, const float *dist, int *outi, float *out) {
  for (int s = 0; s < n; ++s) { out[s] = dist[s]; outi[s] = s; }   /* :96-101 */
  int kk = k < n ? k : n;
  for (int s = 0; s < kk; ++s) {                                     /* :107 */
    int min = s;
    for (int t = s + 1; t < n; ++t)
      if (out[t] < out[min]) min = t;                                /* :111 strict '<' */
    if (min != s) {                                                  /* :116-121 */
      float tmp = out[min]; out[min] = out[s]; out[s] = tmp;
      int tmpi = outi[min]; outi[min] = outi[s]; outi[s] = tmpi;
    }
  }
}

ORC_API void orc_selection_sort(int b, int n, int m, int k, const float *dist, int *outi, float *out) {
  for (size_t r = 0; r < (size_t)b * m; ++r) selsort_row(n, k, dist + r * n, outi + r * n, out + r * n);
}

/* kNN distance matrix -- grouping/tf_grouping.py:62-66:
 *   dist[b,j,k] = reduce_sum((xyz1[b,k,:] - xyz2[b,j,:])**2, -1), channels summed in order.
 * TensorFlow op; parity unpinned (TF absent), restated as left-to-right fp32. */
static inline float knn_d(const float *p1, const float *p2, int c) {
  float s = 0.0f;
  for (int l = 0; l < c; ++l) {
    float t = p1[l] - p2[l];
    float q = t * t;
    s = (l == 0) ? q : s + q;
  }
  return s;
}

ORC_API void orc_knn_dist(int b, int n, int m, int c, const float *xyz1, const float *xyz2, float *dist) {
  for (int i = 0; i < b; ++i)
    for (int j = 0; j < m; ++j)
      for (int k = 0; k < n; ++k)
        dist[((size_t)i * m + j) * n + k] = knn_d(xyz1 + ((size_t)i * n + k) * c, xyz2 + ((size_t)i * m + j) * c, c);
}

/* knn_point -- grouping/tf_grouping.py:48-73: distance matrix, SelectionSort, slice [:k].
 * Returns (val, idx) each (b,m,k). */
static void knn_rows(int i, int n, int m, int k, int c, const float *xyz1, const float *xyz2, float *val, int *idx,
                     float *row, float *orow, int *irow) {
  for (int j = 0; j < m; ++j) {
    for (int t = 0; t < n; ++t) row[t] = knn_d(xyz1 + ((size_t)i * n + t) * c, xyz2 + ((size_t)i * m + j) * c, c);
    selsort_row(n, k, row, irow, orow);
    for (int s = 0; s < k; ++s) {
      val[((size_t)i * m + j) * k + s] = orow[s];
      idx[((size_t)i * m + j) * k + s] = irow[s];
    }
  }
}

ORC_API void orc_knn(int b, int n, int m, int k, int c, const float *xyz1, const float *xyz2, float *val, int *idx) {
#pragma omp parallel
  {
    float *row = (float *)malloc(sizeof(float) * (size_t)(n + 1));
    float *orow = (float *)malloc(sizeof(float) * (size_t)(n + 1));
    int *irow = (int *)malloc(sizeof(int) * (size_t)(n + 1));
#pragma omp for schedule(dynamic, 1)
    for (int i = 0; i < b; ++i) knn_rows(i, n, m, k, c, xyz1, xyz2, val, idx, row, orow, irow);
    free(row); free(orow); free(irow);
  }
}

/* ------------------------------------------------------------------------------------------
 * a8  three_nn -- interpolation_3d/tf_interpolate.cpp:60-103.  The distance is evaluated in
 * float (all operands are float, :73) and widened to double only for the comparisons; best
 * slots start at 1e40 (:66) which becomes +inf when stored to the float output (:91-95).
 * ------------------------------------------------------------------------------------------ */
static void threenn_one(int n, int m, const float *xyz1, const float *xyz2, float *dist, int *idx) {
  for (int j = 0; j < n; ++j) {
    float x1 = xyz1[j * 3 + 0], y1 = xyz1[j * 3 + 1], z1 = xyz1[j * 3 + 2];
    double best1 = 1e40, best2 = 1e40, best3 = 1e40;
    int besti1 = 0, besti2 = 0, besti3 = 0;
    for (int k = 0; k < m; ++k) {
      float x2 = xyz2[k * 3 + 0], y2 = xyz2[k * 3 + 1], z2 = xyz2[k * 3 + 2];
      float dx = x2 - x1, dy = y2 - y1, dz = z2 - z1;
      float df = (dx * dx + dy * dy) + dz * dz;
      double d = df;
      if (d < best1) {
        best3 = best2; besti3 = besti2; best2 = best1; besti2 = besti1; best1 = d; besti1 = k;
      } else if (d < best2) {
        best3 = best2; besti3 = besti2; best2 = d; besti2 = k;
      } else if (d < best3) {
        best3 = d; besti3 = k;
      }
    }
    dist[j * 3 + 0] = (float)best1; idx[j * 3 + 0] = besti1;
    dist[j * 3 + 1] = (float)best2; idx[j * 3 + 1] = besti2;
    dist[j * 3 + 2] = (float)best3; idx[j * 3 + 2] = besti3;
  }
}

ORC_API void orc_three_nn(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int *idx) {
  for (int i = 0; i < b; ++i)
    threenn_one(n, m, xyz1 + (size_t)i * n * 3, xyz2 + (size_t)i * m * 3, dist + (size_t)i * n * 3,
                idx + (size_t)i * n * 3);
}

ORC_API void orc_three_nn_omp(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int *idx) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int i = 0; i < b; ++i)
    threenn_one(n, m, xyz1 + (size_t)i * n * 3, xyz2 + (size_t)i * m * 3, dist + (size_t)i * n * 3,
                idx + (size_t)i * n * 3);
}

/* a11 inverse-distance weights -- utils/pointnet_util.py:219-222 (TensorFlow elementwise ops):
 *   dist = max(dist, 1e-10); norm = sum(1/dist, axis=2); weight = (1/dist)/norm
 * parity unpinned (TF absent); restated in fp32, sum left to right. */
ORC_API void orc_three_weights(size_t rows, const float *dist, float *weight) {
  for (size_t r = 0; r < rows; ++r) {
    float d0 = fmaxf(dist[r * 3 + 0], 1e-10f), d1 = fmaxf(dist[r * 3 + 1], 1e-10f), d2 = fmaxf(dist[r * 3 + 2], 1e-10f);
    float r0 = 1.0f / d0, r1 = 1.0f / d1, r2 = 1.0f / d2;
    float norm = (r0 + r1) + r2;
    weight[r * 3 + 0] = r0 / norm;
    weight[r * 3 + 1] = r1 / norm;
    weight[r * 3 + 2] = r2 / norm;
  }
}

/* a9  three_interpolate -- interpolation_3d/tf_interpolate.cpp:107-127 */
static void interp_one(int m, int c, int n, const float *points, const int *idx, const float *weight, float *out) {
  (void)m;
  for (int j = 0; j < n; ++j) {
    float w1 = weight[j * 3], w2 = weight[j * 3 + 1], w3 = weight[j * 3 + 2];
    int i1 = idx[j * 3], i2 = idx[j * 3 + 1], i3 = idx[j * 3 + 2];
    for (int l = 0; l < c; ++l)
      out[(size_t)j * c + l] =
          (points[(size_t)i1 * c + l] * w1 + points[(size_t)i2 * c + l] * w2) + points[(size_t)i3 * c + l] * w3; /* :119 */
  }
}

ORC_API void orc_three_interpolate(int b, int m, int c, int n, const float *points, const int *idx,
                                   const float *weight, float *out) {
  for (int i = 0; i < b; ++i)
    interp_one(m, c, n, points + (size_t)i * m * c, idx + (size_t)i * n * 3, weight + (size_t)i * n * 3,
               out + (size_t)i * n * c);
}

ORC_API void orc_three_interpolate_omp(int b, int m, int c, int n, const float *points, const int *idx,
                                       const float *weight, float *out) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int i = 0; i < b; ++i)
    interp_one(m, c, n, points + (size_t)i * m * c, idx + (size_t)i * n * 3, weight + (size_t)i * n * 3,
               out + (size_t)i * n * c);
}

/* a10 three_interpolate_grad -- interpolation_3d/tf_interpolate.cpp:131-153 after the memset at :258 */
static void interp_grad_one(int n, int c, int m, const float *grad_out, const int *idx, const float *weight,
                            float *grad_points) {
  memset(grad_points, 0, sizeof(float) * (size_t)m * c);
  for (int j = 0; j < n; ++j) {
    float w1 = weight[j * 3], w2 = weight[j * 3 + 1], w3 = weight[j * 3 + 2];
    int i1 = idx[j * 3], i2 = idx[j * 3 + 1], i3 = idx[j * 3 + 2];
    for (int l = 0; l < c; ++l) {
      float g = grad_out[(size_t)j * c + l];
      grad_points[(size_t)i1 * c + l] += g * w1;          /* :145-147 */
      grad_points[(size_t)i2 * c + l] += g * w2;
      grad_points[(size_t)i3 * c + l] += g * w3;
    }
  }
}

ORC_API void orc_three_interpolate_grad(int b, int n, int c, int m, const float *grad_out, const int *idx,
                                        const float *weight, float *grad_points) {
  for (int i = 0; i < b; ++i)
    interp_grad_one(n, c, m, grad_out + (size_t)i * n * c, idx + (size_t)i * n * 3, weight + (size_t)i * n * 3,
                    grad_points + (size_t)i * m * c);
}

ORC_API void orc_three_interpolate_grad_omp(int b, int n, int c, int m, const float *grad_out, const int *idx,
                                            const float *weight, float *grad_points) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int i = 0; i < b; ++i)
    interp_grad_one(n, c, m, grad_out + (size_t)i * n * c, idx + (size_t)i * n * 3, weight + (size_t)i * n * 3,
                    grad_points + (size_t)i * m * c);
}

/* ------------------------------------------------------------------------------------------
 * a13 attention contraction -- /root/reference/attention_points/attention_scannet/attention_layer.py:29-45
 *
 * G neighbourhoods, S samples each, H heads of key_dim D, width HD = H*D.
 *   Q (G,HD)  K (G,S,HD)  V (G,S,HD)  ->  out (G,HD)
 * Heads come from a RAW row-major reshape (:35) of each neighbourhood's (S,HD) buffer to
 * (H,S,D):  K'[h,s,d] = Kflat[h*S*D + s*D + d];  Q'[h,d] = Q[h*D+d].
 *   logit[h,s] = sum_d Q'[h,d]*K'[h,s,d] / sqrt(D)   (:37-38)
 *   a = softmax_s(logit)                              (:39)
 *   out[h*D+d] = sum_s a[h,s]*V'[h,s,d]               (:40,:42)
 * TensorFlow matmul/softmax; parity unpinned (TF absent).  Accumulated in double so that
 * any fp32 evaluation order is within 1e-5 relative of it.
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_attention_fwd(int G, int S, int H, int D, const float *Q, const float *K, const float *V, float *out,
                               float *attn /* (G,H,S) or NULL */) {
  const int HD = H * D;
  const double scale = 1.0 / sqrt((double)D);
#pragma omp parallel
  {
    double *lg = (double *)malloc(sizeof(double) * (size_t)S);
#pragma omp for schedule(static)
    for (int g = 0; g < G; ++g) {
      const float *Kg = K + (size_t)g * S * HD, *Vg = V + (size_t)g * S * HD, *Qg = Q + (size_t)g * HD;
      for (int h = 0; h < H; ++h) {
        double mx = -INFINITY;
        for (int s = 0; s < S; ++s) {
          double acc = 0.0;
          for (int d = 0; d < D; ++d) acc += (double)Qg[h * D + d] * (double)Kg[(size_t)h * S * D + s * D + d];
          lg[s] = acc * scale;
          if (lg[s] > mx) mx = lg[s];
        }
        double den = 0.0;
        for (int s = 0; s < S; ++s) { lg[s] = exp(lg[s] - mx); den += lg[s]; }
        for (int s = 0; s < S; ++s) {
          lg[s] /= den;
          if (attn) attn[((size_t)g * H + h) * S + s] = (float)lg[s];
        }
        for (int d = 0; d < D; ++d) {
          double acc = 0.0;
          for (int s = 0; s < S; ++s) acc += lg[s] * (double)Vg[(size_t)h * S * D + s * D + d];
          out[(size_t)g * HD + h * D + d] = (float)acc;
        }
      }
    }
    free(lg);
  }
}

/* Gradient of the contraction above w.r.t. Q, K, V (what TF autodiff derives from :37-40). */
ORC_API void orc_attention_bwd(int G, int S, int H, int D, const float *Q, const float *K, const float *V,
                               const float *dout, float *dQ, float *dK, float *dV) {
  const int HD = H * D;
  const double scale = 1.0 / sqrt((double)D);
#pragma omp parallel
  {
    double *a = (double *)malloc(sizeof(double) * (size_t)S);
    double *da = (double *)malloc(sizeof(double) * (size_t)S);
#pragma omp for schedule(static)
    for (int g = 0; g < G; ++g) {
      const float *Kg = K + (size_t)g * S * HD, *Vg = V + (size_t)g * S * HD, *Qg = Q + (size_t)g * HD;
      const float *dog = dout + (size_t)g * HD;
      float *dKg = dK + (size_t)g * S * HD, *dVg = dV + (size_t)g * S * HD, *dQg = dQ + (size_t)g * HD;
      for (int h = 0; h < H; ++h) {
        double mx = -INFINITY;
        for (int s = 0; s < S; ++s) {
          double acc = 0.0;
          for (int d = 0; d < D; ++d) acc += (double)Qg[h * D + d] * (double)Kg[(size_t)h * S * D + s * D + d];
          a[s] = acc * scale;
          if (a[s] > mx) mx = a[s];
        }
        double den = 0.0;
        for (int s = 0; s < S; ++s) { a[s] = exp(a[s] - mx); den += a[s]; }
        double dot = 0.0;
        for (int s = 0; s < S; ++s) {
          a[s] /= den;
          double acc = 0.0;
          for (int d = 0; d < D; ++d) acc += (double)dog[h * D + d] * (double)Vg[(size_t)h * S * D + s * D + d];
          da[s] = acc;
          dot += a[s] * acc;
        }
        for (int d = 0; d < D; ++d) {
          double accq = 0.0;
          for (int s = 0; s < S; ++s) {
            double dl = a[s] * (da[s] - dot) * scale;
            accq += dl * (double)Kg[(size_t)h * S * D + s * D + d];
            dKg[(size_t)h * S * D + s * D + d] = (float)(dl * (double)Qg[h * D + d]);
            dVg[(size_t)h * S * D + s * D + d] = (float)(a[s] * (double)dog[h * D + d]);
          }
          dQg[h * D + d] = (float)accq;
        }
      }
    }
    free(a); free(da);
  }
}

/* Dense layer y = x W + b (tf.layers.Dense; attention_layer.py:24-26,31-34), rows x Cin -> Cout,
 * W stored (Cin, Cout) row-major as TensorFlow does.  Double accumulation. */
ORC_API void orc_dense(size_t rows, int cin, int cout, const float *x, const float *W, const float *bias, float *y) {
#pragma omp parallel for schedule(static)
  for (size_t r = 0; r < rows; ++r)
    for (int o = 0; o < cout; ++o) {
      double acc = bias ? (double)bias[o] : 0.0;
      for (int i = 0; i < cin; ++i) acc += (double)x[r * cin + i] * (double)W[(size_t)i * cout + o];
      y[r * cout + o] = (float)acc;
    }
}
