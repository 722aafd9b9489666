/*
 * shdr_oracle.c -- plain-C restatement of the reference's Linearization-Net per-pixel path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- PARITY UNPINNED: the reference delegates all arithmetic to
 * TensorFlow, which cannot be installed in this project's image, and ships no tests or golden vectors.  This file
 * is a second, independent restatement next to oracle/np_oracle.py (the two are cross-checked bit for bit in
 * tests/test_oracle_c.py) and the faster CPU baseline of bench.py (OpenMP over image rows; every output element is
 * still computed with the reference's own sequence of fp32 operations).
 *
 * Build:  gcc -O3 -mavx2 -fopenmp -fno-fast-math -ffp-contract=off -shared -fPIC oracle/shdr_oracle.c -o oracle/_build/libshdr_oracle.so
 *         (-ffp-contract=off: no FMA contraction, every fp32 operation rounds like the TF op it stands for)
 *
 * file:line citations are into the reference repository (ShinYwings/SingleHDR-tf2).  [TF-sem] marks TensorFlow
 * semantics that are not visible in the reference's own files.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* launchers such as torchrun export OMP_NUM_THREADS=1; the baseline is meant to use all the host threads it can */
void shdr_oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int shdr_oracle_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* tf.pad(..., 'REFLECT') with pad 1: index -1 -> 1, n -> n-2   [TF-sem] */
static inline int reflect1(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

/* tf.image.sobel_edges + reshape to [n,h,w,2c]   linearization_net.py:312-314
 * cross-correlation with Ky = [[-1,-2,-1],[0,0,0],[1,2,1]] (k=0) and Kx = Ky^T (k=1), taps accumulated from 0 in
 * row-major order like a depthwise VALID convolution [TF-sem]; output channel = ch*2 + k. */
void shdr_oracle_sobel6(const float* img, float* out, int n, int h, int w, int c) {
  static const float KY[3][3] = {{-1, -2, -1}, {0, 0, 0}, {1, 2, 1}};
  static const float KX[3][3] = {{-1, 0, 1}, {-2, 0, 2}, {-1, 0, 1}};
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < n; ++b)
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x)
        for (int ch = 0; ch < c; ++ch) {
          float dy = 0.0f, dx = 0.0f;
          for (int r = 0; r < 3; ++r)
            for (int s = 0; s < 3; ++s) {
              const int yy = reflect1(y + r - 1, h), xx = reflect1(x + s - 1, w);
              const float p = img[(((size_t)b * h + yy) * w + xx) * c + ch];
              if (KY[r][s] != 0.0f) dy = dy + KY[r][s] * p;
              if (KX[r][s] != 0.0f) dx = dx + KX[r][s] * p;
            }
          float* o = out + (((size_t)b * h + y) * w + x) * 2 * c + ch * 2;
          o[0] = dy;
          o[1] = dx;
        }
}

/* model.histogram_layer(img, bins)   linearization_net.py:336-350
 * out[.., (i-1)*c + ch] = (d < 1/B) ? 1 - d*B : 0 with d = |img - (2i-1)/(2B)|; constants are fp32 [TF-sem]. */
void shdr_oracle_hist(const float* img, float* out, long long npx, int c, int bins) {
  const float thr = (float)(1.0 / (double)bins);        /* :339  python double -> fp32 tensor */
  const float nb = (float)bins, two_b = (float)(2.0 * bins);
#pragma omp parallel for schedule(static)
  for (long long p = 0; p < npx; ++p)
    for (int i = 0; i < bins; ++i) {
      const float centre = (float)(2.0 * (i + 1) - 1.0) / two_b;   /* :342,345  tf.divide of two fp32 values */
      for (int ch = 0; ch < c; ++ch) {
        const float d = fabsf(img[p * c + ch] - centre);            /* :345 */
        const float hv = 1.0f - d * nb;                             /* :346  multiply, then subtract */
        out[p * (long long)c * bins + (long long)i * c + ch] = (d < thr) ? hv : 0.0f;
      }
    }
}

/* average_pooling2d(x, k, 1, 'same')   linearization_net.py:351 (dead code there), README.md:51
 * [TF-sem] SAME: (k-1)/2 taps before, the rest after; mean over the in-bounds taps only; a window is accumulated in
 * input raster order starting from 0. */
void shdr_oracle_avg_pool_same(const float* x, float* out, int n, int h, int w, int c, int k) {
  const int pb = (k - 1) / 2, pa = k - 1 - pb;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < n; ++b)
    for (int y = 0; y < h; ++y) {
      const int y0 = y - pb < 0 ? 0 : y - pb, y1 = y + pa > h - 1 ? h - 1 : y + pa;
      for (int xx = 0; xx < w; ++xx) {
        const int x0 = xx - pb < 0 ? 0 : xx - pb, x1 = xx + pa > w - 1 ? w - 1 : xx + pa;
        const float cnt = (float)((y1 - y0 + 1) * (x1 - x0 + 1));
        float* o = out + (((size_t)b * h + y) * w + xx) * c;
        for (int ch = 0; ch < c; ++ch) o[ch] = 0.0f;
        for (int yy = y0; yy <= y1; ++yy)
          for (int xs = x0; xs <= x1; ++xs) {
            const float* p = x + (((size_t)b * h + yy) * w + xs) * c;
            for (int ch = 0; ch < c; ++ch) o[ch] = o[ch] + p[ch];
          }
        for (int ch = 0; ch < c; ++ch) o[ch] = o[ch] / cnt;
      }
    }
}

/* AEInvcrfDecodeNet.invcrf_pca_w_2_invcrf   linearization_net.py:231-253: g0 + hinv[s,11] . w[b,11] */
void shdr_oracle_pca(const float* w, const float* g0, const float* hinv, float* out, int b, int s, int ncomp) {
  for (int i = 0; i < b; ++i)
    for (int k = 0; k < s; ++k) {
      float acc = 0.0f;
      for (int j = 0; j < ncomp; ++j) acc = acc + hinv[(size_t)k * ncomp + j] * w[(size_t)i * ncomp + j];
      out[(size_t)i * s + k] = g0[k] + acc;
    }
}

/* model._increase   linearization_net.py:368-392 (sum and cumsum sequential in fp32) */
void shdr_oracle_increase(const float* rf, float* out, int b, int k) {
  float* g = (float*)malloc(sizeof(float) * (size_t)(k > 1 ? k - 1 : 1));
  for (int i = 0; i < b; ++i) {
    const float* r = rf + (size_t)i * k;
    float mn = INFINITY;
    for (int j = 0; j < k - 1; ++j) { g[j] = r[j + 1] - r[j]; if (g[j] < mn) mn = g[j]; }     /* :370-373 */
    const float rl = (-mn > 0.0f) ? -mn : 0.0f;                                                   /* :377 relu(-min) */
    float sum = 0.0f;
    for (int j = 0; j < k - 1; ++j) { g[j] = g[j] + rl; sum = sum + g[j]; }                       /* :380-383 */
    float acc = 0.0f;
    out[(size_t)i * k] = 0.0f;                                                                    /* :389 pad */
    for (int j = 0; j < k - 1; ++j) { acc = acc + g[j] / sum; out[(size_t)i * k + j + 1] = acc; } /* :383-386 */
  }
  free(g);
}

/* tf_utils.apply_rf -> interp_1d -> sample_1d   tf_utils.py:95-105, 70-93, 54-68 */
void shdr_oracle_apply_rf(const float* x, const float* rf, float* y, int b, long long per, int k) {
  const float km1 = (float)(k - 1);
#pragma omp parallel for schedule(static)
  for (long long e = 0; e < (long long)b * per; ++e) {
    const float* r = rf + (e / per) * k;
    const float yy = km1 * x[e];                    /* :103 */
    const float y0 = floorf(yy), y1 = y0 + 1.0f;    /* :77-78 */
    /* tf.cast(float -> int32) then clip (:82, :66); values far outside int32 are saturated here (TF: undefined) */
    long long i0 = (y0 != y0) ? 0 : (y0 < -2e9f ? -2000000000LL : (y0 > 2e9f ? 2000000000LL : (long long)y0));
    long long i1 = (y1 != y1) ? 0 : (y1 < -2e9f ? -2000000000LL : (y1 > 2e9f ? 2000000000LL : (long long)y1));
    if (i0 < 0) i0 = 0;
    if (i0 > k - 1) i0 = k - 1;
    if (i1 < 0) i1 = 0;
    if (i1 > k - 1) i1 = k - 1;
    const float w0 = y1 - yy, w1 = yy - y0;          /* :87-88 */
    const float a = w0 * r[i0], c = w1 * r[i1];      /* :93  mul, mul, add */
    y[e] = a + c;
  }
}

/* concat([img, edge6, hist(bins[0]), ...], -1), each histogram optionally pooled   linearization_net.py:312-322 */
void shdr_oracle_frontend(const float* img, float* out, int n, int h, int w, const int* bins, int nbins, int pool_k,
                          int with_img_edge) {
  const long long npx = (long long)n * h * w;
  int C = with_img_edge ? 9 : 0;
  for (int i = 0; i < nbins; ++i) C += 3 * bins[i];
  int off = 0;
  if (with_img_edge) {
    float* e = (float*)malloc(sizeof(float) * (size_t)npx * 6);
    shdr_oracle_sobel6(img, e, n, h, w, 3);
#pragma omp parallel for schedule(static)
    for (long long p = 0; p < npx; ++p) {
      for (int ch = 0; ch < 3; ++ch) out[p * C + ch] = img[p * 3 + ch];
      for (int ch = 0; ch < 6; ++ch) out[p * C + 3 + ch] = e[p * 6 + ch];
    }
    free(e);
    off = 9;
  }
  for (int i = 0; i < nbins; ++i) {
    const int cb = 3 * bins[i];
    float* hst = (float*)malloc(sizeof(float) * (size_t)npx * cb);
    shdr_oracle_hist(img, hst, npx, 3, bins[i]);
    if (pool_k) {
      float* pl = (float*)malloc(sizeof(float) * (size_t)npx * cb);
      shdr_oracle_avg_pool_same(hst, pl, n, h, w, cb, pool_k);
      free(hst);
      hst = pl;
    }
#pragma omp parallel for schedule(static)
    for (long long p = 0; p < npx; ++p) memcpy(out + p * C + off, hst + p * cb, sizeof(float) * (size_t)cb);
    free(hst);
    off += cb;
  }
}
