// backward.cu -- reverse-mode gradients of the per-pixel path, so that the native ops can
// sit inside the reference's training steps behind tf.custom_gradient:
//   tf_utils.apply_rf        w.r.t. x and rf   train.py:186-194, finetune_real_dataset.py:149-178
//   model._increase          w.r.t. rf         linearization_net.py:368-392
//   invcrf_pca_w_2_invcrf    w.r.t. w          linearization_net.py:231-253
//   front end (img, Sobel, soft histograms) w.r.t. img   linearization_net.py:312-350 (finetune_real_dataset.py only:
//                                                         there the front end's input is the Dequantization-Net's output)
// Each kernel restates what TensorFlow's autodiff does for the reference's op sequence (floor / cast / compare have no
// gradient; tf.abs -> sign; tf.where routes the gradient to the taken branch; reduce_min splits the gradient evenly
// among ties; gather_nd -> scatter-add).
#include "common.cuh"

namespace shdr {

// ------------------------------------------------------------------ apply_rf backward
// out = (y1 - y) * rf[i0] + (y - y0) * rf[i1],  y = (k-1) x,  y0 = floor(y), y1 = y0 + 1, i* = clip(int(y*), 0, k-1)
//   d/dx      = g (k-1) (rf[i1] - rf[i0])
//   d/drf[i0] += g (y1 - y),   d/drf[i1] += g (y - y0)     -- a per-image k-bin weighted histogram:
// accumulated in shared memory per CTA, flushed with one global atomic per bin per CTA.
constexpr int BWD_THREADS = 256;

template <bool NEED_GX, bool NEED_GRF>
__global__ void __launch_bounds__(BWD_THREADS)
k_apply_rf_bwd(const float* __restrict__ x, const float* __restrict__ rf, const float* __restrict__ gy,
               float* __restrict__ gx, float* __restrict__ grf, long long elems_per_item, int k, int chunks_per_item,
               long long elems_per_chunk) {
  extern __shared__ float sm[];
  float* tab = sm;           // rf[k]
  float* acc = sm + k;       // d/drf[k]
  const int tid = threadIdx.x;
  const long long item = blockIdx.x / chunks_per_item;
  const int chunk = blockIdx.x - (int)(item * chunks_per_item);
  const float* r = rf + item * k;
  for (int i = tid; i < k; i += BWD_THREADS) {
    tab[i] = r[i];
    if (NEED_GRF) acc[i] = 0.0f;
  }
  __syncthreads();
  const float km1 = (float)(k - 1);
  const int kmax = k - 1;
  const long long e0 = (long long)chunk * elems_per_chunk;
  const long long e1 = min(e0 + elems_per_chunk, elems_per_item);
  const float* xi = x + item * elems_per_item;
  const float* gi = gy + item * elems_per_item;
  float* gxi = NEED_GX ? gx + item * elems_per_item : nullptr;
  if (!NEED_GRF) {                                                 // d/dx only: a plain stream
    for (long long e = e0 + tid; e < e1; e += BWD_THREADS) {
      const float xv = __ldg(xi + e), g = __ldg(gi + e);
      const float y = __fmul_rn(km1, xv);
      const float y0 = floorf(y);
      const float y1 = __fadd_rn(y0, 1.0f);
      const int i0 = min(max(__float2int_rz(y0), 0), kmax);
      const int i1 = min(max(__float2int_rz(y1), 0), kmax);
      gxi[e] = km1 * (g * (tab[i1] - tab[i0]));
    }
    return;
  }
  const int lane = tid & 31;
#pragma unroll 2
  for (long long eb = e0; eb < e1; eb += BWD_THREADS) {           // warp-uniform trip count (shuffles below)
    const long long e = eb + tid;
    const bool valid = e < e1;
    const float xv = valid ? __ldg(xi + e) : 0.0f, g = valid ? __ldg(gi + e) : 0.0f;
    const float y = __fmul_rn(km1, xv);
    const float y0 = floorf(y);
    const float y1 = __fadd_rn(y0, 1.0f);
    const int i0 = min(max(__float2int_rz(y0), 0), kmax);
    const int i1 = min(max(__float2int_rz(y1), 0), kmax);
    // difference first: neighbouring curve samples are close, so their difference is (nearly) exact in fp32, while
    // g*rf[i1] - g*rf[i0] would amplify the products' rounding by k - 1
    if (NEED_GX && valid) gxi[e] = km1 * (g * (tab[i1] - tab[i0]));
    if (NEED_GRF) {
      float a0 = g * __fsub_rn(y1, y), a1 = g * __fsub_rn(y, y0);
      // Natural images are smooth: neighbouring elements fall into the same bin, and 32 same-address shared-memory
      // atomics serialise (measured 3.2 ms vs 0.27 ms for uniform-random input on 16 x 1024^2).  When much of the
      // warp repeats its neighbour's bin, each RUN of equal consecutive bins is summed with a segmented scan and the
      // last lane of the run issues the two atomics.  (Runs, not all equal keys: a non-adjacent repeat just costs one
      // more atomic.)
      const unsigned full = 0xffffffffu;
      const int key = valid ? i0 : -1 - lane;                       // invalid lanes: runs of their own, no atomics
      const int prev = __shfl_up_sync(full, key, 1);
      const unsigned cont = __ballot_sync(full, lane > 0 && prev == key);   // lanes that continue a run
      bool tail = true;
      if (__popc(cont) >= 12) {
        const int start = 31 - __clz(~cont & (0xffffffffu >> (31 - lane)));   // first lane of this lane's run
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float b0 = __shfl_up_sync(full, a0, o), b1 = __shfl_up_sync(full, a1, o);
          if (lane - o >= start) { a0 += b0; a1 += b1; }
        }
        tail = (lane == 31) || !((cont >> (lane + 1)) & 1u);
      }
      if (tail && valid) {
        atomicAdd(acc + i0, a0);
        atomicAdd(acc + i1, a1);
      }
    }
  }
  if (NEED_GRF) {
    __syncthreads();
    float* go = grf + item * k;
    for (int i = tid; i < k; i += BWD_THREADS) {
      const float v = acc[i];
      if (v != 0.0f) atomicAdd(go + i, v);
    }
  }
}

template <bool NEED_GX, bool NEED_GRF>
static int launch_apply_bwd_t(const float* x, const float* rf, const float* gy, float* gx, float* grf, int b,
                              long long n, int k, cudaStream_t st, int dev) {
  const long long quantum = 4096;
  long long per_chunk = quantum * 2;       // 8192 elements per CTA (one flush of k atomics each): ~3 waves of CTAs
  const long long want = (long long)sm_count(dev) * 4;
  while (per_chunk > quantum && (long long)b * ((n + per_chunk - 1) / per_chunk) < want) per_chunk >>= 1;
  const long long chunks = (n + per_chunk - 1) / per_chunk;
  const long long grid = (long long)b * chunks;
  SHDR_REQUIRE(grid > 0 && grid <= 0x7fffffffLL, "apply_rf_bwd: grid of %lld CTAs is out of range", grid);
  const size_t smem = (size_t)k * 2 * sizeof(float);
  if (smem > 48 * 1024)
    SHDR_CUDA(cudaFuncSetAttribute(k_apply_rf_bwd<NEED_GX, NEED_GRF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
  k_apply_rf_bwd<NEED_GX, NEED_GRF><<<(unsigned)grid, BWD_THREADS, smem, st>>>(x, rf, gy, gx, grf, n, k, (int)chunks,
                                                                              per_chunk);
  SHDR_LAUNCH_CHECK("k_apply_rf_bwd");
  return SHDR_OK;
}

// ------------------------------------------------------------------ _increase / PCA backward
constexpr int CB_THREADS = 1024;

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float wmin(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum / min of one value per thread; every thread gets the result
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = wsum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float t = (lane < (int)(blockDim.x >> 5)) ? red[lane] : 0.0f;
  return wsum(t);
}
__device__ __forceinline__ float block_min(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = wmin(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float t = (lane < (int)(blockDim.x >> 5)) ? red[lane] : __int_as_float(0x7f800000);
  return wmin(t);
}

// One CTA per curve.
//   w != NULL : the curve is the PCA reconstruction g0 + hinv.w (k == 1024); the result is d/dw[11]
//   w == NULL : the curve is rf_in[k]; the result is d/drf[k]
//   monotone  : gout is the gradient of _increase(curve); else gout is the gradient of the curve itself (w != NULL)
// _increase forward: g = diff(v); m = min g; r = relu(-m); u = g + r; s = sum u; n = u / s; out = [0, cumsum(n)]
// backward:  dn_i = sum_{j >= i} gout[j+1];  D = sum dn_i n_i;  du_i = (dn_i - D) / s;  dr = sum du_i;
//            dg_i = du_i - [m < 0] dr [g_i == m] / ties;   dv_j = dg_{j-1} - dg_j
__global__ void __launch_bounds__(CB_THREADS)
k_curve_bwd(const float* __restrict__ w, const float* __restrict__ rf_in, const float* __restrict__ g0,
            const float* __restrict__ hinv, const float* __restrict__ gout, float* __restrict__ gres, int k,
            int monotone) {
  extern __shared__ float sm[];
  float* v = sm;              // curve, k
  float* dg = sm + k;         // gradient w.r.t. the diffs (k - 1), later w.r.t. the curve (k)
  __shared__ float red[32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const long long item = blockIdx.x;
  const float* go = gout + item * k;

  if (w != nullptr) {
    float wj[SHDR_EMOR_NCOMP];
#pragma unroll
    for (int j = 0; j < SHDR_EMOR_NCOMP; ++j) wj[j] = w[item * SHDR_EMOR_NCOMP + j];
    for (int s = tid; s < k; s += CB_THREADS) {
      float a = 0.0f;
#pragma unroll
      for (int j = 0; j < SHDR_EMOR_NCOMP; ++j) a = fmaf(hinv[j * SHDR_EMOR_SAMPLES + s], wj[j], a);
      v[s] = __fadd_rn(g0[s], a);
    }
  } else {
    for (int s = tid; s < k; s += CB_THREADS) v[s] = rf_in[item * k + s];
  }
  __syncthreads();

  if (monotone) {
    const int n = k - 1;
    const int seg = (n + CB_THREADS - 1) / CB_THREADS;
    const int lo = min(tid * seg, n), hi = min(lo + seg, n);
    float m = __int_as_float(0x7f800000);
    for (int i = lo; i < hi; ++i) m = fminf(m, __fsub_rn(v[i + 1], v[i]));
    m = block_min(m, red);
    float ties = 0.0f;
    for (int i = lo; i < hi; ++i) ties += (__fsub_rn(v[i + 1], v[i]) == m) ? 1.0f : 0.0f;
    ties = block_sum(ties, red);
    const float r = fmaxf(-m, 0.0f);
    float s = 0.0f;
    for (int i = lo; i < hi; ++i) s += __fadd_rn(__fsub_rn(v[i + 1], v[i]), r);
    s = block_sum(s, red);
    // reverse inclusive cumsum of gout[1..k-1]: dn_i = sum_{j >= i} gout[j + 1]
    float local = 0.0f;
    for (int i = lo; i < hi; ++i) local += go[i + 1];
    // exclusive scan from the right over threads: suffix = sum of `local` of threads with a larger id
    float incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_down_sync(0xffffffffu, incl, o);
      if (lane + o < 32) incl += t;
    }
    __syncthreads();
    if (lane == 0) red[wid] = incl;        // sum of this warp
    __syncthreads();
    float after = 0.0f;                    // sum of the warps after this one
    for (int q = wid + 1; q < (int)(blockDim.x >> 5); ++q) after += red[q];
    float run = after + incl - local;      // sum over threads strictly after this one
    float dsum = 0.0f;                     // sum dn_i n_i  (n_i = u_i / s)
    for (int i = hi - 1; i >= lo; --i) {
      run += go[i + 1];
      dg[i] = run;                         // dn_i for now
      dsum += run * (__fadd_rn(__fsub_rn(v[i + 1], v[i]), r) / s);
    }
    const float D = block_sum(dsum, red);
    float dr = 0.0f;
    for (int i = lo; i < hi; ++i) {
      const float du = (dg[i] - D) / s;
      dg[i] = du;
      dr += du;
    }
    dr = block_sum(dr, red);
    const float dm = (m < 0.0f) ? -dr / ties : 0.0f;       // relu'(-m) = 1 for -m > 0; reduce_min splits among ties
    for (int i = lo; i < hi; ++i)
      if (__fsub_rn(v[i + 1], v[i]) == m) dg[i] += dm;
    __syncthreads();
    // d/dv_j = dg_{j-1} - dg_j; written over v (no longer needed once dg is complete)
    for (int j = tid; j < k; j += CB_THREADS) {
      const float a = (j > 0) ? dg[j - 1] : 0.0f;
      const float b2 = (j < n) ? dg[j] : 0.0f;
      v[j] = a - b2;
    }
    __syncthreads();
  } else {
    for (int j = tid; j < k; j += CB_THREADS) v[j] = go[j];
    __syncthreads();
  }

  if (w != nullptr) {
    // d/dw_j = sum_s hinv[s][j] d/dcurve_s
    for (int j = 0; j < SHDR_EMOR_NCOMP; ++j) {
      float a = 0.0f;
      for (int s = tid; s < k; s += CB_THREADS) a = fmaf(hinv[j * SHDR_EMOR_SAMPLES + s], v[s], a);
      a = block_sum(a, red);
      if (tid == 0) gres[item * SHDR_EMOR_NCOMP + j] = a;
    }
  } else {
    for (int j = tid; j < k; j += CB_THREADS) gres[item * k + j] = v[j];
  }
}

static int launch_curve_bwd(const float* w, const float* rf_in, const float* gout, float* gres, int b, int k,
                            int monotone, cudaStream_t st, int dev) {
  const float *g0 = nullptr, *hinv = nullptr;
  if (w != nullptr) {
    int rc = emor_device_table(dev, &g0, &hinv);
    if (rc != SHDR_OK) return rc;
  }
  const size_t smem = (size_t)k * 2 * sizeof(float);
  if (smem > 48 * 1024)
    SHDR_CUDA(cudaFuncSetAttribute(k_curve_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_curve_bwd<<<b, CB_THREADS, smem, st>>>(w, rf_in, g0, hinv, gout, gres, k, monotone);
  SHDR_LAUNCH_CHECK("k_curve_bwd");
  return SHDR_OK;
}

// ------------------------------------------------------------------ front end backward (un-pooled)
// gimg[p][c] = gfeat[p][c]                                                    (the img slice of the concat)
//            + sum over bins of the three histograms: -B sign(I - centre) gfeat[p][hist channel]  where |I - centre| < 1/B
//            + the transpose of the REFLECT-padded Sobel correlation applied to gfeat[.][3..8].
// (row, tap) pairs whose source row is py: the direct ones (py - r + 1, r) and the reflected ones: output row 0 reads
// row 1 through tap 0 (index -1 -> 1), output row h-1 reads row h-2 through tap 2 (index h -> h-2).
__device__ __forceinline__ int src_pairs(int p, int n, int (&o)[5], int (&t)[5]) {
  int c = 0;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int y = p - r + 1;
    if (y >= 0 && y < n) { o[c] = y; t[c] = r; ++c; }
  }
  if (p == 1) { o[c] = 0; t[c] = 0; ++c; }
  if (p == n - 2) { o[c] = n - 1; t[c] = 2; ++c; }
  return c;
}

template <bool FULL>
__global__ void __launch_bounds__(256)
k_frontend_bwd(const float* __restrict__ img, const float* __restrict__ gfeat, float* __restrict__ gimg, long long npx,
               int h, int w, int c, int bins) {
  // FULL: gfeat is [npx, 93] (c == 3, histograms 4 / 8 / 16); else gfeat is [npx, c * bins] of one histogram_layer
  const float ky[3][3] = {{-1.f, -2.f, -1.f}, {0.f, 0.f, 0.f}, {1.f, 2.f, 1.f}};
  const float kx[3][3] = {{-1.f, 0.f, 1.f}, {-2.f, 0.f, 2.f}, {-1.f, 0.f, 1.f}};
  const long long total = npx * c;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long p = e / c;
    const int ch = (int)(e - p * c);
    const float v = __ldg(img + e);
    float acc = 0.0f;
    if (FULL) {
      const float* gf = gfeat + p * SHDR_FRONTEND_CH;
      acc = __ldg(gf + ch);
      int off = 9;
#pragma unroll
      for (int bi = 0; bi < 3; ++bi) {
        const int B = 4 << bi;
        const float thr = 1.0f / (float)B;
        for (int b = 0; b < B; ++b) {
          const float centre = (float)(2 * b + 1) / (float)(2 * B);
          const float d = __fsub_rn(v, centre);
          if (fabsf(d) < thr) {
            const float sg = (d > 0.0f) ? 1.0f : ((d < 0.0f) ? -1.0f : 0.0f);
            acc -= (float)B * sg * __ldg(gf + off + b * 3 + ch);
          }
        }
        off += 3 * B;
      }
      // Sobel transpose
      const long long row = p / w;
      const int px = (int)(p - row * w);
      const int py = (int)(row % h);
      const long long img0 = (row - py) * w;             // first pixel of this image
      int oy[5], ty[5], ox[5], tx[5];
      const int ny = src_pairs(py, h, oy, ty), nx = src_pairs(px, w, ox, tx);
      for (int a = 0; a < ny; ++a)
        for (int b2 = 0; b2 < nx; ++b2) {
          const float* ge = gfeat + (img0 + (long long)oy[a] * w + ox[b2]) * SHDR_FRONTEND_CH + 3 + ch * 2;
          acc += ky[ty[a]][tx[b2]] * __ldg(ge) + kx[ty[a]][tx[b2]] * __ldg(ge + 1);
        }
    } else {
      const float* gh = gfeat + p * (long long)c * bins;
      const float thr = (float)(1.0 / (double)bins);
      const float nb = (float)bins, two_b = (float)(2 * bins);
      for (int b = 0; b < bins; ++b) {
        const float centre = __fdiv_rn((float)(2 * b + 1), two_b);
        const float d = __fsub_rn(v, centre);
        if (fabsf(d) < thr) {
          const float sg = (d > 0.0f) ? 1.0f : ((d < 0.0f) ? -1.0f : 0.0f);
          acc -= nb * sg * __ldg(gh + b * c + ch);
        }
      }
    }
    gimg[e] = acc;
  }
}

}  // namespace shdr

using namespace shdr;

extern "C" int shdr_apply_rf_bwd_f32(const float* x, const float* rf, const float* gy, float* gx, float* grf, int b,
                                     long long elems_per_item, int k, void* stream) {
  SHDR_REQUIRE(b >= 0 && elems_per_item >= 0, "apply_rf_bwd: b=%d elems_per_item=%lld", b, elems_per_item);
  SHDR_REQUIRE(k >= 1 && k <= 24576, "apply_rf_bwd: k=%d (need 1..24576)", k);
  SHDR_REQUIRE(gx || grf, "apply_rf_bwd: both gradient outputs are NULL");
  if (b == 0) return SHDR_OK;
  SHDR_REQUIRE(x && rf && gy, "apply_rf_bwd: NULL pointer");
  DeviceGuard g(gx ? gx : grf);
  if (g.status != SHDR_OK) return g.status;
  cudaStream_t st = (cudaStream_t)stream;
  if (grf) SHDR_CUDA(cudaMemsetAsync(grf, 0, (size_t)b * k * sizeof(float), st));
  if (elems_per_item == 0) return SHDR_OK;
  if (gx && grf) return launch_apply_bwd_t<true, true>(x, rf, gy, gx, grf, b, elems_per_item, k, st, g.dev);
  if (gx) return launch_apply_bwd_t<true, false>(x, rf, gy, gx, grf, b, elems_per_item, k, st, g.dev);
  return launch_apply_bwd_t<false, true>(x, rf, gy, gx, grf, b, elems_per_item, k, st, g.dev);
}

extern "C" int shdr_increase_bwd_f32(const float* rf, const float* gout, float* grf, int b, int k, void* stream) {
  SHDR_REQUIRE(rf && gout && grf, "increase_bwd: NULL pointer");
  SHDR_REQUIRE(b >= 0 && k >= 2 && k <= 24576, "increase_bwd: b=%d k=%d (need b>=0, 2<=k<=24576)", b, k);
  if (b == 0) return SHDR_OK;
  DeviceGuard g(grf);
  if (g.status != SHDR_OK) return g.status;
  return launch_curve_bwd(nullptr, rf, gout, grf, b, k, 1, (cudaStream_t)stream, g.dev);
}

extern "C" int shdr_invcrf_build_bwd_f32(const float* w, const float* gcurve, float* gw, int b, int monotone,
                                         void* stream) {
  SHDR_REQUIRE(w && gcurve && gw, "invcrf_build_bwd: NULL pointer");
  SHDR_REQUIRE(b >= 0, "invcrf_build_bwd: b=%d", b);
  if (b == 0) return SHDR_OK;
  DeviceGuard g(gw);
  if (g.status != SHDR_OK) return g.status;
  return launch_curve_bwd(w, nullptr, gcurve, gw, b, SHDR_EMOR_SAMPLES, monotone, (cudaStream_t)stream, g.dev);
}

static unsigned bwd_grid(long long total, int dev) {
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count(dev) * 32;
  return (unsigned)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

extern "C" int shdr_frontend_bwd_f32(const float* img, const float* gfeat, float* gimg, int n, int h, int w,
                                     void* stream) {
  SHDR_REQUIRE(n >= 0 && h >= 0 && w >= 0, "frontend_bwd: bad shape n=%d h=%d w=%d", n, h, w);
  const long long npx = (long long)n * h * w;
  if (npx == 0) return SHDR_OK;
  SHDR_REQUIRE(img && gfeat && gimg, "frontend_bwd: NULL pointer");
  SHDR_REQUIRE(h >= 2 && w >= 2, "frontend_bwd: REFLECT padding needs h >= 2 and w >= 2 (got %d x %d)", h, w);
  SHDR_REQUIRE(npx < 0x7fffffffLL, "frontend_bwd: n*h*w = %lld pixels does not fit int32", npx);
  DeviceGuard g(gimg);
  if (g.status != SHDR_OK) return g.status;
  k_frontend_bwd<true><<<bwd_grid(npx * 3, g.dev), 256, 0, (cudaStream_t)stream>>>(img, gfeat, gimg, npx, h, w, 3, 0);
  SHDR_LAUNCH_CHECK("k_frontend_bwd");
  return SHDR_OK;
}

extern "C" int shdr_soft_hist_bwd_f32(const float* img, const float* ghist, float* gimg, int n, int h, int w, int c,
                                      int bins, void* stream) {
  SHDR_REQUIRE(n >= 0 && h >= 0 && w >= 0 && c >= 1, "soft_hist_bwd: bad shape n=%d h=%d w=%d c=%d", n, h, w, c);
  SHDR_REQUIRE(bins >= 1 && bins <= 4096, "soft_hist_bwd: bins=%d (need 1..4096)", bins);
  const long long npx = (long long)n * h * w;
  if (npx == 0) return SHDR_OK;
  SHDR_REQUIRE(img && ghist && gimg, "soft_hist_bwd: NULL pointer");
  DeviceGuard g(gimg);
  if (g.status != SHDR_OK) return g.status;
  k_frontend_bwd<false><<<bwd_grid(npx * c, g.dev), 256, 0, (cudaStream_t)stream>>>(img, ghist, gimg, npx, h, w, c,
                                                                                     bins);
  SHDR_LAUNCH_CHECK("k_frontend_bwd");
  return SHDR_OK;
}
