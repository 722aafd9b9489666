// invcrf.cu -- inverse-CRF stage: EMoR PCA reconstruction, monotonic enforcement,
// per-pixel linear-interpolated curve lookup.
//
// Reference behaviour restated (ShinYwings/SingleHDR-tf2):
//   invcrf_pca_w_2_invcrf   linearization_net.py:231-253   curve = g0 + hinv . w
//   _increase               linearization_net.py:368-392   diff, shift by relu(-min), normalise, cumsum, pad
//   apply_rf/interp_1d      tf_utils.py:95-105, 70-93      y=(k-1)x, floor, clip, lerp (mul, mul, add)
//
// Roofline: apply_rf is a pure stream (4 B read + 4 B written per element, 24 B/px for RGB)
// -> HBM-bound.  The curve kernels move ~4 KB per image and are latency-bound; they exist
// to keep the whole stage on the device and in one stream.
#include "common.cuh"

namespace shdr {

// ------------------------------------------------------------------ curve build
constexpr int CURVE_THREADS = 1024;

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_incl_scan(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v = __fadd_rn(v, t);
  }
  return v;
}

// One CTA per curve.  Either `w` (PCA weights, k == 1024) or `rf_in` (raw curve, any k) is given.
__global__ void __launch_bounds__(CURVE_THREADS)
k_curve(const float* __restrict__ w, const float* __restrict__ rf_in,
        const float* __restrict__ g0, const float* __restrict__ hinv,
        float* __restrict__ out, int k, int monotone) {
  extern __shared__ float v[];            // k samples of the un-enforced curve
  __shared__ float red[32];
  __shared__ float bcast[2];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const long long item = blockIdx.x;
  // programmatic dependent launch: the apply kernel that follows in shdr_linearize_f32 may start its CTAs now;
  // it blocks in griddepcontrol.wait until this grid has completed and its curves are visible
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (w != nullptr) {
    float wj[SHDR_EMOR_NCOMP];
#pragma unroll
    for (int j = 0; j < SHDR_EMOR_NCOMP; ++j) wj[j] = w[item * SHDR_EMOR_NCOMP + j];
    for (int s = tid; s < k; s += CURVE_THREADS) {
      float acc = 0.0f;                   // matmul row: 11-deep dot, j ascending (:246-249)
#pragma unroll
      for (int j = 0; j < SHDR_EMOR_NCOMP; ++j) acc = fmaf(hinv[j * SHDR_EMOR_SAMPLES + s], wj[j], acc);   // hinv is [11][1024] on the device
      v[s] = __fadd_rn(g0[s], acc);       // G0 + matmul(...)
    }
  } else {
    for (int s = tid; s < k; s += CURVE_THREADS) v[s] = rf_in[item * k + s];
  }
  __syncthreads();
  float* o = out + item * k;
  if (!monotone) {
    for (int s = tid; s < k; s += CURVE_THREADS) o[s] = v[s];
    return;
  }

  // thread t owns the contiguous diffs [lo, hi)
  const int n = k - 1;
  const int seg = (n + CURVE_THREADS - 1) / CURVE_THREADS;
  const int lo = min(tid * seg, n), hi = min(lo + seg, n);

  // min over g = rf[1:] - rf[:-1]                                            (:370-373)
  float m = __int_as_float(0x7f800000);
  for (int i = lo; i < hi; ++i) m = fminf(m, __fsub_rn(v[i + 1], v[i]));
  m = warp_min(m);
  if (lane == 0) red[wid] = m;
  __syncthreads();
  if (wid == 0) {
    m = warp_min(red[lane]);
    if (lane == 0) bcast[0] = fmaxf(-m, 0.0f);   // r = relu(-min_g)            (:377)
  }
  __syncthreads();
  const float r = bcast[0];

  // sum of new_g = g + r                                                      (:380-383)
  float s = 0.0f;
  for (int i = lo; i < hi; ++i) s = __fadd_rn(s, __fadd_rn(__fsub_rn(v[i + 1], v[i]), r));
  s = warp_sum(s);
  __syncthreads();                         // red[] reuse
  if (lane == 0) red[wid] = s;
  __syncthreads();
  if (wid == 0) {
    s = warp_sum(red[lane]);
    if (lane == 0) bcast[1] = s;
  }
  __syncthreads();
  const float total = bcast[1];

  // inclusive cumsum of new_g / total, left-padded with one zero              (:383-389)
  float local = 0.0f;
  for (int i = lo; i < hi; ++i)
    local = __fadd_rn(local, __fdiv_rn(__fadd_rn(__fsub_rn(v[i + 1], v[i]), r), total));
  float incl = warp_incl_scan(local, lane);
  __syncthreads();
  if (lane == 31) red[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    float t = warp_incl_scan(red[lane], lane);
    red[lane] = t;
  }
  __syncthreads();
  float excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) excl = 0.0f;
  float acc = __fadd_rn(excl, wid > 0 ? red[wid - 1] : 0.0f);   // exclusive prefix of this thread
  if (seg == 1) {
    // k == 1024 fast path: the value is exactly the block-scan's inclusive result
    if (lo < hi) o[lo + 1] = __fadd_rn(wid > 0 ? red[wid - 1] : 0.0f, incl);
  } else {
    for (int i = lo; i < hi; ++i) {
      acc = __fadd_rn(acc, __fdiv_rn(__fadd_rn(__fsub_rn(v[i + 1], v[i]), r), total));
      o[i + 1] = acc;
    }
  }
  if (tid == 0) o[0] = 0.0f;
}

static int launch_curve(const float* w, const float* rf_in, float* out, int b, int k,
                        int monotone, cudaStream_t st, int dev) {
  const float *g0 = nullptr, *hinv = nullptr;
  if (w != nullptr) {
    int rc = emor_device_table(dev, &g0, &hinv);
    if (rc != SHDR_OK) return rc;
  }
  size_t smem = (size_t)k * sizeof(float);
  if (smem > 48 * 1024)
    SHDR_CUDA(cudaFuncSetAttribute(k_curve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_curve<<<b, CURVE_THREADS, smem, st>>>(w, rf_in, g0, hinv, out, k, monotone);
  SHDR_LAUNCH_CHECK("k_curve");
  return SHDR_OK;
}

// ------------------------------------------------------------------ apply_rf
constexpr int APPLY_THREADS = 256;
constexpr int APPLY_UNROLL = 4;

// One lookup, the reference's rounding sequence:  tf_utils.py:103 (scale), :77-78 (floor, +1),
// :82/:66 (cast + clip), :87-88 (weights), :93 (w0*v0 + w1*v1 = mul, mul, add).
// tab[i] = {rf[i], rf[min(i+1,k-1)]}: the clipped pair always is (i0c, i0c) or (i0c, i0c+1).
template <bool SMEM>
__device__ __forceinline__ float lerp_lookup(float x, const float2* tab, const float* rf,
                                             float km1, int kmax) {
  float y = __fmul_rn(km1, x);
  float y0 = floorf(y);
  float y1 = __fadd_rn(y0, 1.0f);
  int i0 = min(max(__float2int_rz(y0), 0), kmax);
  int i1 = min(max(__float2int_rz(y1), 0), kmax);
  float v0, v1;
  if (SMEM) {
    float2 p = tab[i0];
    v0 = p.x;
    v1 = (i1 == i0) ? p.x : p.y;
  } else {
    v0 = __ldg(rf + i0);
    v1 = __ldg(rf + i1);
  }
  float w0 = __fsub_rn(y1, y);
  float w1 = __fsub_rn(y, y0);
  return __fadd_rn(__fmul_rn(w0, v0), __fmul_rn(w1, v1));
}

// grid.x = items * chunks_per_item; a CTA stays inside one item, so one curve per CTA.
template <bool SMEM, bool VEC>
__global__ void __launch_bounds__(APPLY_THREADS)
k_apply_rf(const float* __restrict__ x, const float* __restrict__ rf, float* __restrict__ y,
           long long elems_per_item, int k, int chunks_per_item, long long elems_per_chunk) {
  extern __shared__ float2 tab[];
  const int tid = threadIdx.x;
  const long long item = blockIdx.x / chunks_per_item;
  const int chunk = blockIdx.x - (int)(item * chunks_per_item);
  const float* r = rf + item * k;
  const float km1 = (float)(k - 1);
  const int kmax = k - 1;
  const long long e0 = (long long)chunk * elems_per_chunk;
  const long long e1 = min(e0 + elems_per_chunk, elems_per_item);
  const float* xi = x + item * elems_per_item;
  float* yi = y + item * elems_per_item;

  if (VEC) {
    // elems_per_item % 4 == 0, chunk bounds % 4 == 0, bases 16-B aligned (checked on host)
    const float4* x4 = reinterpret_cast<const float4*>(xi);
    float4* y4 = reinterpret_cast<float4*>(yi);
    const long long v1 = e1 >> 2;
    long long v = (e0 >> 2) + tid;
    // x does not depend on the curve kernel: when this grid runs as a programmatic dependent of k_curve
    // (shdr_linearize_f32) the first batch of loads is in flight while k_curve finishes
    float4 in[APPLY_UNROLL];
#pragma unroll
    for (int u = 0; u < APPLY_UNROLL; ++u) {
      const long long vv = v + u * APPLY_THREADS;
      if (vv < v1) in[u] = ld_stream4(x4 + vv);
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");   // no-op unless launched as a programmatic dependent
    if (SMEM) {
      for (int i = tid; i < k; i += APPLY_THREADS) tab[i] = make_float2(r[i], r[min(i + 1, k - 1)]);
      __syncthreads();
    }
    for (; v < v1; v += APPLY_THREADS * APPLY_UNROLL) {
#pragma unroll
      for (int u = 0; u < APPLY_UNROLL; ++u) {
        const long long vv = v + u * APPLY_THREADS;
        if (vv < v1) {
          float4 o;
          o.x = lerp_lookup<SMEM>(in[u].x, tab, r, km1, kmax);
          o.y = lerp_lookup<SMEM>(in[u].y, tab, r, km1, kmax);
          o.z = lerp_lookup<SMEM>(in[u].z, tab, r, km1, kmax);
          o.w = lerp_lookup<SMEM>(in[u].w, tab, r, km1, kmax);
          st_stream4(y4 + vv, o);
        }
      }
      const long long vn = v + APPLY_THREADS * APPLY_UNROLL;
#pragma unroll
      for (int u = 0; u < APPLY_UNROLL; ++u) {
        const long long vv = vn + u * APPLY_THREADS;
        if (vv < v1) in[u] = ld_stream4(x4 + vv);
      }
    }
  } else {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (SMEM) {
      for (int i = tid; i < k; i += APPLY_THREADS) tab[i] = make_float2(r[i], r[min(i + 1, k - 1)]);
      __syncthreads();
    }
    for (long long e = e0 + tid; e < e1; e += APPLY_THREADS)
      yi[e] = lerp_lookup<SMEM>(__ldg(xi + e), tab, r, km1, kmax);
  }
}

// ---- the steps either side of apply_rf in the inference graph, fused:
//   c = clip_by_value(x, 0, 1)                                        test_real_refinement.py:91
//   y = apply_rf(c, rf)                                               :95
//   alpha = min(1, max(0, max_c(y) - 1 + thr) / thr) tiled x 3        :98-101
// One thread handles whole RGB pixels (4 at a time when vectorised: 12 floats = 3 x 128 bit).
__device__ __forceinline__ float alpha_of(float a, float b, float c, float thr) {
  float m = fmaxf(fmaxf(a, b), c);                          // reduce_max over the channel axis
  float t = __fadd_rn(__fsub_rn(m, 1.0f), thr);             // alpha - 1.0 + THRESHOLD
  t = __fdiv_rn(fmaxf(0.0f, t), thr);                       // maximum(0.0, .) / THRESHOLD
  return fminf(1.0f, t);
}
__device__ __forceinline__ float clip01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }

template <bool SMEM, bool VEC>
__global__ void __launch_bounds__(APPLY_THREADS)
k_apply_rf_px(const float* __restrict__ x, const float* __restrict__ rf, float* __restrict__ y,
              float* __restrict__ clipped, float* __restrict__ alpha, long long px_per_item, int k,
              int chunks_per_item, long long px_per_chunk, int clip, float thr) {
  extern __shared__ float2 tab[];
  const int tid = threadIdx.x;
  const long long item = blockIdx.x / chunks_per_item;
  const int chunk = blockIdx.x - (int)(item * chunks_per_item);
  const float* r = rf + item * k;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (SMEM) {
    for (int i = tid; i < k; i += APPLY_THREADS) tab[i] = make_float2(r[i], r[min(i + 1, k - 1)]);
    __syncthreads();
  }
  const float km1 = (float)(k - 1);
  const int kmax = k - 1;
  const long long p0 = (long long)chunk * px_per_chunk;
  const long long p1 = min(p0 + px_per_chunk, px_per_item);
  const long long base = item * px_per_item * 3;
  if (VEC) {
    const float4* x4 = reinterpret_cast<const float4*>(x + base);
    float4* y4 = reinterpret_cast<float4*>(y + base);
    float4* c4 = clipped ? reinterpret_cast<float4*>(clipped + base) : nullptr;
    float4* a4 = alpha ? reinterpret_cast<float4*>(alpha + base) : nullptr;
    for (long long g = (p0 >> 2) + tid; g < (p1 >> 2); g += APPLY_THREADS) {   // group of 4 pixels = 3 float4
      float v[12], o[12];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const float4 t = ld_stream4(x4 + g * 3 + q);
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        if (clip) v[i] = clip01(v[i]);
        o[i] = lerp_lookup<SMEM>(v[i], tab, r, km1, kmax);
      }
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        st_stream4(y4 + g * 3 + q, make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]));
        if (c4) st_stream4(c4 + g * 3 + q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
      }
      if (a4) {
        float a[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = alpha_of(o[3 * i], o[3 * i + 1], o[3 * i + 2], thr);
        st_stream4(a4 + g * 3 + 0, make_float4(a[0], a[0], a[0], a[1]));
        st_stream4(a4 + g * 3 + 1, make_float4(a[1], a[1], a[2], a[2]));
        st_stream4(a4 + g * 3 + 2, make_float4(a[2], a[3], a[3], a[3]));
      }
    }
  } else {
    for (long long px = p0 + tid; px < p1; px += APPLY_THREADS) {
      float o[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float v = __ldg(x + base + px * 3 + c);
        if (clip) v = clip01(v);
        if (clipped) clipped[base + px * 3 + c] = v;
        o[c] = lerp_lookup<SMEM>(v, tab, r, km1, kmax);
        y[base + px * 3 + c] = o[c];
      }
      if (alpha) {
        const float a = alpha_of(o[0], o[1], o[2], thr);
        alpha[base + px * 3] = a; alpha[base + px * 3 + 1] = a; alpha[base + px * 3 + 2] = a;
      }
    }
  }
}

template <bool SMEM, bool VEC>
static int launch_apply_px_t(const float* x, const float* rf, float* y, float* clipped, float* alpha, int b,
                             long long npx, int k, int clip, float thr, cudaStream_t st, int dev, bool pdl) {
  const long long quantum = 4LL * APPLY_THREADS;            // pixels: every thread one group of 4
  long long per_chunk = quantum * 4;                        // 4096 pixels = 12 K elements per CTA (see launch_apply_t)
  const long long want = (long long)sm_count(dev) * 8;
  while (per_chunk > quantum && (long long)b * ((npx + per_chunk - 1) / per_chunk) < want) per_chunk >>= 1;
  const long long chunks = (npx + per_chunk - 1) / per_chunk;
  const long long grid = (long long)b * chunks;
  SHDR_REQUIRE(grid > 0 && grid <= 0x7fffffffLL, "apply_rf_ex: grid of %lld CTAs is out of range", grid);
  const size_t smem = SMEM ? (size_t)k * sizeof(float2) : 0;
  if (smem > 48 * 1024)
    SHDR_CUDA(cudaFuncSetAttribute(k_apply_rf_px<SMEM, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(APPLY_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  SHDR_CUDA(cudaLaunchKernelEx(&cfg, k_apply_rf_px<SMEM, VEC>, x, rf, y, clipped, alpha, npx, k, (int)chunks, per_chunk,
                               clip, thr));
  SHDR_LAUNCH_CHECK("k_apply_rf_px");
  return SHDR_OK;
}

static int launch_apply_px(const float* x, const float* rf, float* y, float* clipped, float* alpha, int b,
                           long long npx, int k, int clip, float thr, cudaStream_t st, int dev, bool pdl = false) {
  const bool vec = (npx % 4 == 0) && aligned16(x) && aligned16(y) && (!clipped || aligned16(clipped)) &&
                   (!alpha || aligned16(alpha));
  const bool smem = (size_t)k * sizeof(float2) <= 200 * 1024;
  if (smem) return vec ? launch_apply_px_t<true, true>(x, rf, y, clipped, alpha, b, npx, k, clip, thr, st, dev, pdl)
                       : launch_apply_px_t<true, false>(x, rf, y, clipped, alpha, b, npx, k, clip, thr, st, dev, pdl);
  return vec ? launch_apply_px_t<false, true>(x, rf, y, clipped, alpha, b, npx, k, clip, thr, st, dev, pdl)
             : launch_apply_px_t<false, false>(x, rf, y, clipped, alpha, b, npx, k, clip, thr, st, dev, pdl);
}

template <bool SMEM, bool VEC>
static int launch_apply_t(const float* x, const float* rf, float* y, int b, long long n, int k,
                          cudaStream_t st, int dev, bool pdl) {
  // chunk: multiple of 4 * threads * unroll elements; shrink until the grid has >= 8 CTAs per SM
  const long long quantum = 4LL * APPLY_THREADS * APPLY_UNROLL;   // 4096 elements
  // 8192 elements per CTA: several waves of short CTAs.  The grid of 32 K-element chunks was 1.7 waves and its tail
  // cost 12 % (A/B on config 3: 32 K 0.0770 ms, 16 K 0.0731, 8 K 0.0685, 4 K 0.0712).
  long long per_chunk = quantum * 2;
  const long long want = (long long)sm_count(dev) * 8;
  while (per_chunk > quantum && (long long)b * ((n + per_chunk - 1) / per_chunk) < want) per_chunk >>= 1;
  long long chunks = (n + per_chunk - 1) / per_chunk;
  long long grid = (long long)b * chunks;
  SHDR_REQUIRE(grid > 0 && grid <= 0x7fffffffLL, "apply_rf: grid of %lld CTAs is out of range", grid);
  size_t smem = SMEM ? (size_t)k * sizeof(float2) : 0;
  if (smem > 48 * 1024)
    SHDR_CUDA(cudaFuncSetAttribute(k_apply_rf<SMEM, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(APPLY_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;      // overlap this launch with the tail of the curve kernel before it
  SHDR_CUDA(cudaLaunchKernelEx(&cfg, k_apply_rf<SMEM, VEC>, x, rf, y, n, k, (int)chunks, per_chunk));
  SHDR_LAUNCH_CHECK("k_apply_rf");
  return SHDR_OK;
}

static int launch_apply(const float* x, const float* rf, float* y, int b, long long n, int k,
                        cudaStream_t st, int dev, bool pdl = false) {
  const bool vec = (n % 4 == 0) && aligned16(x) && aligned16(y);
  const bool smem = (size_t)k * sizeof(float2) <= 200 * 1024;
  if (smem) return vec ? launch_apply_t<true, true>(x, rf, y, b, n, k, st, dev, pdl)
                       : launch_apply_t<true, false>(x, rf, y, b, n, k, st, dev, pdl);
  return vec ? launch_apply_t<false, true>(x, rf, y, b, n, k, st, dev, pdl)
             : launch_apply_t<false, false>(x, rf, y, b, n, k, st, dev, pdl);
}


// ---- synthetic-LDR generator, the per-pixel part of _preprocessing (train.py:28-50, joint_training.py:26-46), fused:
//   x  = hdr * t[b]                                   exposure                          (:31)
//   x  = relu(x + n_s * (sigma_s[b,c] * x) + sigma_c[b,c] * n_c)   Poisson + Gaussian noise (:34-42)
//   c  = clip_by_value(x, 0, 1)                       dynamic-range clipping            (:45)
//   l  = apply_rf(c, crf[b])                          camera response                   (:48)
//   q  = round(l * 255)                               quantisation (tf.round: half to even) (:51)
// The unit-normal samples n_s, n_c are INPUTS (TensorFlow's random generator stays in TensorFlow); sigma_s / sigma_c
// are the reference's per-image, per-channel factors [b,3].  Outputs (each nullable): x (_hdr_t), c, l, q as float.
template <bool SMEM>
__global__ void __launch_bounds__(APPLY_THREADS)
k_synth_ldr(const float* __restrict__ hdr, const float* __restrict__ t, const float* __restrict__ sig_s,
            const float* __restrict__ sig_c, const float* __restrict__ n_s, const float* __restrict__ n_c,
            const float* __restrict__ rf, float* __restrict__ o_x, float* __restrict__ o_c, float* __restrict__ o_l,
            float* __restrict__ o_q, long long px_per_item, int k, int chunks_per_item, long long px_per_chunk) {
  extern __shared__ float2 tab[];
  const int tid = threadIdx.x;
  const long long item = blockIdx.x / chunks_per_item;
  const int chunk = blockIdx.x - (int)(item * chunks_per_item);
  const float* r = rf + item * k;
  if (SMEM) {
    for (int i = tid; i < k; i += APPLY_THREADS) tab[i] = make_float2(r[i], r[min(i + 1, k - 1)]);
    __syncthreads();
  }
  const float km1 = (float)(k - 1);
  const int kmax = k - 1;
  const float ti = t[item];
  float ss[3], sc[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) { ss[c] = sig_s[item * 3 + c]; sc[c] = sig_c[item * 3 + c]; }
  const long long p0 = (long long)chunk * px_per_chunk;
  const long long p1 = min(p0 + px_per_chunk, px_per_item);
  const long long base = item * px_per_item * 3;
  for (long long e = p0 * 3 + tid; e < p1 * 3; e += APPLY_THREADS) {
    const int c = (int)(e % 3);
    const float x0 = __fmul_rn(__ldg(hdr + base + e), ti);                              // hdr * t
    const float ns = __fmul_rn(__ldg(n_s + base + e), __fmul_rn(ss[c], x0));            // normal * (sigma_s * _hdr_t)
    float x = __fadd_rn(x0, ns);
    x = __fadd_rn(x, __fmul_rn(sc[c], __ldg(n_c + base + e)));                          // + sigma_c * normal
    x = fmaxf(x, 0.0f);                                                                 // relu
    const float cl = clip01(x);
    const float l = lerp_lookup<SMEM>(cl, tab, r, km1, kmax);
    if (o_x) o_x[base + e] = x;
    if (o_c) o_c[base + e] = cl;
    if (o_l) o_l[base + e] = l;
    if (o_q) o_q[base + e] = rintf(__fmul_rn(l, 255.0f));                               // round half to even
  }
}

}  // namespace shdr

using namespace shdr;

extern "C" int shdr_invcrf_build_f32(const float* w, float* curve, int b, int monotone, void* stream) {
  SHDR_REQUIRE(w && curve, "invcrf_build: NULL pointer");
  SHDR_REQUIRE(b >= 0, "invcrf_build: b=%d", b);
  if (b == 0) return SHDR_OK;
  DeviceGuard g(curve);
  if (g.status != SHDR_OK) return g.status;
  return launch_curve(w, nullptr, curve, b, SHDR_EMOR_SAMPLES, monotone, (cudaStream_t)stream, g.dev);
}

extern "C" int shdr_increase_f32(const float* rf, float* out, int b, int k, void* stream) {
  SHDR_REQUIRE(rf && out, "increase: NULL pointer");
  SHDR_REQUIRE(b >= 0 && k >= 2 && k <= 49152, "increase: b=%d k=%d (need b>=0, 2<=k<=49152)", b, k);
  if (b == 0) return SHDR_OK;
  DeviceGuard g(out);
  if (g.status != SHDR_OK) return g.status;
  return launch_curve(nullptr, rf, out, b, k, 1, (cudaStream_t)stream, g.dev);
}

extern "C" int shdr_apply_rf_f32(const float* x, const float* rf, float* y, int b,
                                 long long elems_per_item, int k, void* stream) {
  SHDR_REQUIRE(b >= 0 && elems_per_item >= 0, "apply_rf: b=%d elems_per_item=%lld", b, elems_per_item);
  SHDR_REQUIRE(k >= 1, "apply_rf: k=%d", k);
  if (b == 0 || elems_per_item == 0) return SHDR_OK;
  SHDR_REQUIRE(x && rf && y, "apply_rf: NULL pointer");
  DeviceGuard g(y);
  if (g.status != SHDR_OK) return g.status;
  return launch_apply(x, rf, y, b, elems_per_item, k, (cudaStream_t)stream, g.dev);
}

extern "C" int shdr_linearize_f32(const float* x, const float* w, float* y, float* curve_out,
                                  int b, long long elems_per_item, void* stream) {
  SHDR_REQUIRE(b >= 0 && elems_per_item >= 0, "linearize: b=%d elems_per_item=%lld", b, elems_per_item);
  if (b == 0) return SHDR_OK;
  SHDR_REQUIRE(x && w && y && curve_out, "linearize: NULL pointer");
  DeviceGuard g(y);
  if (g.status != SHDR_OK) return g.status;
  int rc = launch_curve(w, nullptr, curve_out, b, SHDR_EMOR_SAMPLES, 1, (cudaStream_t)stream, g.dev);
  if (rc != SHDR_OK || elems_per_item == 0) return rc;
  return launch_apply(x, curve_out, y, b, elems_per_item, SHDR_EMOR_SAMPLES, (cudaStream_t)stream, g.dev, true);
}

extern "C" int shdr_apply_rf_ex_f32(const float* x, const float* rf, float* y, float* clipped_out, float* alpha_out,
                                    int b, long long pixels_per_item, int k, int clip, float thr, void* stream) {
  SHDR_REQUIRE(b >= 0 && pixels_per_item >= 0, "apply_rf_ex: b=%d pixels_per_item=%lld", b, pixels_per_item);
  SHDR_REQUIRE(k >= 1, "apply_rf_ex: k=%d", k);
  SHDR_REQUIRE(!alpha_out || thr > 0.0f, "apply_rf_ex: the alpha mask needs thr > 0 (got %g)", (double)thr);
  if (b == 0 || pixels_per_item == 0) return SHDR_OK;
  SHDR_REQUIRE(x && rf && y, "apply_rf_ex: NULL pointer");
  DeviceGuard g(y);
  if (g.status != SHDR_OK) return g.status;
  return launch_apply_px(x, rf, y, clipped_out, alpha_out, b, pixels_per_item, k, clip, thr, (cudaStream_t)stream, g.dev);
}

extern "C" int shdr_linearize_ex_f32(const float* x, const float* w, float* y, float* curve_out, float* clipped_out,
                                     float* alpha_out, int b, long long pixels_per_item, int clip, float thr,
                                     void* stream) {
  SHDR_REQUIRE(b >= 0 && pixels_per_item >= 0, "linearize_ex: b=%d pixels_per_item=%lld", b, pixels_per_item);
  SHDR_REQUIRE(!alpha_out || thr > 0.0f, "linearize_ex: the alpha mask needs thr > 0 (got %g)", (double)thr);
  if (b == 0) return SHDR_OK;
  SHDR_REQUIRE(x && w && y && curve_out, "linearize_ex: NULL pointer");
  DeviceGuard g(y);
  if (g.status != SHDR_OK) return g.status;
  int rc = launch_curve(w, nullptr, curve_out, b, SHDR_EMOR_SAMPLES, 1, (cudaStream_t)stream, g.dev);
  if (rc != SHDR_OK || pixels_per_item == 0) return rc;
  return launch_apply_px(x, curve_out, y, clipped_out, alpha_out, b, pixels_per_item, SHDR_EMOR_SAMPLES, clip, thr,
                         (cudaStream_t)stream, g.dev, true);
}

extern "C" int shdr_synth_ldr_f32(const float* hdr, const float* t, const float* sigma_s, const float* sigma_c,
                                  const float* noise_s, const float* noise_c, const float* crf, float* out_hdr_t,
                                  float* out_clipped, float* out_ldr, float* out_quant, int b,
                                  long long pixels_per_item, int k, void* stream) {
  SHDR_REQUIRE(b >= 0 && pixels_per_item >= 0, "synth_ldr: b=%d pixels_per_item=%lld", b, pixels_per_item);
  SHDR_REQUIRE(k >= 1, "synth_ldr: k=%d", k);
  SHDR_REQUIRE(out_hdr_t || out_clipped || out_ldr || out_quant, "synth_ldr: every output is NULL");
  if (b == 0 || pixels_per_item == 0) return SHDR_OK;
  SHDR_REQUIRE(hdr && t && sigma_s && sigma_c && noise_s && noise_c && crf, "synth_ldr: NULL input");
  float* any = out_ldr ? out_ldr : (out_quant ? out_quant : (out_clipped ? out_clipped : out_hdr_t));
  DeviceGuard g(any);
  if (g.status != SHDR_OK) return g.status;
  const long long quantum = 4LL * APPLY_THREADS;
  long long per_chunk = quantum * 4;
  const long long want = (long long)sm_count(g.dev) * 8;
  while (per_chunk > quantum && (long long)b * ((pixels_per_item + per_chunk - 1) / per_chunk) < want) per_chunk >>= 1;
  const long long chunks = (pixels_per_item + per_chunk - 1) / per_chunk;
  const long long grid = (long long)b * chunks;
  SHDR_REQUIRE(grid > 0 && grid <= 0x7fffffffLL, "synth_ldr: grid of %lld CTAs is out of range", grid);
  const bool smem = (size_t)k * sizeof(float2) <= 200 * 1024;
  const size_t bytes = smem ? (size_t)k * sizeof(float2) : 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (smem) {
    if (bytes > 48 * 1024)
      SHDR_CUDA(cudaFuncSetAttribute(k_synth_ldr<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    k_synth_ldr<true><<<(unsigned)grid, APPLY_THREADS, bytes, st>>>(hdr, t, sigma_s, sigma_c, noise_s, noise_c, crf,
                                                                   out_hdr_t, out_clipped, out_ldr, out_quant,
                                                                   pixels_per_item, k, (int)chunks, per_chunk);
  } else {
    k_synth_ldr<false><<<(unsigned)grid, APPLY_THREADS, 0, st>>>(hdr, t, sigma_s, sigma_c, noise_s, noise_c, crf,
                                                                out_hdr_t, out_clipped, out_ldr, out_quant,
                                                                pixels_per_item, k, (int)chunks, per_chunk);
  }
  SHDR_LAUNCH_CHECK("k_synth_ldr");
  return SHDR_OK;
}
