// frontend.cu -- Linearization-Net feature front end, un-pooled (as the reference ships it):
// Sobel edges + spatial-aware soft histograms + the 93-channel concat, one pass over the image.
//
// Reference behaviour restated (ShinYwings/SingleHDR-tf2):
//   tf.image.sobel_edges + reshape     linearization_net.py:312-314
//   model.histogram_layer              linearization_net.py:336-350
//   tf.concat([img, edge, h4, h8, h16]) linearization_net.py:322
//
// Layout / roofline: NHWC fp32 in (12 B/px), NHWC fp32 out (372 B/px for the 93-channel tensor):
// a write-dominated stream -> HBM-bound.  A pixel's 93 floats are 372 B, which is not a multiple
// of 16 B, so no per-pixel vector store exists.  The strip kernel therefore builds the FINAL
// interleaved layout for 128 consecutive pixels in shared memory (one thread per pixel writes its
// channels at an odd word stride -> bank-conflict free) and then streams the 47.6 KB strip to HBM
// as flat, fully coalesced 128-bit stores.  Strips are taken over the flattened pixel index of the
// whole batch, so strip bases are 16-B aligned for any image width.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace shdr {

constexpr int STRIP = 128;   // pixels (= threads) per CTA

// HMASK: bit0 -> B=4, bit1 -> B=8, bit2 -> B=16.  FULL adds img(3)+edge(6) in front.
// EDGES: the six Sobel channels alone (stand-alone tf.image.sobel_edges + reshape); needs FULL and HMASK == 0.
template <bool FULL, int HMASK, bool EDGES = false>
struct StripLayout {
  static constexpr int HB = EDGES ? 6 : (FULL ? 9 : 0);
  static constexpr int CH = HB + ((HMASK & 1) ? 12 : 0) + ((HMASK & 2) ? 24 : 0) + ((HMASK & 4) ? 48 : 0);
  // odd per-pixel stride in shared memory (EDGES: 6, a 2-way conflict on six stores, so that the strip leaves as 128-bit words)
  static constexpr int CHP = ((CH & 1) || EDGES) ? CH : CH + 1;
};

template <int B>
__device__ __forceinline__ void hist_bins_pow2(float v, float* o /* stride 3 between bins */) {
#pragma unroll
  for (int b = 0; b < B; ++b) {
    const float centre = (float)(2 * b + 1) / (float)(2 * B);   // exact for power-of-two B
    o[b * 3] = hist_vote_pow2(v, centre, (float)B);
  }
}

// BF16: the strip is rounded to bfloat16 (round to nearest even) while it is streamed out -- the reduced-precision
// output flag of SURVEY.md 8(f) rank 2 (halves the 372 B/px write that dominates this kernel's roofline time; changes
// numerics, so it is a separate entry point and never the parity-gated default).
// OUT16: 0 = fp32 output, 1 = bfloat16, 2 = IEEE half precision (fp16)
template <bool FULL, int HMASK, int OUT16 = 0, bool EDGES = false>
__global__ void __launch_bounds__(STRIP)
k_frontend_strip(const float* __restrict__ img, float* __restrict__ out, int npx, int h, int w, int vec_ok) {
  using L = StripLayout<FULL, HMASK, EDGES>;
  constexpr int CH = L::CH, CHP = L::CHP;
  __shared__ __align__(16) float s[STRIP * CHP];
  const int tid = threadIdx.x;
  const int p0 = blockIdx.x * STRIP;
  const int p = p0 + tid;

  if (p < npx) {
    const float* base = img + (size_t)p * 3;
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = __ldg(base + c);
    float* o = s + tid * CHP;

    if (FULL) {
      const int row = p / w;          // n*h + y
      const int x = p - row * w;
      const int y = row % h;
      // neighbour offsets in floats, REFLECT at the true image border (-1 -> 1, n -> n-2)
      const long long oym = (long long)(reflect1(y - 1, h) - y) * w * 3;
      const long long oyp = (long long)(reflect1(y + 1, h) - y) * w * 3;
      const int oxm = (reflect1(x - 1, w) - x) * 3;
      const int oxp = (reflect1(x + 1, w) - x) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* q = base + c;
        const float p00 = __ldg(q + oym + oxm), p01 = __ldg(q + oym), p02 = __ldg(q + oym + oxp);
        const float p10 = __ldg(q + oxm), p12 = __ldg(q + oxp);
        const float p20 = __ldg(q + oyp + oxm), p21 = __ldg(q + oyp), p22 = __ldg(q + oyp + oxp);
        // cross-correlation, taps accumulated in row-major order from 0 (products by 1, 2 are exact)
        float dy = -p00;
        dy = __fadd_rn(dy, -2.0f * p01);
        dy = __fsub_rn(dy, p02);
        dy = __fadd_rn(dy, p20);
        dy = __fadd_rn(dy, 2.0f * p21);
        dy = __fadd_rn(dy, p22);
        float dx = -p00;
        dx = __fadd_rn(dx, p02);
        dx = __fadd_rn(dx, -2.0f * p10);
        dx = __fadd_rn(dx, 2.0f * p12);
        dx = __fsub_rn(dx, p20);
        dx = __fadd_rn(dx, p22);
        if (!EDGES) o[c] = v[c];
        o[(EDGES ? 0 : 3) + c * 2 + 0] = dy;
        o[(EDGES ? 0 : 3) + c * 2 + 1] = dx;
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      int off = L::HB + c;
      if (HMASK & 1) { hist_bins_pow2<4>(v[c], o + off);  off += 12; }
      if (HMASK & 2) { hist_bins_pow2<8>(v[c], o + off);  off += 24; }
      if (HMASK & 4) { hist_bins_pow2<16>(v[c], o + off); }
    }
  }
  __syncthreads();

  // stream the strip out: flat over [pixels in strip] x CH
  const int npx_strip = min(STRIP, npx - p0);
  const int nfl = npx_strip * CH;
  if (OUT16) {
    // CHP == CH here (93 is odd); the strip base p0 * CH * 2 bytes is 16-byte aligned (p0 is a multiple of 128)
    unsigned short* ob = reinterpret_cast<unsigned short*>(out) + (size_t)p0 * CH;
    // consecutive lanes read consecutive 128-bit words of the strip (conflict-free) and write 64 bits each
    const int nv = vec_ok ? (nfl >> 2) : 0;               // 4 values in, 2 x bf16x2 out
    const float4* s4 = reinterpret_cast<const float4*>(s);
    uint2* o2 = reinterpret_cast<uint2*>(ob);
#pragma unroll 4
    for (int i = tid; i < nv; i += STRIP) {
      const float4 q = s4[i];
      uint2 v;
      if (OUT16 == 1) {
        __nv_bfloat162 a = __floats2bfloat162_rn(q.x, q.y), b = __floats2bfloat162_rn(q.z, q.w);
        v.x = *reinterpret_cast<unsigned*>(&a); v.y = *reinterpret_cast<unsigned*>(&b);
      } else {
        __half2 a = __floats2half2_rn(q.x, q.y), b = __floats2half2_rn(q.z, q.w);
        v.x = *reinterpret_cast<unsigned*>(&a); v.y = *reinterpret_cast<unsigned*>(&b);
      }
      __stcs(o2 + i, v);
    }
    for (int i = (nv << 2) + tid; i < nfl; i += STRIP)
      ob[i] = OUT16 == 1 ? __bfloat16_as_ushort(__float2bfloat16_rn(s[i])) : __half_as_ushort(__float2half_rn(s[i]));
    return;
  }
  float* og = out + (size_t)p0 * CH;
  if (CHP == CH) {
    int done = 0;
    if (vec_ok) {
      const int nv = nfl >> 2;
      const float4* s4 = reinterpret_cast<const float4*>(s);
      float4* o4 = reinterpret_cast<float4*>(og);
#pragma unroll 4
      for (int i = tid; i < nv; i += STRIP) st_stream4(o4 + i, s4[i]);
      done = nv << 2;
    }
    for (int i = done + tid; i < nfl; i += STRIP) st_stream1(og + i, s[i]);
  } else {
    // padded shared layout: skip one word per pixel while reading; 4-B coalesced stores
#pragma unroll 4
    for (int i = tid; i < nfl; i += STRIP) {
      const int px = i / CH;
      st_stream1(og + i, s[i + px]);   // px*CHP + (i - px*CH)
    }
  }
}

template <bool FULL, int HMASK, int OUT16 = 0, bool EDGES = false>
static int launch_strip(const float* img, float* out, long long npx, int h, int w, cudaStream_t st) {
  const unsigned grid = (unsigned)((npx + STRIP - 1) / STRIP);
  k_frontend_strip<FULL, HMASK, OUT16, EDGES><<<grid, STRIP, 0, st>>>(img, out, (int)npx, h, w, aligned16(out) ? 1 : 0);
  SHDR_LAUNCH_CHECK("k_frontend_strip");
  return SHDR_OK;
}

// ---------------------------------------------------------------- generic (any c / any B / strided out)
// Sobel: one thread per (pixel, channel) -> writes {dy, dx}.
__global__ void __launch_bounds__(256)
k_sobel_generic(const float* __restrict__ img, float* __restrict__ out, long long total /* npx*c */,
                int h, int w, int c, int ostride, int ooff) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long p = e / c;
    const int ch = (int)(e - p * c);
    const long long row = p / w;
    const int x = (int)(p - row * w);
    const int y = (int)(row % h);
    const long long oym = (long long)(reflect1(y - 1, h) - y) * w * c;
    const long long oyp = (long long)(reflect1(y + 1, h) - y) * w * c;
    const int oxm = (reflect1(x - 1, w) - x) * c;
    const int oxp = (reflect1(x + 1, w) - x) * c;
    const float* q = img + e;
    const float p00 = __ldg(q + oym + oxm), p01 = __ldg(q + oym), p02 = __ldg(q + oym + oxp);
    const float p10 = __ldg(q + oxm), p12 = __ldg(q + oxp);
    const float p20 = __ldg(q + oyp + oxm), p21 = __ldg(q + oyp), p22 = __ldg(q + oyp + oxp);
    float dy = -p00;
    dy = __fadd_rn(dy, -2.0f * p01);
    dy = __fsub_rn(dy, p02);
    dy = __fadd_rn(dy, p20);
    dy = __fadd_rn(dy, 2.0f * p21);
    dy = __fadd_rn(dy, p22);
    float dx = -p00;
    dx = __fadd_rn(dx, p02);
    dx = __fadd_rn(dx, -2.0f * p10);
    dx = __fadd_rn(dx, 2.0f * p12);
    dx = __fsub_rn(dx, p20);
    dx = __fadd_rn(dx, p22);
    float* o = out + p * ostride + ooff + ch * 2;
    o[0] = dy;
    o[1] = dx;
  }
}

// Soft histogram: one thread per output element; runtime B and c; bit-exact generic vote.
__global__ void __launch_bounds__(256)
k_hist_generic(const float* __restrict__ img, float* __restrict__ out, long long total /* npx*c*B */,
               int c, int bins, float thr, int ostride, int ooff) {
  const int cb = c * bins;
  const float nb = (float)bins, two_b = (float)(2 * bins);
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long p = e / cb;
    const int ch = (int)(e - p * cb);
    const int b = ch / c;
    const int cc = ch - b * c;
    const float centre = __fdiv_rn((float)(2 * b + 1), two_b);      // tf.divide(2i-1, 2B) in fp32
    out[p * ostride + ooff + ch] = hist_vote(__ldg(img + p * c + cc), centre, thr, nb);
  }
}

// copy img[npx, c] into a channel slice of a wider tensor
__global__ void __launch_bounds__(256)
k_copy_strided(const float* __restrict__ img, float* __restrict__ out, long long total, int c,
               int ostride, int ooff) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long p = e / c;
    out[p * ostride + ooff + (int)(e - p * c)] = __ldg(img + e);
  }
}

static unsigned generic_grid(long long total, int dev) {
  long long blocks = (total + 255) / 256;
  long long cap = (long long)sm_count(dev) * 32;
  return (unsigned)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

int launch_sobel_generic(const float* img, float* out, long long npx, int h, int w, int c, int ostride,
                         int ooff, cudaStream_t st, int dev) {
  const long long total = npx * c;
  k_sobel_generic<<<generic_grid(total, dev), 256, 0, st>>>(img, out, total, h, w, c, ostride, ooff);
  SHDR_LAUNCH_CHECK("k_sobel_generic");
  return SHDR_OK;
}

int launch_copy_strided(const float* img, float* out, long long npx, int c, int ostride, int ooff,
                        cudaStream_t st, int dev) {
  const long long total = npx * c;
  k_copy_strided<<<generic_grid(total, dev), 256, 0, st>>>(img, out, total, c, ostride, ooff);
  SHDR_LAUNCH_CHECK("k_copy_strided");
  return SHDR_OK;
}

static int launch_hist_generic(const float* img, float* out, long long npx, int c, int bins, int ostride,
                               int ooff, cudaStream_t st, int dev) {
  const long long total = npx * c * bins;
  const float thr = (float)(1.0 / (double)bins);   // python double 1./max_bin -> fp32 tensor (:339)
  k_hist_generic<<<generic_grid(total, dev), 256, 0, st>>>(img, out, total, c, bins, thr, ostride, ooff);
  SHDR_LAUNCH_CHECK("k_hist_generic");
  return SHDR_OK;
}

// pooled_slide.cu
bool pool_slide_supported(const float* out, int w, const int* bins, int nbins, bool full93);
int launch_pool_slide(const float* img, float* out, int n, int h, int w, bool full93, int dev, cudaStream_t st);

// pooled.cu
int launch_hist_pooled(const float* img, float* out, int n, int h, int w, const int* bins, int nbins,
                       int ostride, int ooff, cudaStream_t st);

static int check_image(const char* who, const float* img, const float* out, int n, int h, int w, int c) {
  SHDR_REQUIRE(n >= 0 && h >= 0 && w >= 0 && c >= 1, "%s: bad shape n=%d h=%d w=%d c=%d", who, n, h, w, c);
  if ((long long)n * h * w == 0) return 1;   // empty: nothing to do
  SHDR_REQUIRE(img && out, "%s: NULL pointer", who);
  SHDR_REQUIRE((long long)n * h * w < 0x7fffffffLL, "%s: n*h*w = %lld pixels does not fit int32", who,
               (long long)n * h * w);
  return SHDR_OK;
}

}  // namespace shdr

using namespace shdr;

extern "C" int shdr_sobel6_f32(const float* img, float* out, int n, int h, int w, int c,
                               int out_ch_stride, int out_ch_off, void* stream) {
  int rc = check_image("sobel6", img, out, n, h, w, c);
  if (rc != SHDR_OK) return rc < 0 ? rc : SHDR_OK;
  SHDR_REQUIRE(h >= 2 && w >= 2, "sobel6: REFLECT padding needs h >= 2 and w >= 2 (got %d x %d)", h, w);
  SHDR_REQUIRE(out_ch_off >= 0 && out_ch_stride >= out_ch_off + 2 * c,
               "sobel6: out_ch_stride=%d out_ch_off=%d cannot hold %d channels", out_ch_stride, out_ch_off, 2 * c);
  DeviceGuard g(out);
  if (g.status != SHDR_OK) return g.status;
  // the stand-alone RGB tensor takes the strip kernel (edges staged in shared memory, flat 128-bit stores)
  if (c == 3 && out_ch_off == 0 && out_ch_stride == 6)
    return launch_strip<true, 0, 0, true>(img, out, (long long)n * h * w, h, w, (cudaStream_t)stream);
  return launch_sobel_generic(img, out, (long long)n * h * w, h, w, c, out_ch_stride, out_ch_off,
                              (cudaStream_t)stream, g.dev);
}

extern "C" int shdr_soft_hist_f32(const float* img, float* out, int n, int h, int w, int c, int bins,
                                  int pool_k, int out_ch_stride, int out_ch_off, void* stream) {
  int rc = check_image("soft_hist", img, out, n, h, w, c);
  if (rc != SHDR_OK) return rc < 0 ? rc : SHDR_OK;
  SHDR_REQUIRE(bins >= 1 && bins <= 4096, "soft_hist: bins=%d (need 1..4096)", bins);
  SHDR_REQUIRE(pool_k == 0 || pool_k == 16, "soft_hist: pool_k=%d (need 0 or 16)", pool_k);
  SHDR_REQUIRE(out_ch_off >= 0 && (long long)out_ch_stride >= (long long)out_ch_off + (long long)c * bins,
               "soft_hist: out_ch_stride=%d out_ch_off=%d cannot hold %d channels", out_ch_stride,
               out_ch_off, c * bins);
  DeviceGuard g(out);
  if (g.status != SHDR_OK) return g.status;
  cudaStream_t st = (cudaStream_t)stream;
  const long long npx = (long long)n * h * w;
  if (pool_k == 16) {
    if (c != 3) {
      set_error("soft_hist: the pooled form is implemented for c == 3 only (got c=%d)", c);
      return SHDR_ERR_UNSUPPORTED;
    }
    return launch_hist_pooled(img, out, n, h, w, &bins, 1, out_ch_stride, out_ch_off, st);
  }
  const bool dense = (out_ch_off == 0 && out_ch_stride == c * bins);
  if (c == 3 && dense) {
    if (bins == 4) return launch_strip<false, 1>(img, out, npx, h, w, st);
    if (bins == 8) return launch_strip<false, 2>(img, out, npx, h, w, st);
    if (bins == 16) return launch_strip<false, 4>(img, out, npx, h, w, st);
  }
  return launch_hist_generic(img, out, npx, c, bins, out_ch_stride, out_ch_off, st, g.dev);
}

extern "C" int shdr_hist_multi_f32(const float* img, float* out, int n, int h, int w, int pool_k,
                                   void* stream) {
  int rc = check_image("hist_multi", img, out, n, h, w, 3);
  if (rc != SHDR_OK) return rc < 0 ? rc : SHDR_OK;
  SHDR_REQUIRE(pool_k == 0 || pool_k == 16, "hist_multi: pool_k=%d (need 0 or 16)", pool_k);
  DeviceGuard g(out);
  if (g.status != SHDR_OK) return g.status;
  cudaStream_t st = (cudaStream_t)stream;
  if (pool_k == 0) return launch_strip<false, 7>(img, out, (long long)n * h * w, h, w, st);
  const int bins[3] = {4, 8, 16};
  return launch_hist_pooled(img, out, n, h, w, bins, 3, SHDR_HIST_CH, 0, st);
}

extern "C" int shdr_frontend_f32(const float* img, float* out, int n, int h, int w, int pool_k,
                                 void* stream) {
  int rc = check_image("frontend", img, out, n, h, w, 3);
  if (rc != SHDR_OK) return rc < 0 ? rc : SHDR_OK;
  SHDR_REQUIRE(h >= 2 && w >= 2, "frontend: REFLECT padding needs h >= 2 and w >= 2 (got %d x %d)", h, w);
  SHDR_REQUIRE(pool_k == 0 || pool_k == 16, "frontend: pool_k=%d (need 0 or 16)", pool_k);
  DeviceGuard g(out);
  if (g.status != SHDR_OK) return g.status;
  cudaStream_t st = (cudaStream_t)stream;
  const long long npx = (long long)n * h * w;
  if (pool_k == 0) return launch_strip<true, 7>(img, out, npx, h, w, st);
  // pooled histograms: img + edges go to channels 0..8, pooled histograms to 9..92 -- one launch of the sliding-window
  // kernel when its per-row bulk copies can be 16-byte aligned (w % 4 == 0), else three generic launches
  const int bins[3] = {4, 8, 16};
  if (pool_slide_supported(out, w, bins, 3, true)) return launch_pool_slide(img, out, n, h, w, true, g.dev, st);
  rc = launch_copy_strided(img, out, npx, 3, SHDR_FRONTEND_CH, 0, st, g.dev);
  if (rc != SHDR_OK) return rc;
  rc = launch_sobel_generic(img, out, npx, h, w, 3, SHDR_FRONTEND_CH, 3, st, g.dev);
  if (rc != SHDR_OK) return rc;
  return launch_hist_pooled(img, out, n, h, w, bins, 3, SHDR_FRONTEND_CH, 9, st);
}

extern "C" int shdr_frontend_bf16(const float* img, void* out_bf16, int n, int h, int w, void* stream) {
  int rc = check_image("frontend_bf16", img, (const float*)out_bf16, n, h, w, 3);
  if (rc != SHDR_OK) return rc < 0 ? rc : SHDR_OK;
  SHDR_REQUIRE(h >= 2 && w >= 2, "frontend_bf16: REFLECT padding needs h >= 2 and w >= 2 (got %d x %d)", h, w);
  DeviceGuard g(out_bf16);
  if (g.status != SHDR_OK) return g.status;
  return launch_strip<true, 7, 1>(img, (float*)out_bf16, (long long)n * h * w, h, w, (cudaStream_t)stream);
}

extern "C" int shdr_frontend_f16(const float* img, void* out_f16, int n, int h, int w, void* stream) {
  int rc = check_image("frontend_f16", img, (const float*)out_f16, n, h, w, 3);
  if (rc != SHDR_OK) return rc < 0 ? rc : SHDR_OK;
  SHDR_REQUIRE(h >= 2 && w >= 2, "frontend_f16: REFLECT padding needs h >= 2 and w >= 2 (got %d x %d)", h, w);
  DeviceGuard g(out_f16);
  if (g.status != SHDR_OK) return g.status;
  return launch_strip<true, 7, 2>(img, (float*)out_f16, (long long)n * h * w, h, w, (cudaStream_t)stream);
}
