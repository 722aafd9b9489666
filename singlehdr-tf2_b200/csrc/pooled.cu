// pooled.cu -- soft histogram fused with the 16x16 / stride 1 / 'same' average pool.
//
// Reference behaviour restated (ShinYwings/SingleHDR-tf2):
//   model.histogram_layer                          linearization_net.py:336-350
//   average_pooling2d(h, 16, 1, 'same') (optional) linearization_net.py:351, README.md:51
//   TF SAME semantics: window rows [y-7, y+8], cols [x-7, x+8] clipped to the image, average over
//   the in-bounds elements only (81 at the top-left corner, 64 at bottom-right, 256 inside).
//
// Design: the 3B-channel un-pooled histogram NEVER goes to HBM.  A CTA owns a 16x64 output tile
// (two CTAs per SM), stages the (16+15)x(64+15)x3 input tile in shared memory once with cp.async
// (transposed: one contiguous 31-row column per (channel, x), so a thread fetches its column with
// 8 LDS.128), and for every group of 12 output channels (4 bins x RGB) evaluates the votes on the
// fly and box-filters them separably:
//   pass 1 (vertical)   one thread per (column, channel): 31 votes in registers -> 16 window sums
//   pass 2 (horizontal) one thread per (16-column block, row, channel): 31 column sums (8 LDS.128)
//                       -> 16 outputs, scaled by 1/count, stored with the group's 12 channels of a
//                       pixel in adjacent lanes
// Window sums use the van Herk / Gil-Werman split (suffix sums of one 16-block + prefix sums of the
// next): ~2.8 adds per output, only ADDITIONS of non-negative votes -- no running-sum subtraction,
// so no cancellation and an exactly-zero window stays exactly zero (the 1e-5 RELATIVE gate).
// Out-of-image taps hold a sentinel whose vote is 0 for every bin, so the border needs no branches
// in the sums.  Interior tiles scale by the exact 1/256; border tiles by (1/rows)*(1/cols) from two
// small per-tile tables (<= 2 ulp from the reference's divide by rows*cols).
//
// Roofline: 12 B/px read + 12*B B/px written (348 B/px for B = 4, 8, 16) -> HBM-bound by intent;
// the instruction budget per pixel (~13.5 issue clocks per pixel per SM at the roofline) is what the
// kernel has to fit under, which is why every inner loop is register-resident and vectorised.
#include <stdlib.h>

#include "common.cuh"

namespace shdr {

constexpr int PK = 16;                 // pool window
constexpr int PB = (PK - 1) / 2;       // 7 taps before   [TF-sem] SAME: (k-1)//2 before, rest after
constexpr int PA = PK - 1 - PB;        // 8 taps after
constexpr int PT_H = 16, PT_W = 64;    // output tile
constexpr int IN_H = PT_H + PK - 1;    // 31
constexpr int IN_W = PT_W + PK - 1;    // 79
constexpr int IPITCH = 36;             // floats per transposed input column (31 + pad); 9 x 16 B -> LDS.128, odd chunk stride
constexpr int VPITCH = 84;             // floats per (row, channel) line of column sums (79 + pad); 21 x 16 B
constexpr int CG = 12;                 // channels per group: 4 bins x RGB
constexpr int POOL_THREADS = 512;
constexpr int MAX_GROUPS = 16;
constexpr float SENTINEL = -8.0f;      // |(-8) - centre| >= 8 > 1/B  ->  vote 0 for every bin
constexpr int SI_FLOATS = 3 * IN_W * IPITCH;       // 8532
constexpr int SV_FLOATS = PT_H * CG * VPITCH;      // 16128
constexpr int SC_FLOATS = PT_W + PT_H;             // border scale tables: 1/cols per column, 1/rows per row

struct PoolGroup {
  float nbins;     // float(B)
  float thr;       // float(1.0 / B)
  int bin0;        // first bin (0-based) of this group
  int nch;         // 3 * bins in this group (<= 12)
  int out_off;     // channel offset of the group's first channel in the output pixel
};
struct PoolParams {
  PoolGroup g[MAX_GROUPS];
  int ngroups;
};

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gmem_src) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// FAST: every B is a power of two and every group is a full 4-bin x RGB group (the B = 4, 8, 16 case).
// OSTRIDE: compile-time pixel stride of the output (84, 93) or 0 = runtime.
template <bool FAST, int OSTRIDE>
__global__ void __launch_bounds__(POOL_THREADS, 2)
k_hist_pooled(const float* __restrict__ img, float* __restrict__ out, int h, int w, int ostride_rt,
              const __grid_constant__ PoolParams prm) {
  extern __shared__ __align__(16) float smem[];
  float* sI = smem;                          // [3][IN_W][IPITCH]  transposed input tile (+halo)
  float* sV = smem + SI_FLOATS;              // [PT_H][CG][VPITCH] column sums of the current channel group
  float* sRcx = smem + SI_FLOATS + SV_FLOATS;   // [PT_W] 1 / (#in-bounds window columns)
  float* sRcy = sRcx + PT_W;                    // [PT_H] 1 / (#in-bounds window rows)
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * PT_W, y0 = blockIdx.y * PT_H;
  const long long n = blockIdx.z;
  const float* im = img + n * h * w * 3;
  const long long ostride = OSTRIDE ? OSTRIDE : ostride_rt;

  // ---- stage the input tile: coalesced global reads -> transposed shared layout; sentinel outside
  for (int i = tid; i < IN_H * IN_W * 3; i += POOL_THREADS) {
    const int r = i / (IN_W * 3);
    const int rem = i - r * (IN_W * 3);
    const int xc = rem / 3;
    const int c = rem - xc * 3;
    const int gy = y0 - PB + r;
    const int gx = x0 - PB + xc;
    float* dst = sI + (c * IN_W + xc) * IPITCH + r;
    if (gy >= 0 && gy < h && gx >= 0 && gx < w) cp_async4(dst, im + ((long long)gy * w + (x0 - PB)) * 3 + rem);
    else *dst = SENTINEL;
  }
  for (int i = tid; i < 3 * IN_W * (IPITCH - IN_H); i += POOL_THREADS) {   // pad rows 31..35
    const int col = i / (IPITCH - IN_H);
    sI[col * IPITCH + IN_H + (i - col * (IPITCH - IN_H))] = SENTINEL;
  }
  if (tid < PT_W) {
    const int gx = min(x0 + tid, w - 1);
    sRcx[tid] = __fdiv_rn(1.0f, (float)(min(gx + PA, w - 1) - max(gx - PB, 0) + 1));
  } else if (tid < PT_W + PT_H) {
    const int gy = min(y0 + tid - PT_W, h - 1);
    sRcy[tid - PT_W] = __fdiv_rn(1.0f, (float)(min(gy + PA, h - 1) - max(gy - PB, 0) + 1));
  }
  cp_async_wait_all();
  __syncthreads();

  // every window of this tile is complete (count 256) and every output pixel is in bounds?
  const bool interior = (y0 >= PB) && (y0 + PT_H + PA <= h) && (x0 >= PB) && (x0 + PT_W + PA <= w);

  for (int gi = 0; gi < prm.ngroups; ++gi) {
    const PoolGroup g = prm.g[gi];
    const int nch = FAST ? CG : g.nch;
    // ---- pass 1: vertical 16-window sums; item = (channel-in-group, column)
    const int items1 = IN_W * nch;
    for (int it = tid; it < items1; it += POOL_THREADS) {
      const int cb = it / IN_W;            // channel within group = bin*3 + c
      const int xc = it - cb * IN_W;
      const int bin = cb / 3;
      const int c = cb - bin * 3;
      const float k2 = (float)(2 * (g.bin0 + bin) + 1);
      // tf.divide(2i-1, 2B) in fp32; for a power-of-two B the product with the exact 1/(2B) is the same value
      const float centre = FAST ? k2 * (0.5f / g.nbins) : __fdiv_rn(k2, 2.0f * g.nbins);
      const float4* col = reinterpret_cast<const float4*>(sI + (c * IN_W + xc) * IPITCH);
      float a[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 t = col[q];
        if (FAST) {
          a[4 * q + 0] = hist_vote_pow2(t.x, centre, g.nbins);
          a[4 * q + 1] = hist_vote_pow2(t.y, centre, g.nbins);
          a[4 * q + 2] = hist_vote_pow2(t.z, centre, g.nbins);
          a[4 * q + 3] = hist_vote_pow2(t.w, centre, g.nbins);
        } else {
          a[4 * q + 0] = hist_vote(t.x, centre, g.thr, g.nbins);
          a[4 * q + 1] = hist_vote(t.y, centre, g.thr, g.nbins);
          a[4 * q + 2] = hist_vote(t.z, centre, g.thr, g.nbins);
          a[4 * q + 3] = hist_vote(t.w, centre, g.thr, g.nbins);
        }
      }
      // blocks A = a[0..15], B = a[16..30]; output row r sums input rows r..r+15
#pragma unroll
      for (int i = 14; i >= 0; --i) a[i] = __fadd_rn(a[i], a[i + 1]);            // suffix of A
#pragma unroll
      for (int i = 17; i < 31; ++i) a[i] = __fadd_rn(a[i], a[i - 1]);            // prefix of B
      float* vcol = sV + cb * VPITCH + xc;
      vcol[0] = a[0];
#pragma unroll
      for (int r = 1; r < 16; ++r) vcol[r * (CG * VPITCH)] = __fadd_rn(a[r], a[15 + r]);
    }
    __syncthreads();

    // ---- pass 2: horizontal 16-window sums, scale, store; item = (16-col block, row, channel)
    const int lines = PT_H * nch;          // (row, channel) lines, channel fastest
    const int items2 = lines * (PT_W / 16);
    for (int it = tid; it < items2; it += POOL_THREADS) {
      const int xb = it / lines;
      const int line = it - xb * lines;
      const int r = line / nch;
      const int ch = line - r * nch;
      const int gy = y0 + r;
      const int gx0 = x0 + xb * 16;
      if (!interior && (gy >= h || gx0 >= w)) continue;
      const float4* vl = reinterpret_cast<const float4*>(sV + (r * CG + ch) * VPITCH + xb * 16);
      float a[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 t = vl[q];
        a[4 * q + 0] = t.x; a[4 * q + 1] = t.y; a[4 * q + 2] = t.z; a[4 * q + 3] = t.w;
      }
#pragma unroll
      for (int i = 14; i >= 0; --i) a[i] = __fadd_rn(a[i], a[i + 1]);            // suffix of cols 0..15
#pragma unroll
      for (int i = 17; i < 31; ++i) a[i] = __fadd_rn(a[i], a[i - 1]);            // prefix of cols 16..30
      float* o = out + ((n * h + gy) * w + gx0) * ostride + g.out_off + ch;
      if (interior) {
        o[0] = a[0] * (1.0f / 256.0f);     // exact: power-of-two count
#pragma unroll
        for (int j = 1; j < 16; ++j) o[j * ostride] = __fadd_rn(a[j], a[15 + j]) * (1.0f / 256.0f);
      } else {
        const float rcy = sRcy[r];
        const float4* rc4 = reinterpret_cast<const float4*>(sRcx + xb * 16);
        float rc[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 t = rc4[q];
          rc[4 * q + 0] = t.x; rc[4 * q + 1] = t.y; rc[4 * q + 2] = t.z; rc[4 * q + 3] = t.w;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (gx0 + j < w) {
            const float sum = (j == 0) ? a[0] : __fadd_rn(a[j], a[15 + j]);
            o[j * ostride] = sum * (rcy * rc[j]);
          }
        }
      }
    }
    __syncthreads();
  }
}

constexpr size_t POOL_SMEM = (size_t)(SI_FLOATS + SV_FLOATS + SC_FLOATS) * sizeof(float);

template <bool FAST, int OSTRIDE>
static int launch_pooled_t(const float* img, float* out, int h, int w, int ostride, dim3 grid,
                           const PoolParams& prm, cudaStream_t st) {
  SHDR_CUDA(cudaFuncSetAttribute(k_hist_pooled<FAST, OSTRIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)POOL_SMEM));
  k_hist_pooled<FAST, OSTRIDE><<<grid, POOL_THREADS, POOL_SMEM, st>>>(img, out, h, w, ostride, prm);
  SHDR_LAUNCH_CHECK("k_hist_pooled");
  return SHDR_OK;
}

bool hist_pooled_ws_supported(int w, const int* bins, int nbins, int ostride, int ooff);
int launch_hist_pooled_ws(const float* img, float* out, int n, int h, int w, const int* bins, int nbins, int dev,
                          cudaStream_t st);
bool pool_slide_supported(const float* out, int w, const int* bins, int nbins, bool full93);
int launch_pool_slide(const float* img, float* out, int n, int h, int w, bool full93, int dev, cudaStream_t st);

int launch_hist_pooled(const float* img, float* out, int n, int h, int w, const int* bins, int nbins,
                       int ostride, int ooff, cudaStream_t st) {
  int dev = 0;
  cudaGetDevice(&dev);
  // 1. the {4, 8, 16} histograms into a dense 84-channel tensor: exact sliding-window kernel (pooled_slide.cu)
  if (ostride == SHDR_HIST_CH && ooff == 0 && pool_slide_supported(out, w, bins, nbins, false))
    return launch_pool_slide(img, out, n, h, w, false, dev, st);
  // 2. power-of-two B, dense output, even width -> warp-specialised whole-sector kernel (pooled_ws.cu);
  //    its TMA stores need 16-B alignment
  if (aligned16(out) && hist_pooled_ws_supported(w, bins, nbins, ostride, ooff))
    return launch_hist_pooled_ws(img, out, n, h, w, bins, nbins, dev, st);
  dim3 grid((w + PT_W - 1) / PT_W, (h + PT_H - 1) / PT_H, n);
  SHDR_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "hist_pooled: grid (%u,%u,%u) out of range", grid.x, grid.y, grid.z);

  PoolParams prm;
  prm.ngroups = 0;
  bool fast = true;
  int off = ooff;
  auto flush = [&]() -> int {
    if (prm.ngroups == 0) return SHDR_OK;
    int rc;
    if (fast && ostride == SHDR_HIST_CH) rc = launch_pooled_t<true, SHDR_HIST_CH>(img, out, h, w, ostride, grid, prm, st);
    else if (fast && ostride == SHDR_FRONTEND_CH) rc = launch_pooled_t<true, SHDR_FRONTEND_CH>(img, out, h, w, ostride, grid, prm, st);
    else if (fast) rc = launch_pooled_t<true, 0>(img, out, h, w, ostride, grid, prm, st);
    else rc = launch_pooled_t<false, 0>(img, out, h, w, ostride, grid, prm, st);
    prm.ngroups = 0;
    fast = true;
    return rc;
  };
  for (int i = 0; i < nbins; ++i) {
    const int B = bins[i];
    for (int b0 = 0; b0 < B; b0 += CG / 3) {
      PoolGroup& g = prm.g[prm.ngroups++];
      g.nbins = (float)B;
      g.thr = (float)(1.0 / (double)B);     // python double 1./max_bin -> fp32 tensor (:339)
      g.bin0 = b0;
      g.nch = 3 * (B - b0 < CG / 3 ? B - b0 : CG / 3);
      g.out_off = off + 3 * b0;
      fast = fast && ((B & (B - 1)) == 0) && g.nch == CG;
      if (prm.ngroups == MAX_GROUPS) {
        int rc = flush();
        if (rc != SHDR_OK) return rc;
      }
    }
    off += 3 * B;
  }
  return flush();
}

}  // namespace shdr
