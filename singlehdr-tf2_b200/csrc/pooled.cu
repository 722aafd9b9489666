// pooled.cu -- soft histogram fused with the 16x16 / stride 1 / 'same' average pool.
//
// Reference behaviour restated (ShinYwings/SingleHDR-tf2):
//   model.histogram_layer                          linearization_net.py:336-350
//   average_pooling2d(h, 16, 1, 'same') (optional) linearization_net.py:351, README.md:51
//   TF SAME semantics: window rows [y-7, y+8], cols [x-7, x+8] clipped to the image, average over
//   the in-bounds elements only (81 at the top-left corner, 64 at bottom-right, 256 inside).
//
// Design: the 3B-channel un-pooled histogram NEVER goes to HBM.  A CTA owns a 32x64 output tile,
// stages the (32+15)x(64+15)x3 input tile in shared memory once (transposed: one contiguous
// 47-row column per (channel, x), so a thread fetches its column with 12 LDS.128), and for every
// group of 12 output channels (4 bins x RGB) evaluates the votes on the fly and box-filters them
// separably:
//   pass 1 (vertical)   one thread per (column, channel): 47 votes in registers -> 32 window sums
//   pass 2 (horizontal) one thread per (row, 16-column block, channel): 31 column sums (8 LDS.128)
//                       -> 16 outputs, scaled by 1/count, stored with the group's 12 channels of a
//                       pixel in adjacent lanes
// Window sums use the van Herk / Gil-Werman split (suffix sums of one 16-block + prefix sums of the
// next): ~2.8 adds per output, only ADDITIONS of non-negative votes -- no running-sum subtraction,
// so no cancellation and an exactly-zero window stays exactly zero (the 1e-5 RELATIVE gate).
// Out-of-image taps hold a sentinel whose vote is 0 for every bin, so the border needs no branches
// in the sums; border tiles divide by the true in-bounds count, interior tiles multiply by the exact
// 1/256.
//
// Roofline: 12 B/px read + 12*B B/px written (348 B/px for B = 4, 8, 16) -> HBM-bound by intent;
// the instruction budget per pixel (~13.5 issue clocks per pixel per SM at the roofline) is what the
// kernel has to fit under, which is why every inner loop is register-resident and vectorised.
#include "common.cuh"

namespace shdr {

constexpr int PK = 16;                 // pool window
constexpr int PB = (PK - 1) / 2;       // 7 taps before   [TF-sem] SAME: (k-1)//2 before, rest after
constexpr int PA = PK - 1 - PB;        // 8 taps after
constexpr int PT_H = 32, PT_W = 64;    // output tile
constexpr int IN_H = PT_H + PK - 1;    // 47
constexpr int IN_W = PT_W + PK - 1;    // 79
constexpr int IPITCH = 52;             // floats per transposed input column (47 + pad); 13 x 16 B -> LDS.128, odd chunk stride
constexpr int VPITCH = 84;             // floats per (row, channel) line of column sums (79 + pad); 21 x 16 B
constexpr int CG = 12;                 // channels per group: 4 bins x RGB
constexpr int POOL_THREADS = 512;
constexpr int MAX_GROUPS = 16;
constexpr float SENTINEL = -8.0f;      // |(-8) - centre| >= 8 > 1/B  ->  vote 0 for every bin
constexpr int SI_FLOATS = 3 * IN_W * IPITCH;       // 12324
constexpr int SV_FLOATS = PT_H * CG * VPITCH;      // 32256

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

template <bool POW2>
__device__ __forceinline__ float vote(float v, float centre, float thr, float nbins) {
  return POW2 ? hist_vote_pow2(v, centre, nbins) : hist_vote(v, centre, thr, nbins);
}

template <bool POW2>
__global__ void __launch_bounds__(POOL_THREADS, 1)
k_hist_pooled(const float* __restrict__ img, float* __restrict__ out, int h, int w, int ostride,
              const __grid_constant__ PoolParams prm) {
  extern __shared__ __align__(16) float smem[];
  float* sI = smem;                 // [3][IN_W][IPITCH]  transposed input tile (+halo)
  float* sV = smem + SI_FLOATS;     // [PT_H][CG][VPITCH] column sums of the current channel group
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * PT_W, y0 = blockIdx.y * PT_H;
  const long long n = blockIdx.z;
  const float* im = img + n * h * w * 3;

  // ---- stage the input tile: coalesced global reads, transposed shared writes; sentinel outside
  for (int i = tid; i < IN_H * IN_W * 3; i += POOL_THREADS) {
    const int r = i / (IN_W * 3);
    const int rem = i - r * (IN_W * 3);
    const int xc = rem / 3;
    const int c = rem - xc * 3;
    const int gy = y0 - PB + r;
    const int gx = x0 - PB + xc;
    float v = SENTINEL;
    if (gy >= 0 && gy < h && gx >= 0 && gx < w) v = __ldg(im + ((long long)gy * w + (x0 - PB)) * 3 + rem);
    sI[(c * IN_W + xc) * IPITCH + r] = v;
  }
  for (int i = tid; i < 3 * IN_W * (IPITCH - IN_H); i += POOL_THREADS) {   // pad rows 47..51
    const int col = i / (IPITCH - IN_H);
    sI[col * IPITCH + IN_H + (i - col * (IPITCH - IN_H))] = SENTINEL;
  }
  __syncthreads();

  // every window of this tile is complete and in bounds?
  const bool interior = (y0 >= PB) && (y0 + PT_H + PA <= h) && (x0 >= PB) && (x0 + PT_W + PA <= w);

  for (int gi = 0; gi < prm.ngroups; ++gi) {
    const PoolGroup g = prm.g[gi];
    // ---- pass 1: vertical 16-window sums; item = (channel-in-group, column)
    const int items1 = IN_W * g.nch;
    for (int it = tid; it < items1; it += POOL_THREADS) {
      const int cb = it / IN_W;            // channel within group = bin*3 + c
      const int xc = it - cb * IN_W;
      const int bin = cb / 3;
      const int c = cb - bin * 3;
      const float centre = __fdiv_rn((float)(2 * (g.bin0 + bin) + 1), 2.0f * g.nbins);
      const float4* col = reinterpret_cast<const float4*>(sI + (c * IN_W + xc) * IPITCH);
      float a[48];
#pragma unroll
      for (int q = 0; q < 12; ++q) {
        const float4 t = col[q];
        a[4 * q + 0] = vote<POW2>(t.x, centre, g.thr, g.nbins);
        a[4 * q + 1] = vote<POW2>(t.y, centre, g.thr, g.nbins);
        a[4 * q + 2] = vote<POW2>(t.z, centre, g.thr, g.nbins);
        a[4 * q + 3] = vote<POW2>(t.w, centre, g.thr, g.nbins);
      }
      // blocks A = a[0..15], B = a[16..31], C = a[32..46]; output row r sums input rows r..r+15
      float bs[16];                        // suffix sums of B
      bs[15] = a[31];
#pragma unroll
      for (int i = 14; i >= 0; --i) bs[i] = __fadd_rn(a[16 + i], bs[i + 1]);
#pragma unroll
      for (int i = 14; i >= 0; --i) a[i] = __fadd_rn(a[i], a[i + 1]);            // suffix of A
#pragma unroll
      for (int i = 17; i < 32; ++i) a[i] = __fadd_rn(a[i], a[i - 1]);            // prefix of B
#pragma unroll
      for (int i = 33; i < 47; ++i) a[i] = __fadd_rn(a[i], a[i - 1]);            // prefix of C
      float* vcol = sV + cb * VPITCH + xc;
      vcol[0] = a[0];
#pragma unroll
      for (int r = 1; r < 16; ++r) vcol[r * (CG * VPITCH)] = __fadd_rn(a[r], a[15 + r]);
      vcol[16 * (CG * VPITCH)] = bs[0];
#pragma unroll
      for (int r = 17; r < 32; ++r) vcol[r * (CG * VPITCH)] = __fadd_rn(bs[r - 16], a[15 + r]);
    }
    __syncthreads();

    // ---- pass 2: horizontal 16-window sums, scale, store; item = (16-col block, row, channel)
    const int lines = PT_H * g.nch;        // (row, channel) lines, channel fastest
    const int items2 = lines * (PT_W / 16);
    for (int it = tid; it < items2; it += POOL_THREADS) {
      const int xb = it / lines;
      const int line = it - xb * lines;
      const int r = line / g.nch;
      const int ch = line - r * g.nch;
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
        for (int j = 1; j < 16; ++j) o[(long long)j * ostride] = __fadd_rn(a[j], a[15 + j]) * (1.0f / 256.0f);
      } else {
        const int cy = min(gy + PA, h - 1) - max(gy - PB, 0) + 1;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int gx = gx0 + j;
          if (gx < w) {
            const int cx = min(gx + PA, w - 1) - max(gx - PB, 0) + 1;
            const float sum = (j == 0) ? a[0] : __fadd_rn(a[j], a[15 + j]);
            o[(long long)j * ostride] = __fdiv_rn(sum, (float)(cy * cx));
          }
        }
      }
    }
    __syncthreads();
  }
}

template <bool POW2>
static int launch_pooled_t(const float* img, float* out, int h, int w, int ostride, dim3 grid,
                           const PoolParams& prm, cudaStream_t st) {
  const size_t smem = (size_t)(SI_FLOATS + SV_FLOATS) * sizeof(float);
  SHDR_CUDA(cudaFuncSetAttribute(k_hist_pooled<POW2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_hist_pooled<POW2><<<grid, POOL_THREADS, smem, st>>>(img, out, h, w, ostride, prm);
  SHDR_LAUNCH_CHECK("k_hist_pooled");
  return SHDR_OK;
}

int launch_hist_pooled(const float* img, float* out, int n, int h, int w, const int* bins, int nbins,
                       int ostride, int ooff, cudaStream_t st) {
  dim3 grid((w + PT_W - 1) / PT_W, (h + PT_H - 1) / PT_H, n);
  SHDR_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "hist_pooled: grid (%u,%u,%u) out of range", grid.x, grid.y, grid.z);

  PoolParams prm;
  prm.ngroups = 0;
  bool pow2 = true;
  int off = ooff;
  auto flush = [&]() -> int {
    if (prm.ngroups == 0) return SHDR_OK;
    int rc = pow2 ? launch_pooled_t<true>(img, out, h, w, ostride, grid, prm, st)
                  : launch_pooled_t<false>(img, out, h, w, ostride, grid, prm, st);
    prm.ngroups = 0;
    pow2 = true;
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
      pow2 = pow2 && ((B & (B - 1)) == 0);
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
