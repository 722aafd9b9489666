// pooled.cu -- soft histogram fused with the 16x16 / stride 1 / 'same' average pool.
//
// Reference behaviour restated (ShinYwings/SingleHDR-tf2):
//   model.histogram_layer                          linearization_net.py:336-350
//   average_pooling2d(h, 16, 1, 'same') (optional) linearization_net.py:351, README.md:51
//   TF SAME semantics: window rows [y-7, y+8], cols [x-7, x+8] clipped to the image, average over
//   the in-bounds elements only (81 at the top-left corner, 64 at bottom-right, 256 inside).
//
// Design: the 3B-channel un-pooled histogram NEVER goes to HBM.  A CTA owns a 16x64 output tile,
// stages the (16+15)x(64+15)x3 input tile in shared memory once, and for every group of 12 output
// channels (4 bins x RGB) evaluates the votes on the fly and box-filters them separably:
//   pass 1 (vertical)   one thread per (column, channel): 31 votes in registers -> 16 window sums
//   pass 2 (horizontal) one thread per (row, 16-column block, channel): 31 column sums -> 16 outputs
// Window sums use the van Herk / Gil-Werman split (suffix sums of one 16-block + prefix sums of the
// next): ~2.8 adds per output, only ADDITIONS of non-negative votes -- no running-sum subtraction,
// so no cancellation and an exactly-zero window stays exactly zero (the 1e-5 RELATIVE gate).
// Out-of-image taps hold a sentinel whose vote is 0 for every bin, so the border needs no branches
// in the sums; the divide uses the true in-bounds count.
//
// Roofline: 12 B/px read + 12*B B/px written (348 B/px for B = 4, 8, 16) -> HBM-bound by intent;
// the instruction/shared-memory budget per pixel is what the kernel has to fit under.
#include "common.cuh"

namespace shdr {

constexpr int PK = 16;                 // pool window
constexpr int PB = (PK - 1) / 2;       // 7 taps before   [TF-sem] SAME: (k-1)//2 before, rest after
constexpr int PT_H = 16, PT_W = 64;    // output tile
constexpr int IN_H = PT_H + PK - 1;    // 31
constexpr int IN_W = PT_W + PK - 1;    // 79
constexpr int CG = 12;                 // channels per group: 4 bins x RGB
constexpr int VPX = 13;                // odd per-pixel stride of the column-sum buffer
constexpr int VROW = 1036;             // >= IN_W*VPX (1027) and == 12 (mod 32): pass-2 lanes hit 32 banks
constexpr int POOL_THREADS = 256;
constexpr int MAX_GROUPS = 16;
constexpr float SENTINEL = -8.0f;      // |(-8) - centre| >= 8 > 1/B  ->  vote 0 for every bin

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

// window sums of 16 over 31 values held in registers, in place:
//   a[0..15]  <- suffix sums of block A,  a[16..30] <- prefix sums of block B
//   result r  =  a[0] (r = 0)  |  a[r] + a[15 + r] (r = 1..15)
__device__ __forceinline__ void vanherk31(float (&a)[IN_H]) {
#pragma unroll
  for (int i = 14; i >= 0; --i) a[i] = __fadd_rn(a[i], a[i + 1]);
#pragma unroll
  for (int i = 17; i < 31; ++i) a[i] = __fadd_rn(a[i], a[i - 1]);
}

__global__ void __launch_bounds__(POOL_THREADS)
k_hist_pooled(const float* __restrict__ img, float* __restrict__ out, int h, int w, int ostride,
              const __grid_constant__ PoolParams prm) {
  extern __shared__ float smem[];
  float* sI = smem;                        // [IN_H][IN_W][3]
  float* sV = smem + IN_H * IN_W * 3;      // [PT_H][VROW]   (pixel stride VPX, channel fastest)
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * PT_W, y0 = blockIdx.y * PT_H;
  const long long n = blockIdx.z;
  const float* im = img + n * h * w * 3;

  // stage the input tile with halo; out-of-image taps get the sentinel
  for (int i = tid; i < IN_H * IN_W * 3; i += POOL_THREADS) {
    const int r = i / (IN_W * 3);
    const int rem = i - r * (IN_W * 3);
    const int gy = y0 - PB + r;
    const int gx = x0 - PB + rem / 3;
    float v = SENTINEL;
    if (gy >= 0 && gy < h && gx >= 0 && gx < w) v = __ldg(im + ((long long)gy * w + (x0 - PB)) * 3 + rem);
    sI[i] = v;
  }
  __syncthreads();

  for (int gi = 0; gi < prm.ngroups; ++gi) {
    const PoolGroup g = prm.g[gi];
    // ---- pass 1: vertical window sums, one (column, channel) per thread
    const int items1 = IN_W * g.nch;
    for (int it = tid; it < items1; it += POOL_THREADS) {
      const int cb = it / IN_W;            // channel within group = bin*3 + c
      const int xc = it - cb * IN_W;
      const int bin = cb / 3;
      const int c = cb - bin * 3;
      const float centre = __fdiv_rn((float)(2 * (g.bin0 + bin) + 1), 2.0f * g.nbins);
      float a[IN_H];
#pragma unroll
      for (int r = 0; r < IN_H; ++r) a[r] = hist_vote(sI[(r * IN_W + xc) * 3 + c], centre, g.thr, g.nbins);
      vanherk31(a);
      float* vcol = sV + xc * VPX + cb;
      vcol[0] = a[0];
#pragma unroll
      for (int r = 1; r < PT_H; ++r) vcol[r * VROW] = __fadd_rn(a[r], a[15 + r]);
    }
    __syncthreads();
    // ---- pass 2: horizontal window sums, divide by the in-bounds count, store
    const int items2 = g.nch * PT_H * (PT_W / 16);
    for (int it = tid; it < items2; it += POOL_THREADS) {
      const int ch = it % g.nch;
      const int t = it / g.nch;
      const int r = t % PT_H;
      const int xb = t / PT_H;
      const int gy = y0 + r;
      const int gx0 = x0 + xb * 16;
      if (gy >= h || gx0 >= w) continue;
      const float* vrow = sV + r * VROW + xb * 16 * VPX + ch;
      float a[IN_H];
#pragma unroll
      for (int j = 0; j < IN_H; ++j) a[j] = vrow[j * VPX];
      vanherk31(a);
      const int cy = min(gy + (PK - 1 - PB), h - 1) - max(gy - PB, 0) + 1;
      float* o = out + ((n * h + gy) * w + gx0) * ostride + g.out_off + ch;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int gx = gx0 + j;
        if (gx < w) {
          const int cx = min(gx + (PK - 1 - PB), w - 1) - max(gx - PB, 0) + 1;
          const float sum = (j == 0) ? a[0] : __fadd_rn(a[j], a[15 + j]);
          o[(long long)j * ostride] = __fdiv_rn(sum, (float)(cy * cx));
        }
      }
    }
    __syncthreads();
  }
}

int launch_hist_pooled(const float* img, float* out, int n, int h, int w, const int* bins, int nbins,
                       int ostride, int ooff, cudaStream_t st) {
  static const size_t smem = (size_t)(IN_H * IN_W * 3 + PT_H * VROW) * sizeof(float);
  SHDR_CUDA(cudaFuncSetAttribute(k_hist_pooled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((w + PT_W - 1) / PT_W, (h + PT_H - 1) / PT_H, n);
  SHDR_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "hist_pooled: grid (%u,%u,%u) out of range", grid.x, grid.y, grid.z);

  PoolParams prm;
  prm.ngroups = 0;
  int off = ooff;
  auto flush = [&]() -> int {
    if (prm.ngroups == 0) return SHDR_OK;
    k_hist_pooled<<<grid, POOL_THREADS, smem, st>>>(img, out, h, w, ostride, prm);
    SHDR_LAUNCH_CHECK("k_hist_pooled");
    prm.ngroups = 0;
    return SHDR_OK;
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
