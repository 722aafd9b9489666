// conv1_fused.cu -- SURVEY.md 8(f) rank 2: the feature front end fused into the INPUT of crfFeatureNet.conv1.
//
// Reference behaviour restated (ShinYwings/SingleHDR-tf2):
//   features = tf.concat([img, sobel6, hist4, hist8, hist16], -1)            linearization_net.py:312-322
//   conv1    = Conv2D(64, (7,7), strides (2,2), padding 'SAME', bias)(features)   linearization_net.py:91,107
// The 93-channel tensor (372 B per input pixel) is the front end's whole HBM cost and conv1 reduces it 4x
// spatially right away.  Here it never exists in HBM: every CTA generates the fp16 feature tile of one 16 x 8
// block of OUTPUT pixels in shared memory and contracts it against the 7x7x93x64 kernel on the tensor cores
// (tcgen05.mma, fp16 x fp16 -> fp32 accumulators in tensor memory).  HBM traffic: 12 B per input pixel in,
// 256 B per output pixel (= 64 B per input pixel) out -- the kernel is TENSOR-bound (583 kFLOP per output pixel),
// unlike everything else in this library.  It changes numerics (half-precision operands, like TF's own mixed_float16
// / TF32 conv paths; fp32 accumulation) and is therefore a separate entry point, never the parity-gated fp32 default.
//
// Implicit GEMM.  D[m, o] = sum_{ky,kx,c} F(2 oy + ky - pt, 2 ox + kx - pl, c) W[ky,kx,c,o];  m = 128 output pixels
// (16 rows x 8 columns), N = 64 output channels, K = 49 taps x 96 (93 channels + 3 zero).  For one tap the A operand
// is a SHIFTED, stride-2 VIEW of the feature tile -- no im2col copy is made.  tcgen05 reads K-major operands as
// 8-row x 16-byte core matrices addressed by (start, LBO, SBO), so the tile is stored as
//      [input row ly 0..36][16-byte channel group cg][column parity px][q = lx >> 1][8 channels]     (fp16)
// in which the 8 pixels of one output row of the tile are 8 consecutive q (16 B apart) for every tap:
//      start = buf + ky RP + cg CGP + (kx & 1) PP + (kx >> 1) 16,   LBO = CGP (next channel group),
//      SBO = 2 RP (next output row = two input rows down).
// K is split into three passes of 32 channels (channels 0-31 | 32-63 | 64-92 + 3 zeros) that alternate between two
// 51 KB feature buffers, so that the CUDA cores generate one pass while the tensor cores consume the other.
//
// Roles of the persistent CTA (one per SM, 20 warps): 13 producer warps (one thread per input pixel of the 37 x 21
// halo tile: Sobel / votes of the pass, fp16 pack, four 16-byte shared stores), 1 thread streaming the packed weights
// (one kernel row of one pass = 28 KB per 1-D bulk copy into a 4-stage ring), 2 warps issuing the MMAs (alternating
// kernel rows, one elected lane, 14 MMAs per row back to back, 294 per tile),
// 4 epilogue warps (tcgen05.ld of the two partial accumulators, 16 columns at a time, add, scale / shift / ReLU, 256 B per output pixel to
// HBM; two tiles' accumulators in tensor memory, so the epilogue of one tile overlaps the MMAs of the next).
//
// CTA pairs (the default; k_frontend_conv1_pair, clusters of 2 = the two SMs of a TPC).  Every MMA is issued with
// cta_group::2 by the even CTA: M = 256 = this CTA's tile and the peer's, each CTA reads its own A and only HALF of B
// (32 of the 64 output channels) from its own shared memory, and each accumulates its 128 rows in its own tensor
// memory.  Per SM that is 5 KB instead of 6 KB of operand reads per MMA and half the weight bytes written into shared
// memory and read from L2.  The peer has no issuer: its producers and epilogue report to the leader's barriers with
// remote arrives, one of its idle threads relays "my half of ring stage s has landed", and the leader's commits are
// multicast to the barriers of both CTAs.  A pair whose second tile index falls off the end runs a dummy tile there.
// Inputs with fewer than 4 tiles take the single-CTA kernel (same code, PAIR = false).
#include <cuda_fp16.h>

#include "common.cuh"
#include "umma.cuh"

namespace shdr {
namespace c1 {

using namespace umma;

constexpr int TR = 16, TC = 8;            // output tile (M = 128)
constexpr int LR = 2 * TR + 5;            // 37 input rows under a tile
constexpr int LC = 2 * TC + 5;            // 21 input columns
constexpr int Q = 11;                     // columns per parity plane (even: 11, odd: 10 used)
constexpr int NPASS = 3;                  // K passes per tile
constexpr int KPASS = 32;                 // channels per pass (96 = 93 + 3 zeros)
constexpr int CG = KPASS / 8;             // 16-byte channel groups per pass
constexpr int PP = Q * 16;                // 176 B   odd-column plane starts right behind the 11 even columns
constexpr int CGP = (2 * Q - 1) * 16;     // 336 B   next channel group = LBO of A (21 columns)
constexpr int RP = (CG * (2 * Q - 1) + 1) * 16;   // 1360 B input row: 85 x 16 B -- 85 = 21 (mod 8), see gen_pass
constexpr int FBUF = LR * RP;             // 50320 B per pass
constexpr int SBO_A = 2 * RP;             // 2720 B
constexpr int NTAP = 49;
constexpr int OC = 64;                    // output channels
constexpr int WTAP = KPASS * OC * 2;      // 4096 B: one tap of one pass
constexpr int WROW = 7 * WTAP;            // 28672 B: one kernel row of one pass = one ring stage
constexpr int NWS = 4;                    // weight ring depth (kernel rows)
// CTA-pair mode (cta_group::2): each CTA holds 32 of the 64 output channels of B -> half the bytes per stage, twice the depth
template <bool PAIR> struct Ring {
  static constexpr int TAP = PAIR ? WTAP / 2 : WTAP;      // bytes of one tap in a stage
  static constexpr int ROW = 7 * TAP;                     // bytes of a stage (one kernel row of one pass)
  static constexpr int DEPTH = PAIR ? 2 * NWS : NWS;      // both: 114688 B of ring
  static constexpr int KSTEP = TAP / 2;                   // bytes of one 16-channel step of a tap
  static constexpr int LBO = KSTEP / 2;                   // K half
};
constexpr int NPROD = 416;                // producer threads (warps 0..12): 777 halo pixels = 2 per thread (1.87)
constexpr int W_LOAD = 13, W_MMA = 14;    // warps 14, 15: MMA issuers; warps 16..19: epilogue
constexpr int NTHREADS = 20 * 32;
constexpr int TMEM_COLS = 256;            // two tiles in flight x two issuer warps x (128 x 64 fp32)
constexpr size_t PACKED_ONE = (size_t)NPASS * NTAP * WTAP;     // 602112: one operand image of the whole kernel
constexpr size_t PACKED_BYTES = 2 * PACKED_ONE;                // the single-CTA image, then the CTA-pair image
constexpr int RAW_R = LR + 2, RAW_C = LC + 2;  // raw fp32 image tile with the Sobel halo: 39 x 23 pixels
constexpr int RAWC = 24;                  // raw tile, per row and colour: 12 odd columns then 12 even columns (23 used)
constexpr int RAWP = 3 * RAWC + 15;       // 87 floats per row: consecutive producer lanes (11 even + 10 odd tile columns,
                                          // then the next row, 87 = 23 (mod 32) words on) read mostly distinct banks
                                          // (ncu: still ~2 wavefronts per LDS -- one colliding pair is enough)
constexpr int RAW_BYTES = ((RAW_R * RAWP * 4 + 15) / 16) * 16;   // 13584
constexpr int SMEM_BYTES = 2 * FBUF + NWS * WROW + RAW_BYTES;    // 228912 (both modes: Ring::DEPTH * Ring::ROW == NWS * WROW)

struct Params {
  const float* img;
  const unsigned char* wpk;   // packed weights [pass][tap][4096]
  const float* scale;         // [64] or null (1)
  const float* shift;         // [64] or null (0): the conv bias, or a folded batch-norm shift
  float* out;                 // [n, oh, ow, 64]
  int n, h, w, oh, ow, pt, pl, tiles_y, tiles_x, ntiles, relu;
};

__device__ __forceinline__ float sat_half(float x);

// ------------------------------------------------------------------------------------------ weight packing
// kernel [7][7][93][64] fp32 (HWIO, what Keras stores) -> for each pass and tap the B operand image tcgen05 reads:
// K-major no-swizzle core matrices, [16-channel step][K half][8-channel n group][n row][8 k] fp16 (LBO 1024, SBO 128).
// K index k of pass p is feature channel 32 p + k (zero beyond 92).
__global__ void __launch_bounds__(256) k_pack_weights(const float* __restrict__ kern, __half* __restrict__ wpk) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= NPASS * NTAP * KPASS * OC) return;
  const int e = i & 7, r = (i >> 3) & 7, ng = (i >> 6) & 7, kh = (i >> 9) & 1, step = (i >> 10) & 1;
  const int ptap = i >> 11;                    // pass * 49 + tap
  const int tap = ptap % NTAP, pass = ptap / NTAP;
  const int ch = pass * KPASS + step * 16 + kh * 8 + e, o = ng * 8 + r;
  wpk[i] = __float2half_rn(ch >= SHDR_FRONTEND_CH ? 0.0f : sat_half(__ldg(kern + ((size_t)tap * SHDR_FRONTEND_CH + ch) * OC + o)));
}
// the same for CTA pairs: [pass][kernel row][half of the output channels][kx][16-channel step][K half][4 n groups][n row][8 k]
// -- each CTA's share of a kernel row is one contiguous 14 KB block (LBO 512, SBO 128)
__global__ void __launch_bounds__(256) k_pack_weights_pair(const float* __restrict__ kern, __half* __restrict__ wpk) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= NPASS * NTAP * KPASS * OC) return;
  const int e = i & 7, r = (i >> 3) & 7, ng = (i >> 6) & 3, kh = (i >> 8) & 1, step = (i >> 9) & 1;
  int rest = i >> 10;                          // ((pass * 7 + ky) * 2 + half) * 7 + kx
  const int kx = rest % 7; rest /= 7;
  const int half = rest & 1; rest >>= 1;
  const int ky = rest % 7, pass = rest / 7;
  const int ch = pass * KPASS + step * 16 + kh * 8 + e, o = half * 32 + ng * 8 + r;
  wpk[i] = __float2half_rn(ch >= SHDR_FRONTEND_CH ? 0.0f
                           : sat_half(__ldg(kern + ((size_t)(ky * 7 + kx) * SHDR_FRONTEND_CH + ch) * OC + o)));
}

// ------------------------------------------------------------------------------------------ producers
// Operands are IEEE half precision (fp16): every feature is in [0, 1] (votes, LDR pixels) or a small multiple of it
// (Sobel: |.| <= 4), and conv weights are O(1), so fp16's 11-bit significand gives a 4x smaller rounding error than
// bf16's 8 bits at the same tensor-core rate and the same bytes (11 bits is also what TF32 -- TensorFlow's default for
// fp32 convolutions on Ampere-or-later GPUs -- keeps).  Values beyond fp16's range saturate at +-65504
// instead of becoming inf (inf x 0 would poison the sum); NaN stays NaN.
__device__ __forceinline__ float sat_half(float x) { return fabsf(x) > 65504.0f ? copysignf(65504.0f, x) : x; }
__device__ __forceinline__ unsigned pack2(float lo, float hi) {
  __half2 t = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<unsigned*>(&t);
}

// feature channel CH (compile-time after unrolling) of a pixel: the concat order of linearization_net.py:322
//   0-2 img | 3-8 Sobel (c*2 + {dy, dx}) | 9-20 hist4 | 21-44 hist8 | 45-92 hist16 ((bin-1)*3 + c) | 93-95 zero
__device__ __forceinline__ float feature(int ch, const float* v, const float* sob) {
  if (ch < 3) return sat_half(v[ch]);
  if (ch < 9) return sat_half(sob[ch - 3]);
  if (ch < 21) { const int i = ch - 9;  return hist_vote_pow2(v[i % 3], (float)(2 * (i / 3) + 1) / 8.0f, 4.0f); }
  if (ch < 45) { const int i = ch - 21; return hist_vote_pow2(v[i % 3], (float)(2 * (i / 3) + 1) / 16.0f, 8.0f); }
  if (ch < 93) { const int i = ch - 45; return hist_vote_pow2(v[i % 3], (float)(2 * (i / 3) + 1) / 32.0f, 16.0f); }
  return 0.0f;
}

// one input pixel, the 32 channels of pass PASS -> four 16-byte groups of the feature tile.  `r` points at the pixel
// in the raw tile (its 3 x 3 neighbourhood is there too, REFLECTed at the image border by the staging step).
// raw tile word of column rx (0..22), colour 0, in its row: odd raw columns (= even tile columns) first, then the even ones
__device__ __forceinline__ int raw_col(int rx) { return ((rx & 1) ^ 1) * (RAWC / 2) + (rx >> 1); }

template <int PASS>
__device__ __forceinline__ void gen_pixel(const float* __restrict__ row, int rx, unsigned char* dst) {
  const int xm = raw_col(rx - 1), x0 = raw_col(rx), xp = raw_col(rx + 1);
  float v[3], sob[6];
#pragma unroll
  for (int c = 0; c < 3; ++c) v[c] = row[c * RAWC + x0];
  if (PASS == 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* q = row + c * RAWC;
      const float p00 = q[-RAWP + xm], p01 = q[-RAWP + x0], p02 = q[-RAWP + xp];
      const float p10 = q[xm], p12 = q[xp];
      const float p20 = q[RAWP + xm], p21 = q[RAWP + x0], p22 = q[RAWP + xp];
      // same tap order as the fp32 front end (frontend.cu), so the value that is rounded to fp16 is the fp32 feature
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
      sob[c * 2] = dy;
      sob[c * 2 + 1] = dx;
    }
  }
#pragma unroll
  for (int g = 0; g < CG; ++g) {
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = feature(PASS * KPASS + g * 8 + e, v, sob);
    uint4 o;
    o.x = pack2(f[0], f[1]);
    o.y = pack2(f[2], f[3]);
    o.z = pack2(f[4], f[5]);
    o.w = pack2(f[6], f[7]);
    *reinterpret_cast<uint4*>(dst + g * CGP) = o;
  }
}

struct Tile { int n, oy0, ox0; };
__device__ __forceinline__ Tile tile_decode(int t, const Params& p) {
  Tile k;
  const int per = p.tiles_y * p.tiles_x;
  k.n = t / per;
  const int r = t - k.n * per;
  const int ty = r / p.tiles_x;
  k.oy0 = ty * TR;
  k.ox0 = (r - ty * p.tiles_x) * TC;
  return k;
}

template <int PASS>
__device__ __forceinline__ void gen_pass(const Params& p, const float* __restrict__ raw, int iy0, int ix0,
                                         unsigned char* fb, int tid) {
#pragma unroll 1
  for (int i = tid; i < LR * LC; i += NPROD) {
    const int ly = i / LC;
    const int rr = i - ly * LC;
    const int px = rr >= Q ? 1 : 0;
    const int q = rr - px * Q;
    const int lx = 2 * q + px;
    const int iy = iy0 + ly, ix = ix0 + lx;
    // consecutive items = consecutive 16-byte slots (11 even columns, 10 odd ones, next row 85 = 21 (mod 8) slots on):
    // every quarter-warp store covers all 32 banks
    unsigned char* dst = fb + ly * RP + px * PP + q * 16;
#if defined(SHDR_C1_DBG) && (SHDR_C1_DBG & 1)   // development A/B only: no feature generation
    if (i < 0) {
#else
    if (iy >= 0 && iy < p.h && ix >= 0 && ix < p.w) {
#endif
      gen_pixel<PASS>(raw + (ly + 1) * RAWP, lx + 1, dst);
    } else {                                               // the convolution's zero padding
      const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int g = 0; g < CG; ++g) *reinterpret_cast<uint4*>(dst + g * CGP) = z;
    }
  }
}

// the fp32 pixels under the tile plus a 1-pixel ring, read from global memory ONCE per tile (the three passes and the
// 3 x 3 Sobel windows then read shared memory: global loads at a 12-byte lane stride cost ~5 L1 wavefronts each and
// the L1 data path is what the tensor cores' operand reads saturate).  Ring pixels that are Sobel neighbours of
// border pixels hold the REFLECTed pixel (-1 -> 1, n -> n-2); positions further out are never used.
__device__ __forceinline__ void stage_raw(const Params& p, const float* __restrict__ img_n, int iy0, int ix0,
                                          float* raw, int tid) {
#pragma unroll 1
  for (int i = tid; i < RAW_R * RAW_C; i += NPROD) {
    const int ry = i / RAW_C, rx = i - ry * RAW_C;
    int iy = iy0 - 1 + ry, ix = ix0 - 1 + rx;
    iy = iy < 0 ? -iy : (iy >= p.h ? 2 * p.h - 2 - iy : iy);
    ix = ix < 0 ? -ix : (ix >= p.w ? 2 * p.w - 2 - ix : ix);
    iy = min(max(iy, 0), p.h - 1);
    ix = min(max(ix, 0), p.w - 1);
    const float* src = img_n + ((size_t)iy * p.w + ix) * 3;
    float* d = raw + ry * RAWP + raw_col(rx);
    d[0] = __ldg(src);
    d[RAWC] = __ldg(src + 1);
    d[2 * RAWC] = __ldg(src + 2);
  }
}

// Tiles of this CTA: t = blockIdx.x + i * gridDim.x.  In pair mode both CTAs of a cluster must run the same number of
// iterations (the leader's MMAs drive both), so every role loops over base = t - rank, and a CTA whose own tile index
// falls off the end runs a dummy tile (no image loads, no stores).
#define SHDR_FOR_TILES(base, rank) for (int base = (int)blockIdx.x - (int)(rank); base < p.ntiles; base += (int)gridDim.x)

template <bool PAIR>
__device__ void producer(const Params& p, unsigned char* fbuf, float* raw, uint64_t* ffull, uint64_t* fempty, int tid,
                         unsigned rank) {
  const int lane = tid & 31;
  const uint32_t ffull_leader = PAIR ? map_to_rank(ffull, 0) : 0;   // pair mode: both CTAs' producers report to the leader
  unsigned gp = 0;                                         // running pass number: buffer gp & 1, use (gp >> 1)
  SHDR_FOR_TILES(base, rank) {
    const int t = base + (int)rank;
    const bool valid = t < p.ntiles;
    int iy0 = 0, ix0 = 0;
    if (valid) {
      const Tile k = tile_decode(t, p);
      const float* img_n = p.img + (size_t)k.n * p.h * p.w * 3;
      iy0 = 2 * k.oy0 - p.pt; ix0 = 2 * k.ox0 - p.pl;
      asm volatile("bar.sync 1, %0;" ::"n"(NPROD) : "memory");   // every producer is done reading the previous raw tile
      stage_raw(p, img_n, iy0, ix0, raw, tid);
      asm volatile("bar.sync 1, %0;" ::"n"(NPROD) : "memory");
    }
#pragma unroll 1
    for (int pass = 0; pass < NPASS; ++pass, ++gp) {
      const unsigned b = gp & 1;
      unsigned char* fb = fbuf + b * FBUF;
      mbar_wait_backoff(fempty + b, ((gp >> 1) & 1) ^ 1);  // the MMAs that read this buffer two passes ago are done
      if (valid) {
        if (pass == 0) gen_pass<0>(p, raw, iy0, ix0, fb, tid);
        else if (pass == 1) gen_pass<1>(p, raw, iy0, ix0, fb, tid);
        else gen_pass<2>(p, raw, iy0, ix0, fb, tid);
      }
      fence_async_smem();                                  // generic-proxy stores -> visible to tcgen05.mma
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(ffull_leader + b * 8);
        else mbar_arrive(ffull + b);
      }
    }
  }
}

// one thread: a whole kernel row of a pass (7 taps; 28 KB, or this CTA's 14 KB half of the output channels in pair
// mode -- contiguous in the packed image) per bulk copy and per barrier
template <bool PAIR>
__device__ void weight_loader(const Params& p, unsigned char* wbuf, uint64_t* wfull, uint64_t* wempty, unsigned rank) {
  using R = Ring<PAIR>;
  unsigned st = 0, ph = 0;
  SHDR_FOR_TILES(base, rank) {
#pragma unroll 1
    for (int s = 0; s < NPASS * 7; ++s) {
      mbar_wait(wempty + st, ph ^ 1);
#if defined(SHDR_C1_DBG) && (SHDR_C1_DBG & 2)   // development A/B only: no weight traffic
      mbar_arrive(wfull + st);
#else
      mbar_expect_tx(wfull + st, R::ROW);
      bulk_load(wbuf + st * R::ROW, p.wpk + (size_t)(PAIR ? s * 2 + (int)rank : s) * R::ROW, R::ROW, wfull + st);
#endif
      if (++st == R::DEPTH) { st = 0; ph ^= 1; }
    }
  }
}

// pair mode, CTA 1, one thread: tells the leader when this CTA's half of a ring stage has landed
__device__ void weight_relay(const Params& p, uint64_t* wfull, unsigned rank) {
  using R = Ring<true>;
  const uint32_t wfull_leader = map_to_rank(wfull, 0);
  unsigned st = 0, ph = 0;
  SHDR_FOR_TILES(base, rank) {
#pragma unroll 1
    for (int s = 0; s < NPASS * 7; ++s) {
      mbar_wait(wfull + st, ph);
      mbar_arrive_cluster(wfull_leader + st * 8);
      if (++st == R::DEPTH) { st = 0; ph ^= 1; }
    }
  }
}

// TWO issuer warps, alternating kernel rows.  Each is executed by the WHOLE warp (so that every address stays on the
// uniform datapath) and one elected lane issues.  The tensor pipe queues only a few MMAs (a 128 x 64 x 16 MMA lasts
// ~48 clk), so with one issuer the barrier poll, election and descriptor set-up between two rows (~350 clk) were lost
// time; with two, one warp prepares its row and blocks at the full queue while the other's 14 MMAs run.  Each warp
// accumulates ITS rows into its OWN tensor-memory tile (the epilogue adds the two), so the summation order -- and with
// it every output bit -- does not depend on how the two issue streams interleave.
// Pair mode: only the even CTA of the cluster runs this; every MMA is M = 256 (this CTA's tile and the peer's), B is
// read half from each CTA, and every commit arrives on the barriers of both CTAs.  The barriers that the PEER's threads
// arrive on (release at cluster scope, after their proxy fence) are waited for with the plain CTA-scope try_wait: this
// thread never reads the peer's data itself -- the tensor core does, in the peer's SM -- and an acquire at cluster
// scope compiles to CCTL.IVALL, which waits for every bulk copy in flight (measured: 0.43 ms instead of 0.25).
template <bool PAIR>
__device__ void mma_issuer(const Params& p, unsigned char* fbuf, unsigned char* wbuf, uint64_t* ffull, uint64_t* fempty,
                           uint64_t* wfull, uint64_t* wempty, uint64_t* afull, uint64_t* aempty,
                           uint32_t tmem, unsigned par) {
  using R = Ring<PAIR>;
  constexpr uint32_t IDESC = idesc_f16_f32(PAIR ? 256 : 128, OC);
  const uint64_t ad0 = smem_desc_nosw(smem_u32(fbuf), CGP, SBO_A);
  const uint64_t bd0 = smem_desc_nosw(smem_u32(wbuf), R::LBO, 128);
  const uint32_t a_hi = (uint32_t)(ad0 >> 32), b_hi = (uint32_t)(bd0 >> 32);
  const uint32_t a_lo0 = (uint32_t)ad0, b_lo0 = (uint32_t)bd0;
  unsigned gr = 0, it = 0, gp = 0;                         // running kernel-row / tile / pass numbers
  SHDR_FOR_TILES(base, 0) {
    const unsigned ab = it & 1, use = it >> 1;
    // the epilogue(s) have drained this accumulator
    mbar_wait(aempty + ab, (use & 1) ^ 1);
    fence_after_sync();
    const uint32_t acc = tmem + (ab * 2 + par) * OC;
    bool first = true;                                     // this warp's first row of the tile zero-initialises its tile
#pragma unroll 1
    for (int pass = 0; pass < NPASS; ++pass, ++gp) {
      const unsigned fbi = gp & 1;
      mbar_wait(ffull + fbi, (gp >> 1) & 1);
      fence_after_sync();
#pragma unroll 1
      for (int ky = 0; ky < 7; ++ky, ++gr) {
        if ((gr & 1) != par) continue;
        const unsigned st = gr % R::DEPTH, ph = (gr / R::DEPTH) & 1;
        mbar_wait(wfull + st, ph);
        fence_after_sync();
        if (elect_one()) {
          const uint32_t a_row = a_lo0 + ((fbi * FBUF + ky * RP) >> 4);
          const uint32_t b_row = b_lo0 + st * (R::ROW >> 4);
#pragma unroll
          for (int kx = 0; kx < 7; ++kx) {
#pragma unroll
            for (int c = 0; c < KPASS / 16; ++c) {
              const uint32_t a_lo = a_row + (((kx & 1) * PP + (kx >> 1) * 16 + 2 * c * CGP) >> 4);
              const uint32_t b_lo = b_row + ((kx * R::TAP + c * R::KSTEP) >> 4);
              const uint32_t accum = (first && (kx | c) == 0) ? 0u : 1u;
              if (PAIR) mma_ss2_pair(acc, a_lo, a_hi, b_lo, b_hi, IDESC, accum);
              else mma_ss2(acc, a_lo, a_hi, b_lo, b_hi, IDESC, accum);
            }
          }
          // ring stage free once these MMAs have read it
          if (PAIR) mma_commit_pair(wempty + st); else mma_commit(wempty + st);
        }
        first = false;
        __syncwarp();
      }
      if (elect_one()) {                                    // this warp's MMAs on the feature buffer are done
        if (PAIR) mma_commit_pair(fempty + fbi); else mma_commit(fempty + fbi);
      }
      __syncwarp();
    }
    if (elect_one()) {                                      // this warp's share of the accumulator is complete
      if (PAIR) mma_commit_pair(afull + ab); else mma_commit(afull + ab);
    }
    __syncwarp();
    ++it;
  }
}

template <bool PAIR>
__device__ void epilogue(const Params& p, uint64_t* afull, uint64_t* aempty, uint32_t tmem, int wq, int lane, unsigned rank) {
  unsigned it = 0;
  const int m = wq * 32 + lane;
  const int r = m >> 3, j = m & 7;
  const uint32_t aempty_leader = PAIR ? map_to_rank(aempty, 0) : 0;
  SHDR_FOR_TILES(base, rank) {
    const int t = base + (int)rank;
    const bool valid = t < p.ntiles;
    Tile k = {0, 0, 0};
    if (valid) k = tile_decode(t, p);
    const unsigned ab = it & 1, use = it >> 1;
    ++it;
    mbar_wait_backoff(afull + ab, use & 1);
    fence_after_sync();
    const int oy = k.oy0 + r, ox = k.ox0 + j;
    const bool ok = valid && oy < p.oh && ox < p.ow;
    float* o = p.out + (((size_t)k.n * p.oh + oy) * p.ow + ox) * OC;
#pragma unroll
    for (int part = 0; part < 4; ++part) {                 // 16 output channels at a time (register budget of 20 warps)
      float v[16], v2[16];                                 // the two issuer warps' partial sums
      tmem_ld16(tmem + ((uint32_t)(wq * 32) << 16) + (ab * 2) * OC + part * 16, v);
      tmem_ld16(tmem + ((uint32_t)(wq * 32) << 16) + (ab * 2 + 1) * OC + part * 16, v2);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] += v2[i];
      if (part == 3) {                                     // every value of this accumulator is in registers
        fence_before_sync();
        if (PAIR) mbar_arrive_cluster(aempty_leader + ab * 8);
        else mbar_arrive(aempty + ab);
      }
      if (ok) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int ch = part * 16 + g * 4;
          float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.scale) sc = __ldg(reinterpret_cast<const float4*>(p.scale + ch));
          if (p.shift) sh = __ldg(reinterpret_cast<const float4*>(p.shift + ch));
          float4 y;
          y.x = fmaf(v[g * 4 + 0], sc.x, sh.x);
          y.y = fmaf(v[g * 4 + 1], sc.y, sh.y);
          y.z = fmaf(v[g * 4 + 2], sc.z, sh.z);
          y.w = fmaf(v[g * 4 + 3], sc.w, sh.w);
          if (p.relu) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
          __stcs(reinterpret_cast<float4*>(o + ch), y);
        }
      }
    }
  }
}

template <bool PAIR>
__device__ __forceinline__ void conv1_body(const Params& p) {
  using R = Ring<PAIR>;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* fbuf = smem;
  unsigned char* wbuf = smem + 2 * FBUF;
  float* raw = reinterpret_cast<float*>(smem + 2 * FBUF + R::DEPTH * R::ROW);
  __shared__ uint64_t bars[2 * Ring<true>::DEPTH + 8];
  __shared__ uint32_t tmem_base;
  uint64_t* wfull = bars;
  uint64_t* wempty = bars + R::DEPTH;
  uint64_t* ffull = bars + 2 * R::DEPTH;
  uint64_t* fempty = ffull + 2;
  uint64_t* afull = ffull + 4;
  uint64_t* aempty = ffull + 6;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned rank = PAIR ? cluster_ctarank() : 0;

  if (tid == 0) {
    for (int s = 0; s < R::DEPTH; ++s) {
      mbar_init(wfull + s, (PAIR && rank == 0) ? 2 : 1);   // leader: its own copy + the peer's relay
      mbar_init(wempty + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(ffull + s, (PAIR ? 2 : 1) * (NPROD / 32)); // pair mode: the producer warps of both CTAs
      mbar_init(fempty + s, 2);                            // one commit per issuer warp
      mbar_init(afull + s, 2);
      mbar_init(aempty + s, (PAIR ? 2 : 1) * 128);         // pair mode: the epilogue threads of both CTAs
    }
    mbar_init_fence();
  }
  if (warp == W_MMA) {
    if (PAIR) tmem_alloc_pair<TMEM_COLS>(&tmem_base); else tmem_alloc<TMEM_COLS>(&tmem_base);
  }
  fence_before_sync();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base;

  if (warp < W_LOAD) producer<PAIR>(p, fbuf, raw, ffull, fempty, tid, rank);
  else if (warp == W_LOAD) { if (lane == 0) weight_loader<PAIR>(p, wbuf, wfull, wempty, rank); __syncwarp(); }
  else if (warp <= W_MMA + 1) {
    if (!PAIR || rank == 0) mma_issuer<PAIR>(p, fbuf, wbuf, ffull, fempty, wfull, wempty, afull, aempty, tmem, warp - W_MMA);
    else if (warp == W_MMA) { if (lane == 0) weight_relay(p, wfull, rank); __syncwarp(); }
  }
  else epilogue<PAIR>(p, afull, aempty, tmem, warp & 3, lane, rank);

  fence_before_sync();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == W_MMA) {
    if (PAIR) tmem_free_pair<TMEM_COLS>(tmem); else tmem_free<TMEM_COLS>(tmem);
  }
}

#undef SHDR_FOR_TILES

__global__ void __launch_bounds__(NTHREADS, 1) k_frontend_conv1(const Params p) { conv1_body<false>(p); }
// CTA pairs: the two CTAs of a cluster sit on the two SMs of one TPC and share every MMA (cta_group::2)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1) k_frontend_conv1_pair(const Params p) {
  conv1_body<true>(p);
}

}  // namespace c1
}  // namespace shdr

using namespace shdr;

extern "C" size_t shdr_conv1_packed_bytes(void) { return c1::PACKED_BYTES; }

extern "C" int shdr_conv1_pack_weights_f32(const float* kernel_hwio, void* packed, void* stream) {
  SHDR_REQUIRE(kernel_hwio && packed, "conv1_pack_weights: NULL pointer");
  SHDR_REQUIRE(aligned16(packed), "conv1_pack_weights: packed must be 16-byte aligned");
  DeviceGuard g(packed);
  if (g.status != SHDR_OK) return g.status;
  const int total = c1::NPASS * c1::NTAP * c1::KPASS * c1::OC;
  c1::k_pack_weights<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(kernel_hwio, (__half*)packed);
  SHDR_LAUNCH_CHECK("k_pack_weights");
  c1::k_pack_weights_pair<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      kernel_hwio, (__half*)((unsigned char*)packed + c1::PACKED_ONE));
  SHDR_LAUNCH_CHECK("k_pack_weights_pair");
  return SHDR_OK;
}

extern "C" int shdr_frontend_conv1_f32(const float* img, const void* packed, const float* scale, const float* shift,
                                       int relu, float* out, int n, int h, int w, void* stream) {
  SHDR_REQUIRE(n >= 0 && h >= 0 && w >= 0, "frontend_conv1: bad shape n=%d h=%d w=%d", n, h, w);
  if ((long long)n * h * w == 0) return SHDR_OK;
  SHDR_REQUIRE(img && packed && out, "frontend_conv1: NULL pointer");
  SHDR_REQUIRE(h >= 2 && w >= 2, "frontend_conv1: REFLECT padding needs h >= 2 and w >= 2 (got %d x %d)", h, w);
  SHDR_REQUIRE(aligned16(packed) && aligned16(out), "frontend_conv1: packed and out must be 16-byte aligned");
  SHDR_REQUIRE((!scale || aligned16(scale)) && (!shift || aligned16(shift)), "frontend_conv1: scale / shift must be 16-byte aligned");
  DeviceGuard g(out);
  if (g.status != SHDR_OK) return g.status;
  c1::Params p;
  p.img = img; p.wpk = (const unsigned char*)packed; p.scale = scale; p.shift = shift; p.out = out;
  p.n = n; p.h = h; p.w = w; p.relu = relu;
  // TF 'SAME', stride 2, 7 taps: out = ceil(in / 2), pad = max((out - 1) 2 + 7 - in, 0), the smaller half in front
  p.oh = (h + 1) / 2; p.ow = (w + 1) / 2;
  p.pt = ((p.oh - 1) * 2 + 7 - h) / 2; p.pl = ((p.ow - 1) * 2 + 7 - w) / 2;
  p.tiles_y = (p.oh + c1::TR - 1) / c1::TR; p.tiles_x = (p.ow + c1::TC - 1) / c1::TC;
  const long long nt = (long long)n * p.tiles_y * p.tiles_x;
  SHDR_REQUIRE(nt < 0x7fffffffLL && (long long)n * h * w < 0x7fffffffLL, "frontend_conv1: too many pixels for int32 tile indices");
  p.ntiles = (int)nt;
  const int sms = sm_count(g.dev);
  if (nt >= 4) {
    // CTA pairs (the two SMs of a TPC share every MMA and each half of the weights): clusters of 2, an even grid
    p.wpk += c1::PACKED_ONE;
    SHDR_CUDA(cudaFuncSetAttribute(c1::k_frontend_conv1_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, c1::SMEM_BYTES));
    const long long pairs = (nt + 1) / 2;
    const int grid = 2 * (int)(pairs < sms / 2 ? pairs : sms / 2);
    c1::k_frontend_conv1_pair<<<grid, c1::NTHREADS, c1::SMEM_BYTES, (cudaStream_t)stream>>>(p);
    SHDR_LAUNCH_CHECK("k_frontend_conv1_pair");
    return SHDR_OK;
  }
  SHDR_CUDA(cudaFuncSetAttribute(c1::k_frontend_conv1, cudaFuncAttributeMaxDynamicSharedMemorySize, c1::SMEM_BYTES));
  const int grid = (int)(nt < sms ? nt : sms);
  c1::k_frontend_conv1<<<grid, c1::NTHREADS, c1::SMEM_BYTES, (cudaStream_t)stream>>>(p);
  SHDR_LAUNCH_CHECK("k_frontend_conv1");
  return SHDR_OK;
}
