// pooled_slide.cu -- soft histograms B = {4, 8, 16} fused with the 16x16 stride-1 'same' average pool, as an EXACT
// integer sliding-window pipeline (the headline kernel, BASELINE.json configs[1]); optionally with img + Sobel in
// front (the pooled 93-channel front end in ONE launch).
//
// Reference behaviour restated (ShinYwings/SingleHDR-tf2): model.histogram_layer, linearization_net.py:336-350, then
// average_pooling2d(h, 16, 1, 'same'), linearization_net.py:351 (README.md:51): window rows [y-7, y+8], columns
// [x-7, x+8] clipped to the image, divided by the number of in-bounds elements; concat order of :322.
//
// Why integers.  For power-of-two B every vote 1 - |I - c_b|*B is a multiple of 2^-24 in [0, 1] (for I >= c_1 it is
// exact in fp32), so votes are 25-bit fixed-point numbers and every window sum is an exact
// 32-bit integer.  Exact sums make a SLIDING window legal (add the entering row, subtract the leaving one -- in fp32
// the cancellation would destroy the pure-relative 1e-5 tolerance and the "all-zero window stays exactly zero"
// property), and a sliding window needs only ONE word of state per (column, channel).  That turns the kernel into a
// stream down the image: no 2-D tile halo to recompute vertically, and every output row of a 64-pixel strip is a
// single contiguous chunk in HBM (64 x 336 B = 21 KB), sent with one bulk copy -- no partial sectors for any channel
// count, which is what made the 84-channel tensor awkward for the tiled kernel (pooled_ws.cu).
//
// Votes without a vote: with T = rne(I * B * 2^24) and G_b = clamp(T - (b - 1.5) * 2^24, 0, 2^24), the triangular
// vote of bin b is G_b - G_{b+1}.  G is ONE instruction (DPX VIADDMNMX.RELU: max(min(a + b, c), 0)), and both the
// clamp and the window sum commute with the difference, so the producers slide B + 1 running sums per (column,
// colour, B) and difference them when a row is emitted.
//
// Pipeline (one persistent CTA per SM, 20 warps, tasks = image x 64-column strip x row segment; 15 warm-up rows
// per task):
//   producers (8 warps, one thread per (TWO adjacent columns incl. 7 + 8 halo, colour, {B4 + B8 | B16})):
//       global (register prefetch two rows ahead, L2 prefetch 8 rows ahead) -> T (fixed point) -> private 16-row ring
//       in shared memory (the leaving row's T); S_b += G_b(T_enter) - G_b(T_leave) (the add / subtract as IMAD on the
//       fma pipe: the alu pipe, which runs the DPX clamps, is the scarce one); column vote sums S_b - S_{b+1} ->
//       stage[channel][column] with 64-bit stores
//   consumers (12 warps = 2 row groups x 3 channel groups x 2 half strips, one LANE PER CHANNEL):
//       47 column sums -> 32 sliding horizontal sums (exact, < 2^32: one column per window is capped at 2^28 - 1) ->
//       fp32 -> x 1/(count * 2^24) -> staging[pixel][channel] (lanes = consecutive channels: conflict-free) -> one
//       cp.async.bulk store per row, staging double-buffered per group (one named barrier per row).
//   (FULL: each consumer lane also computes img + Sobel of one (pixel, colour) of the row into the staging row.)
// Hand-off through an mbarrier ring of 3 column-sum stages; the two consumer groups alternate rows.
// All shared-memory traffic is conflict-free by construction (lanes = consecutive column pairs or consecutive
// channels; stage pitch 84 = 20 mod 32 words for the 128-bit loads of 8 consecutive channels).
// History, measurements and the experiments that did not pay: profiles/README.md (round 2), DESIGN.md section 4.1.
#include <cuda.h>

#include "common.cuh"

// This file is compiled twice: as is (rows leave as bulk copies: the hot path) and, through pooled_slide_nb.cu, with
// SHDR_SLIDE_NOBULK (rows leave as cooperative 4-byte stores: 93 channels with w % 4 != 0, or an unaligned output).
// Two translation units, because merely instantiating a second variant next to the hot kernel changed ptxas' code for
// it (measured: 0.621 instead of 0.597 ms on config2).
#ifndef SHDR_SLIDE_NOBULK
#define SHDR_SL_NS sl
#else
#define SHDR_SL_NS sl_nb
#endif

namespace shdr {
namespace SHDR_SL_NS {

constexpr int PK = 16, HL = 7, HR = 8;       // TF SAME: 7 before, 8 after
constexpr int SW = 64;                       // output columns per strip
constexpr int CW = SW + HL + HR;             // 79 input columns
constexpr int NH = 84;                       // histogram channels: 3 * (4 + 8 + 16)
constexpr int VP = 84;                       // ints per channel line of a stage (79 + pad; 84 = 20 mod 32)
constexpr int NST = 3;                       // column-sum stages
constexpr int STAGE_INTS = NH * VP;          // 7056 ints = 28 KB
constexpr int NCP = (CW + 1) / 2;            // 40 column pairs (the 80th column is padding)
constexpr int NPT = 3 * NCP;                 // 120 (column pair, colour) items per producer half
constexpr int RING_P = 3 * NPT;              // int2 per ring row: T4 | T8 | T16 of every item
constexpr int NPW = 8, NCW = 12;             // producer / consumer warps
constexpr int NPROD = NPW * 32, NCONS = NCW * 32, THREADS = NPROD + NCONS;
constexpr int GROUP = NCONS / 2;             // threads of one consumer row group
constexpr int CAPV = 1 << 24;                // vote 1.0
constexpr int T_OUT = -(1 << 30);            // "no pixel": every G is 0
constexpr int PFD = 8;                       // rows of L2 prefetch distance for the input
constexpr int TAB = 17 * 17 + 3;             // 1 / (ny * nx * 2^24) for ny, nx = 0..16 (+ pad)

template <bool FULL> struct Cfg {
  static constexpr int CO = FULL ? SHDR_FRONTEND_CH : SHDR_HIST_CH;   // floats per output pixel
  static constexpr int CH0 = FULL ? 9 : 0;                            // first histogram channel
  static constexpr size_t SMEM = (size_t)(NST * STAGE_INTS + 16 * RING_P * 2 + 4 * SW * CO + 2 * SW + TAB) * 4 + 2 * NST * 8;
};

struct Params {
  const float* img;
  float* out;
  int n, h, w;
  int nsx, nsy, rseg, ntasks;
  // +1 and -1 as OPAQUE kernel parameters: "s += g * one" compiles to IMAD (fma pipe) where "s += g" would be an IADD3
  // on the alu pipe, which VIADDMNMX / I2FP / compares already saturate (pipe balance)
  int one, mone;
};

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n\t.reg .pred p;\n"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n"
      "DONE_%=:\n\t}" ::"r"(a), "r"(parity) : "memory");
}
// consumers have slack (they wait for the producers ~30 % of the time): after a failed first poll they sleep ~100 ns
// between polls, so that the polls do not take issue slots from the producer warps of the same scheduler
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n\t.reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n"
      "WAIT_%=:\n\t"
      "nanosleep.u32 %2;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n"
      "DONE_%=:\n\t}" ::"r"(a), "r"(parity), "r"(100) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bulk_store(float* gdst, const float* ssrc, unsigned bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct Task { int n, x0, y0, y1; };
__device__ __forceinline__ Task task_decode(int t, const Params& p) {
  Task k;
  const int per_img = p.nsx * p.nsy;
  k.n = t / per_img;
  const int r = t - k.n * per_img;
  const int sy = r / p.nsx;
  k.x0 = (r - sy * p.nsx) * SW;
  k.y0 = sy * p.rseg;
  k.y1 = min(k.y0 + p.rseg, p.h);
  return k;
}

// fixed-point intensity for histogram B: rne(I * B * 2^24); I is already clamped to [-2, 3] (NaN -> -2: no vote)
__device__ __forceinline__ int to_fix(float v, float scale) { return __float2int_rn(v * scale); }

// G_j of a histogram, j = 0..B: clamp(T - (j - 0.5) * 2^24, 0, 2^24)   (bin b = j + 1 has vote G_j - G_{j+1})
__device__ __forceinline__ int gval(int t, int j) { return __viaddmin_s32_relu(t, (1 << 23) - j * (1 << 24), CAPV); }

// ------------------------------------------------------------------------------------------ producers
// One producer thread owns TWO adjacent columns of one colour (64-bit shared-memory accesses, half the per-row
// bookkeeping per column).  HALF 0: histograms B = 4 (5 G) and B = 8 (9 G); HALF 1: B = 16 (17 G).
template <int HALF>
__device__ __forceinline__ void producer(const Params& p, int* __restrict__ sS, int2* __restrict__ ring, uint64_t* bars,
                                         int ptid) {
  constexpr int NG = HALF ? 17 : 14;
  const int rem = ptid - HALF * (NPROD / 2);
  const bool act = rem < NPT;
  const int colour = act ? rem / NCP : 0;
  const int cp = act ? rem - colour * NCP : 0;                   // columns 2*cp, 2*cp + 1
  const int lane = ptid & 31;
  int2* myring = ring + (HALF ? 2 * NPT : 0) + (act ? rem : 0);  // HALF 0: T4 at +0, T8 at +NPT
  int2* mydst = reinterpret_cast<int2*>(sS + (colour + (HALF ? 36 : 0)) * VP) + cp;   // first channel line, stage 0
  const int h = p.h, rs3 = p.w * 3;
  const int pfoff = (PFD - 1) * rs3;
  const int one = p.one, mone = p.mone;
  int S0[NG], S1[NG];
  unsigned s = 0, ph = 1;                  // stage of the next emitted row; parity of its "empty" wait (first use passes)
  for (int t = blockIdx.x; t < p.ntasks; t += gridDim.x) {
    const Task k = task_decode(t, p);
    const int gx = k.x0 - HL + 2 * cp;
    const bool xok0 = act && gx >= 0 && gx < p.w;
    const bool xok1 = act && gx + 1 >= 0 && gx + 1 < p.w && 2 * cp + 1 < CW;
    const float* src = p.img + ((long long)k.n * h * p.w + gx) * 3 + colour;   // column 2*cp; the next one is src + 3
    // The register prefetch (two rows ahead) only covers an L2 hit; under the write stream a DRAM read takes longer
    // than that, so every eighth column also pulls the row PFD rows ahead into L2 (a row of the strip is 948 B).
    const bool pf = xok0 && colour == 0 && ((cp & 3) == 0 || cp == NCP - 1);
#pragma unroll
    for (int j = 0; j < NG; ++j) { S0[j] = 0; S1[j] = 0; }
    if (act) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        myring[i * RING_P] = make_int2(T_OUT, T_OUT);
        if (!HALF) myring[i * RING_P + NPT] = make_int2(T_OUT, T_OUT);
      }
    }
    int r = max(k.y0 - HL, 0);             // the row that enters next; rows are addressed as src[off], off = row * rs3
    int off = r * rs3;
#pragma unroll 1
    for (int i = 1; i < PFD; ++i)
      if (pf && r + i < h) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + off + i * rs3));
    // register prefetch TWO rows ahead: n* is the row that enters next, f* the one after it
    float n0 = (xok0 && r < h) ? __ldg(src + off) : -2.0f;     // out of the image: no vote
    float n1 = (xok1 && r < h) ? __ldg(src + off + 3) : -2.0f;
    float f0 = (xok0 && r + 1 < h) ? __ldg(src + off + rs3) : -2.0f;
    float f1 = (xok1 && r + 1 < h) ? __ldg(src + off + rs3 + 3) : -2.0f;
    float c0, c1;
    // one row step: clamp the row loaded two steps ago (NaN -> -2: votes 0, like tf.where on a NaN compare), start the
    // load of the row two steps ahead and the L2 prefetch PFD rows ahead
    auto advance = [&]() {
      c0 = fminf(fmaxf(n0, -2.0f), 3.0f);
      c1 = fminf(fmaxf(n1, -2.0f), 3.0f);
      n0 = f0;
      n1 = f1;
      off += rs3;
      const bool rok = r + 2 < h;
      f0 = (xok0 && rok) ? __ldg(src + off + rs3) : -2.0f;
      f1 = (xok1 && rok) ? __ldg(src + off + rs3 + 3) : -2.0f;
      if (pf && r + PFD < h) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + off + pfoff));
    };
    // warm-up: rows y0-7 .. y0+7 enter, nothing leaves, nothing is emitted
    for (; r <= k.y0 + HL; ++r) {
      advance();
      int2* rs = myring + (r & 15) * RING_P;
      if (HALF) {
        const int2 tn = make_int2(to_fix(c0, 268435456.0f), to_fix(c1, 268435456.0f));
        if (act) rs[0] = tn;
#pragma unroll
        for (int j = 0; j < 17; ++j) { S0[j] += gval(tn.x, j); S1[j] += gval(tn.y, j); }
      } else {
        const int2 ta = make_int2(to_fix(c0, 67108864.0f), to_fix(c1, 67108864.0f));
        const int2 tb = make_int2(to_fix(c0, 134217728.0f), to_fix(c1, 134217728.0f));
        if (act) { rs[0] = ta; rs[NPT] = tb; }
#pragma unroll
        for (int j = 0; j < 5; ++j) { S0[j] += gval(ta.x, j); S1[j] += gval(ta.y, j); }
#pragma unroll
        for (int j = 0; j < 9; ++j) { S0[5 + j] += gval(tb.x, j); S1[5 + j] += gval(tb.y, j); }
      }
    }
    // steady state: row r = y+8 enters, row y-8 leaves, row y is emitted
    const int rend = k.y1 + HR;
    for (; r < rend; ++r) {
      advance();
      int2* rs = myring + (r & 15) * RING_P;
      if (HALF) {
        const int2 tn = make_int2(to_fix(c0, 268435456.0f), to_fix(c1, 268435456.0f));
        const int2 to = rs[0];
        if (act) rs[0] = tn;
#pragma unroll
        for (int j = 0; j < 17; ++j) {
          S0[j] += gval(tn.x, j) * one; S0[j] += gval(to.x, j) * mone;
          S1[j] += gval(tn.y, j) * one; S1[j] += gval(to.y, j) * mone;
        }
      } else {
        const int2 ta = make_int2(to_fix(c0, 67108864.0f), to_fix(c1, 67108864.0f));
        const int2 tb = make_int2(to_fix(c0, 134217728.0f), to_fix(c1, 134217728.0f));
        const int2 oa = rs[0], ob = rs[NPT];
        if (act) { rs[0] = ta; rs[NPT] = tb; }
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          S0[j] += gval(ta.x, j) * one; S0[j] += gval(oa.x, j) * mone;
          S1[j] += gval(ta.y, j) * one; S1[j] += gval(oa.y, j) * mone;
        }
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          S0[5 + j] += gval(tb.x, j) * one; S0[5 + j] += gval(ob.x, j) * mone;
          S1[5 + j] += gval(tb.y, j) * one; S1[5 + j] += gval(ob.y, j) * mone;
        }
      }
      mbar_wait(bars + NST + s, ph);                           // consumers have read the previous row in this stage
      if (act) {
        // column vote sums over the 16 rows: S_b - S_{b+1} <= 2^28
        int2* dst = mydst + s * (STAGE_INTS / 2);
        if (HALF) {
#pragma unroll
          for (int i = 0; i < 16; ++i) dst[(3 * i) * (VP / 2)] = make_int2(S0[i] - S0[i + 1], S1[i] - S1[i + 1]);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) dst[(3 * i) * (VP / 2)] = make_int2(S0[i] - S0[i + 1], S1[i] - S1[i + 1]);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            dst[(12 + 3 * i) * (VP / 2)] = make_int2(S0[5 + i] - S0[6 + i], S1[5 + i] - S1[6 + i]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + s);                    // release: this warp's column sums are in the stage
      if (++s == NST) { s = 0; ph ^= 1u; }
    }
  }
}

// Sobel of one (pixel, colour): q points at the pixel's colour sample in row y; up / dn are the offsets (in floats)
// of the rows above / below and oxm / oxp those of the columns left / right, REFLECT already applied (-1 -> 1,
// n -> n-2).  Same tap order as k_frontend_strip (frontend.cu).
__device__ __forceinline__ void sobel_at(const float* __restrict__ q, int up, int dn, int oxm, int oxp, float& v,
                                         float& dy, float& dx) {
  const float* qu = q + up;
  const float* qd = q + dn;
  const float p00 = __ldg(qu + oxm), p01 = __ldg(qu), p02 = __ldg(qu + oxp);
  const float p10 = __ldg(q + oxm), p12 = __ldg(q + oxp);
  const float p20 = __ldg(qd + oxm), p21 = __ldg(qd), p22 = __ldg(qd + oxp);
  v = __ldg(q);
  dy = -p00;
  dy = __fadd_rn(dy, -2.0f * p01);
  dy = __fsub_rn(dy, p02);
  dy = __fadd_rn(dy, p20);
  dy = __fadd_rn(dy, 2.0f * p21);
  dy = __fadd_rn(dy, p22);
  dx = -p00;
  dx = __fadd_rn(dx, p02);
  dx = __fadd_rn(dx, -2.0f * p10);
  dx = __fadd_rn(dx, 2.0f * p12);
  dx = __fsub_rn(dx, p20);
  dx = __fadd_rn(dx, p22);
}

// ------------------------------------------------------------------------------------------ consumers
// One output row of one lane, part 1: 47 column sums of its channel -> 32 window sums -> fp32 (-> scaled when one
// scale serves the whole row).  Runs BEFORE the group's staging buffer is known to be free, so it overlaps the bulk
// store of the group's previous row.
template <bool SCALE>
__device__ __forceinline__ void window_sums(const int4* __restrict__ vl, float (&f)[32], float sc, uint64_t* empty_bar,
                                            int lane) {
  unsigned v[48];
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    const int4 x4 = vl[i];
    v[4 * i + 0] = (unsigned)x4.x; v[4 * i + 1] = (unsigned)x4.y;
    v[4 * i + 2] = (unsigned)x4.z; v[4 * i + 3] = (unsigned)x4.w;
  }
  __syncwarp();
  if (lane == 0) mbar_arrive(empty_bar);                 // this warp holds its 47 column sums in registers
  // A column sum can be 2^28 (sixteen exact 1.0 votes) and sixteen of those would wrap to 0.  Every window of 16
  // consecutive columns holds exactly one column whose strip index is a multiple of 16: capping those at 2^28 - 1
  // keeps every window sum below 2^32 and moves an all-ones window by 2^-32 relative (below the fp32 rounding).
  v[0] = min(v[0], (1u << 28) - 1); v[16] = min(v[16], (1u << 28) - 1); v[32] = min(v[32], (1u << 28) - 1);
  // two independent chains (outputs 0..15 and 16..31) halve the dependent-add latency
  unsigned h0 = 0, h1 = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) { h0 += v[i]; h1 += v[16 + i]; }
  unsigned H[32];
  H[0] = h0; H[16] = h1;
#pragma unroll
  for (int j = 1; j < 16; ++j) {
    H[j] = H[j - 1] + v[j + 15] - v[j - 1];
    H[16 + j] = H[16 + j - 1] + v[16 + j + 15] - v[16 + j - 1];
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) f[j] = __uint2float_rn(H[j]);
  if (SCALE) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] *= sc;
  }
}

template <bool FULL>
__device__ __forceinline__ void consumer(const Params& p, const int* __restrict__ sS, float* __restrict__ sStg,
                                         float* __restrict__ sRow, const float* __restrict__ sTab, uint64_t* bars,
                                         int ctid) {
  constexpr int CO = Cfg<FULL>::CO, CH0 = Cfg<FULL>::CH0;
  const int lane = ctid & 31;
  const int cw = ctid >> 5;
  const int g = cw / 6;                    // row group: emitted rows with (q & 1) == g
  const int sub = cw - 6 * g;
  const int seg = sub & 1;                 // half strip: output columns 32*seg .. +31
  const int ch = (sub >> 1) * 32 + lane;   // this lane's histogram channel
  const bool actv = ch < NH;
  const int chr = actv ? ch : NH - 1;
  const int gtid = ctid - g * GROUP;
  const bool elected = (gtid == 0);
  float* stg0 = sStg + g * (2 * SW * CO);  // this group's two staging rows (alternating)
  float* rowsc = sRow + g * SW;
  const int myoff = (seg * 32) * CO + CH0 + ch;
  const int* myS = sS + chr * VP + seg * 32;
  unsigned q = 0;                          // emission index of the next task's first row
  unsigned par = 0;                        // staging buffer of this group's next row
  for (int t = blockIdx.x; t < p.ntasks; t += gridDim.x) {
    const Task k = task_decode(t, p);
    const bool xedge = (k.x0 == 0) || (k.x0 + SW + HR > p.w);
    const int vw = min(SW, p.w - k.x0);
    int nx = PK;                           // in-bounds columns of this thread's scale-table column (gtid < 64)
    if (gtid < SW) {
      const int gx = k.x0 + gtid;
      nx = min(max(min(gx + HR, p.w - 1) - max(gx - HL, 0) + 1, 0), PK);
    }
    // FULL: this lane's (pixel, colour) item of every row of the task (64 pixels x 3 colours over the group's 192 lanes)
    int fpx = -1, fc = 0, oxm = 0, oxp = 0;
    const float* fq = nullptr;             // the item's sample in row 0 of the image
    const int rs3 = p.w * 3;
    if (FULL) {
      const int item = sub * 32 + lane;
      const int px = item / 3;
      fc = item - px * 3;
      const int gx = k.x0 + px;
      if (gx < p.w) {
        fpx = px;
        oxm = (reflect1(gx - 1, p.w) - gx) * 3;
        oxp = (reflect1(gx + 1, p.w) - gx) * 3;
        fq = p.img + ((long long)k.n * p.h * p.w + gx) * 3 + fc;
      }
    }
    // this group's rows of the task: those whose emission index q0 + (y - y0) has parity g
    const unsigned q0 = q;
    const int nrows = k.y1 - k.y0;
    q += (unsigned)nrows;
    const int first = (int)((g - q0) & 1u);
    float* orow = p.out + (((long long)k.n * p.h + k.y0 + first) * p.w + k.x0) * CO;
    const long long ostride = 2LL * p.w * CO;
    for (int yy = first; yy < nrows; yy += 2, orow += ostride, par ^= 1u) {
      const int y = k.y0 + yy;
      const unsigned qq = q0 + (unsigned)yy;
      float* stg = stg0 + par * (SW * CO);
      float fv = 0.f, fdy = 0.f, fdx = 0.f;
      if (FULL && fpx >= 0)                // img + Sobel of this lane's (pixel, colour); the loads are issued early
        sobel_at(fq + (long long)y * rs3, y == 0 ? rs3 : -rs3, y == p.h - 1 ? -rs3 : rs3, oxm, oxp, fv, fdy, fdx);
      const int ny = min(y + HR, p.h - 1) - max(y - HL, 0) + 1;
      const unsigned s = qq % NST, ph = (qq / NST) & 1u;
      mbar_wait_backoff(bars + s, ph);     // producers filled this stage
      const int4* vl = reinterpret_cast<const int4*>(myS + s * STAGE_INTS);
      float f[32];
      if (xedge) {                         // the in-bounds count varies along the row: one scale per column
        window_sums<false>(vl, f, 0.f, bars + NST + s, lane);
        if (gtid < SW) rowsc[gtid] = sTab[ny * 17 + nx];
        named_bar_sync(1 + g, GROUP);      // scales visible (the previous row's readers passed the row barrier)
        const float* rs = rowsc + seg * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] *= rs[j];
      } else {
        window_sums<true>(vl, f, sTab[ny * 17 + PK], bars + NST + s, lane);
      }
      // the staging buffer was last read by the bulk store issued two rows ago, which the elected thread saw complete
      // (its reads) before the previous row barrier
      if (actv) {
        float* my = stg + myoff;
#pragma unroll
        for (int j = 0; j < 32; ++j) my[j * CO] = f[j];
      }
      if (FULL && fpx >= 0) {
        float* o = stg + fpx * CO;
        o[fc] = fv;
        o[3 + fc * 2] = fdy;
        o[4 + fc * 2] = fdx;
      }
#ifndef SHDR_SLIDE_NOBULK
      fence_async_smem();                  // staging writes -> visible to the bulk-copy (async) proxy
      if (elected) bulk_wait_read();       // the previous row's store has left the OTHER staging buffer
      named_bar_sync(1 + g, GROUP);        // the whole row is staged; the other buffer is free for the next row
      if (elected) {
        bulk_store(orow, stg, (unsigned)(vw * CO * 4));
        bulk_commit();
      }
#else
      // rows that are not 16-byte aligned chunks: the group streams the staged row out with coalesced 4-byte stores;
      // every thread is past this loop before it reaches the next row's barrier, i.e. before anybody writes this
      // staging buffer again (two rows on)
      named_bar_sync(1 + g, GROUP);
      const int nfl = vw * CO;
      for (int i = gtid; i < nfl; i += GROUP) st_stream1(orow + i, stg[i]);
#endif
    }
  }
#ifndef SHDR_SLIDE_NOBULK
  if (elected) bulk_wait_all();
#endif
}

template <bool FULL>
__global__ void __launch_bounds__(THREADS, 1) k_pool_slide(const __grid_constant__ Params p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* sStg = reinterpret_cast<float*>(smem_raw);                       // bulk-copy source: 16-byte aligned
  int* sS = reinterpret_cast<int*>(sStg + 4 * SW * Cfg<FULL>::CO);
  int2* ring = reinterpret_cast<int2*>(sS + NST * STAGE_INTS);
  float* sRow = reinterpret_cast<float*>(ring + 16 * RING_P);
  float* sTab = sRow + 2 * SW;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sTab + TAB);               // full[NST], empty[NST]
  const int tid = threadIdx.x;
  if (tid < 17 * 17) {                     // 1 / (in-bounds count * 2^24): 2^-24 undoes the fixed-point vote scale
    const int cnt = (tid / 17) * (tid % 17);
    sTab[tid] = cnt ? __fdiv_rn(1.0f, (float)cnt) * (1.0f / 16777216.0f) : 0.0f;
  }
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(bars + s, NPW);
      mbar_init(bars + NST + s, NCW / 2);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid < NPROD / 2) producer<0>(p, sS, ring, bars, tid);
  else if (tid < NPROD) producer<1>(p, sS, ring, bars, tid);
  else consumer<FULL>(p, sS, sStg, sRow, sTab, bars, tid - NPROD);
}

template <bool FULL>
static int launch_t(const Params& p, int sms, cudaStream_t st) {
  SHDR_CUDA(cudaFuncSetAttribute(k_pool_slide<FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)Cfg<FULL>::SMEM));
  const int grid = p.ntasks < sms ? p.ntasks : sms;
  k_pool_slide<FULL><<<grid, THREADS, Cfg<FULL>::SMEM, st>>>(p);
  SHDR_LAUNCH_CHECK("k_pool_slide");
  return SHDR_OK;
}

}  // namespace sl / sl_nb

#ifdef SHDR_SLIDE_NOBULK
int launch_pool_slide_nobulk(const sl_nb::Params& p0, bool full93, int sms, cudaStream_t st) {
  return full93 ? sl_nb::launch_t<true>(p0, sms, st) : sl_nb::launch_t<false>(p0, sms, st);
}
#else
namespace sl_nb { struct Params; }
int launch_pool_slide_nobulk(const sl_nb::Params& p0, bool full93, int sms, cudaStream_t st);   // pooled_slide_nb.cu

// true when the sliding-window kernel can run this request: the {4, 8, 16} histograms into a dense 84-channel tensor,
// or (full93) into channels 9..92 of the 93-channel front-end tensor together with img + Sobel in channels 0..8.
bool pool_slide_supported(const float* out, int w, const int* bins, int nbins, bool full93) {
  (void)out; (void)w; (void)full93;
  return nbins == 3 && bins[0] == 4 && bins[1] == 8 && bins[2] == 16;
}

int launch_pool_slide(const float* img, float* out, int n, int h, int w, bool full93, int dev, cudaStream_t st) {
  sl::Params p;
  p.img = img; p.out = out; p.n = n; p.h = h; p.w = w;
  p.one = 1; p.mone = -1;
  p.nsx = (w + sl::SW - 1) / sl::SW;
  const int sms = sm_count(dev);
  // row segments: every task pays a warm-up of 15 rows (about 6 rows' worth of work); pick the split that minimises
  // waves x (rows per task + warm-up)
  long long best_cost = -1;
  int best_nsy = 1;
  for (int nsy = 1; nsy <= (h + 31) / 32; ++nsy) {
    const int rseg = (h + nsy - 1) / nsy;
    const long long tasks = (long long)n * p.nsx * ((h + rseg - 1) / rseg);
    const long long waves = (tasks + sms - 1) / sms;
    const long long cost = waves * (rseg + 6);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_nsy = nsy; }
  }
  p.rseg = (h + best_nsy - 1) / best_nsy;
  p.nsy = (h + p.rseg - 1) / p.rseg;
  const long long total = (long long)n * p.nsx * p.nsy;
  SHDR_REQUIRE(total > 0 && total < 0x7fffffffLL, "pool_slide: %lld tasks out of range", total);
  SHDR_REQUIRE((long long)h * w * 3 < 0x7fffffffLL, "pool_slide: one image of %d x %d exceeds the 32-bit row offsets", h, w);
  p.ntasks = (int)total;
  // The per-row bulk copies need 16-byte aligned chunks: always true for 84 channels (336 B per pixel); for 93 channels
  // (372 B per pixel) the image width must be a multiple of 4.  Otherwise: the same kernel with cooperative row stores
  // (identical Params layout, its own translation unit).
  if (!aligned16(out) || (full93 && (w % 4) != 0))
    return launch_pool_slide_nobulk(reinterpret_cast<const sl_nb::Params&>(p), full93, sms, st);
  return full93 ? sl::launch_t<true>(p, sms, st) : sl::launch_t<false>(p, sms, st);
}

#endif  // SHDR_SLIDE_NOBULK

}  // namespace shdr
