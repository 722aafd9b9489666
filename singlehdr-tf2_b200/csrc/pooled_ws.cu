// pooled_ws.cu -- warp-specialised, persistent kernel for the fused soft histogram + 16x16 'same'
// average pool, fast path (power-of-two B, dense [n,h,w,C] output, even image width).  See pooled.cu
// for the reference behaviour restated, the numerics and the generic block-synchronous kernel.
//
// Two measured facts shape this kernel (profiles/r1, tools/microbench/store_patterns.cu):
//  1. B200 absorbs PARTIAL 32-byte sectors ~5x slower than full ones: writing the 84-channel output
//     as 48-byte (12-channel) fragments caps at 1.6 TB/s no matter how little compute runs, while
//     scattered but whole, aligned sectors reach 4.5 TB/s.  A pixel is 336 B = 10.5 sectors, so
//     sector boundaries fall on channel multiples of 8 for even pixels and on 8k+4 for odd pixels.
//     => the channel axis is processed in UNITS OF 8 CHANNELS (one sector); an even pixel stores its
//     8 fresh channels directly; an odd pixel stores {4 channels held from the previous unit in
//     registers, 4 fresh channels} with ONE instruction (8 adjacent lanes = one whole sector).  The
//     sector that straddles an (even, odd) pixel pair is written at the last unit from the even
//     pixel's 4 fresh channels and the odd pixel's first 4 channels held since unit 0.
//  2. A block-synchronous two-pass kernel idles at barriers because the passes have different item
//     counts.  => ONE persistent CTA per SM runs a producer/consumer pipeline over (tile, unit):
//       producers (10 warps)  votes + vertical 16-window sums  -> column-sum ring sV[3 stages]
//       consumers ( 8 warps)  horizontal 16-window sums, scale, whole-sector staging + TMA stores  <- sV[stage]
//     with mbarrier hand-off and no block-wide barrier; the input tile (+halo) is staged transposed with cp.async
//     (the next one is prefetched into L2 while the current one is processed).
//  3. With (1) and (2) the LSU/L1 data pipe became the limiter, and every 32-byte sector stored with STG costs one
//     of its wavefronts.  => each consumer warp stages its results in shared memory with rows innermost (its 32
//     lanes = 8 channels x 4 rows write 128 contiguous bytes: one conflict-free wavefront) and one lane sends them
//     to HBM with TMA bulk-tensor stores (cp.async.bulk.tensor.4d, SASS UTMASTG): a 4-D tensor map over
//     (float, row, pixel pair, image) addresses whole sectors and clips at the image border by itself.
#include <cuda.h>
#include <string.h>

#include "common.cuh"

namespace shdr {

namespace ws {
constexpr int PK = 16, PB = 7, PA = 8;
constexpr int PT_H = 16, PT_W = 64;
constexpr int IN_H = PT_H + PK - 1;    // 31
constexpr int IN_W = PT_W + PK - 1;    // 79
constexpr int IPITCH = 36;             // floats per transposed input column (31 rows + pad): 9 x 16 B
constexpr int VPITCH = 84;             // floats per (row, channel) line of column sums (79 + pad): 21 x 16 B
constexpr int UC = 8;                  // channels per unit = one 32-byte sector
constexpr int NSTAGE = 3;
constexpr int NPROD = 320, NCONS = 256, THREADS = NPROD + NCONS;
constexpr int SI_FLOATS = 3 * IN_W * IPITCH;     // 8532
constexpr int SV_FLOATS = PT_H * UC * VPITCH;    // 10752
constexpr int RC_FLOATS = PT_W + PT_H;           // 80
constexpr int MAXC = 192;                        // up to 64 bins x RGB per launch
constexpr float SENTINEL = -8.0f;
constexpr int HEAD_FLOATS = PT_H * (PT_W / 2) * 4;   // odd pixels' channels 0..3, parked from unit 0 to the last unit
// Output staging for the TMA stores, PER CONSUMER WARP (a warp owns 4 rows x 32 columns x 8 channels of a unit, so
// only __syncwarp is needed around its bulk stores): EO: [2][16 pixel pairs][4 rows][8 floats] (sector of the even
// pixels, sector of the odd pixels); !EO: [32 pixels][4 rows][8 floats].  Rows are the INNER dimension so that the
// 32 lanes of the warp (8 channels x 4 rows) write 128 contiguous bytes: one conflict-free wavefront.
constexpr int WTILE_FLOATS = 16 * 4 * UC;            // one sector of 16 pixel pairs x 4 rows: 512 floats = 2 KB
constexpr int WSTG_FLOATS = 3 * WTILE_FLOATS;        // per consumer warp: even-pixel tile + two odd-pixel tiles (EO)
constexpr int STG_FLOATS = (NCONS / 32) * WSTG_FLOATS;   // 12288 floats = 48 KB (must stay first: 128-byte aligned)
constexpr size_t SMEM_BYTES =
    (size_t)(STG_FLOATS + SI_FLOATS + NSTAGE * SV_FLOATS + 2 * RC_FLOATS + 3 * MAXC + HEAD_FLOATS) * 4 + 2 * NSTAGE * 8;

struct Params {
  float centre[MAXC];        // bin centre of every output channel, fp32 (2i-1)/(2B)
  float nbins[MAXC];         // float(B) of the channel's histogram
  unsigned char col[MAXC];   // colour plane (0..2) the channel reads
  int C;                     // output channels per pixel (multiple of 4)
  int n, h, w;
  int tiles_x, tiles_y, tiles_total;
};

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gmem_src) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
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
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// one sector-aligned box of the staging buffer -> global; coordinates (float, row, pair|pixel, image)
__device__ __forceinline__ void tma_store4(const CUtensorMap* tmap, const float* smem_src, int c0, int c1, int c2, int c3) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_src);
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
               ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(s) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// vote / 256 = max(fma(-|v - c|, B/256, 1/256), 0).  Scaling by a power of two commutes with every rounding that
// follows (no under/overflow: the smallest non-zero vote is 2^-24), so all window sums are exactly 1/256 of the
// un-scaled ones and interior tiles need no multiply at all.
__device__ __forceinline__ float vote256(float v, float centre, float nbs) {
  return fmaxf(fmaf(-fabsf(__fsub_rn(v, centre)), nbs, 1.0f / 256.0f), 0.0f);
}

struct TileCoord { int n, y0, x0; };
__device__ __forceinline__ TileCoord tile_coord(int t, const Params& p) {
  TileCoord c;
  const int per_img = p.tiles_x * p.tiles_y;
  c.n = t / per_img;
  const int r = t - c.n * per_img;
  const int ty = r / p.tiles_x;
  c.y0 = ty * PT_H;
  c.x0 = (r - ty * p.tiles_x) * PT_W;
  return c;
}

// producers: one thread per staged (column, colour) pair copies that column's 31 rows (transposed)
__device__ __forceinline__ void stage_tile(float* sI, const float* __restrict__ img, const TileCoord tc, int h, int w,
                                           int ptid) {
  if (ptid < IN_W * 3) {
    const int xc = ptid / 3;
    const int c = ptid - xc * 3;
    const int gx = tc.x0 - PB + xc;
    float* dst = sI + (c * IN_W + xc) * IPITCH;
    const bool xok = gx >= 0 && gx < w;
    const float* src = img + (((long long)tc.n * h + (tc.y0 - PB)) * w + gx) * 3 + c;
#pragma unroll 4
    for (int r = 0; r < IN_H; ++r) {
      const int gy = tc.y0 - PB + r;
      if (xok && gy >= 0 && gy < h) cp_async4(dst + r, src + (long long)r * w * 3);
      else dst[r] = SENTINEL;
    }
  }
}

// 16 consecutive output columns of one (row, channel) lane -> the warp's staging tiles, in whole-sector groups.  The
// stores to HBM are TMA bulk-tensor stores of the staged sectors (issued once per unit by one lane): they do not
// occupy the LSU data pipe, which is this kernel's limiter, and they clip at the image border by themselves.
// PHASE (unit-uniform): 0 = first unit, 1 = middle, 2 = last unit (4 channels when EO).
//   even pixel (sector-aligned): its 8 fresh channels are sector u of the pixel pair (tile stE); at the last unit
//     lanes 4..7 add the odd neighbour's first 4 channels, parked in shared memory since unit 0 by lanes 0..3;
//   odd pixel: its sector k is {channels 4..7 of unit k-1, channels 0..3 of unit k}: lanes 0..3 complete the CURRENT
//     odd tile (floats 4..7) and lanes 4..7 start the NEXT one (floats 0..3) -- po is this lane's slot in its tile.
// stE/po point at [pair 0 of this 16-column block][row][slot]; a pair is 4 rows * 8 = 32 floats further.
template <bool EO, int PHASE>
__device__ __forceinline__ void emit16(const float (&res)[16], float* __restrict__ stE, float* __restrict__ po,
                                       float* hp, int f) {
  if (!EO) {                               // every pixel is aligned: stE is [pixel][row][8]
#pragma unroll
    for (int j = 0; j < 16; ++j) stE[j * (4 * UC) + f] = res[j];
    return;
  }
  const bool lo = f < 4;
#pragma unroll
  for (int jp = 0; jp < 8; ++jp) {
    const int je = 2 * jp, jo = je + 1;
    if (PHASE == 0) {
      stE[jp * (4 * UC) + f] = res[je];
      if (lo) hp[jp * 4] = res[jo];                        // park channels 0..3 of the odd pixel
      else po[jp * (4 * UC)] = res[jo];                    // channels 4..7 start odd sector 1
    } else if (PHASE == 1) {
      stE[jp * (4 * UC) + f] = res[je];
      po[jp * (4 * UC)] = res[jo];
    } else {
      stE[jp * (4 * UC) + f] = lo ? res[je] : hp[jp * 4];
      if (lo) po[jp * (4 * UC)] = res[jo];                 // channels 80..83 complete the last odd sector
    }
  }
}

// scale 16 window sums: interior tiles need nothing (votes are pre-scaled by the exact 1/256, see vote256)
template <bool INTERIOR>
__device__ __forceinline__ void scale16(float (&res)[16], const float* __restrict__ sRc, int r, int xb) {
  if (INTERIOR) return;
  const float rcy = 256.0f * sRc[PT_W + r];                // undo the 1/256 (exact)
  const float4* rc4 = reinterpret_cast<const float4*>(sRc + xb * 16);
#pragma unroll
  for (int qd = 0; qd < 4; ++qd) {
    const float4 v = rc4[qd];
    res[4 * qd + 0] *= rcy * v.x;                          // <= 2 ulp from sum / (rows * cols)
    res[4 * qd + 1] *= rcy * v.y;
    res[4 * qd + 2] *= rcy * v.z;
    res[4 * qd + 3] *= rcy * v.w;
  }
}

// One consumer thread, TWO adjacent 16-column blocks of one unit: the 32 horizontal 16-window sums of its
// (row, channel) line from 47 column sums (12 LDS.128), chained van Herk: blocks A = cols 0..15, B = 16..31,
// C = 32..46;  out[j] = suffixA[j] + prefixB[j-1] (j < 16),  out[16+j] = suffixB[j] + prefixC[j-1].
template <bool EO, bool INTERIOR, int PHASE>
__device__ __forceinline__ void consume_pair(const float* __restrict__ vline, float* __restrict__ stE,
                                             float* __restrict__ po, float* hp, const float* __restrict__ sRc,
                                             int r, int xb0, int f, bool active) {
  const float4* vl = reinterpret_cast<const float4*>(vline);
  float a[32], bs[16], res[16];
  active = (PHASE < 2) || active;          // every unit but the last is full: no inactive lanes to zero there
  if (active) {
#pragma unroll
    for (int qd = 0; qd < 8; ++qd) {
      const float4 v = vl[qd];
      a[4 * qd + 0] = v.x; a[4 * qd + 1] = v.y; a[4 * qd + 2] = v.z; a[4 * qd + 3] = v.w;
    }
    bs[15] = a[31];
#pragma unroll
    for (int i = 14; i >= 0; --i) bs[i] = __fadd_rn(a[16 + i], bs[i + 1]);   // suffix sums of B
#pragma unroll
    for (int i = 14; i >= 0; --i) a[i] = __fadd_rn(a[i], a[i + 1]);          // suffix sums of A
#pragma unroll
    for (int i = 17; i < 31; ++i) a[i] = __fadd_rn(a[i], a[i - 1]);          // prefix sums of B (to col 30)
    res[0] = a[0];
#pragma unroll
    for (int j = 1; j < 16; ++j) res[j] = __fadd_rn(a[j], a[15 + j]);
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) { res[j] = 0.0f; bs[j] = 0.0f; }
  }
  scale16<INTERIOR>(res, sRc, r, xb0);
  emit16<EO, PHASE>(res, stE, po, hp, f);

  if (active) {
    float c[16];
#pragma unroll
    for (int qd = 0; qd < 4; ++qd) {
      const float4 v = vl[8 + qd];                         // cols 32..47 (47 is padding, never used)
      c[4 * qd + 0] = v.x; c[4 * qd + 1] = v.y; c[4 * qd + 2] = v.z; c[4 * qd + 3] = v.w;
    }
#pragma unroll
    for (int i = 1; i < 15; ++i) c[i] = __fadd_rn(c[i], c[i - 1]);           // prefix sums of C
    res[0] = bs[0];
#pragma unroll
    for (int j = 1; j < 16; ++j) res[j] = __fadd_rn(bs[j], c[j - 1]);
  }
  scale16<INTERIOR>(res, sRc, r, xb0 + 1);
  emit16<EO, PHASE>(res, stE + (EO ? 8 : 16) * (4 * UC), po + 8 * (4 * UC), hp + 32, f);
}

// EO: C == 4 (mod 8) -> pixel pitch is an odd number of half-sectors, even/odd pixels alternate alignment.
//     C == 0 (mod 8) -> every pixel is sector-aligned (EO = false).
// CT: compile-time channel count (pixel pitch in floats) so that store offsets are immediates; 0 = runtime.
template <int CT, bool EO>
__global__ void __launch_bounds__(THREADS, 1)
k_hist_pooled_ws(const float* __restrict__ img, const __grid_constant__ CUtensorMap tmap,
                 const __grid_constant__ Params prm) {
  extern __shared__ __align__(128) float smem[];
  float* sStage = smem;                    // TMA source: must be 128-byte aligned
  float* sI = smem + STG_FLOATS;
  float* sV0 = sI + SI_FLOATS;
  float* sRc0 = sV0 + NSTAGE * SV_FLOATS;
  float* sCentre = sRc0 + 2 * RC_FLOATS;
  float* sNb = sCentre + MAXC;
  int* sCol = reinterpret_cast<int*>(sNb + MAXC);
  float* sHead = reinterpret_cast<float*>(sCol + MAXC);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sHead + HEAD_FLOATS);   // full[NSTAGE], empty[NSTAGE]
  const int tid = threadIdx.x;
  const int h = prm.h, w = prm.w;
  const int C = CT ? CT : prm.C;
  const int units = (C + UC - 1) / UC;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(bars + s, NPROD);
      mbar_init(bars + NSTAGE + s, NCONS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < MAXC; i += THREADS) {
    sCentre[i] = prm.centre[i];
    sNb[i] = prm.nbins[i];
    sCol[i] = prm.col[i];
  }
  // pad rows 31..35 of the input buffer hold the sentinel for the whole kernel
  for (int i = tid; i < 3 * IN_W * (IPITCH - IN_H); i += THREADS) {
    const int col = i / (IPITCH - IN_H);
    sI[col * IPITCH + IN_H + (i - col * (IPITCH - IN_H))] = SENTINEL;
  }
  __syncthreads();

  if (tid < NPROD) {
    // =========================== producers: staging + pass 1 ===========================
    const int ptid = tid;
    int t = blockIdx.x;
    if (t < prm.tiles_total) stage_tile(sI, img, tile_coord(t, prm), h, w, ptid);
    unsigned q = 0;
    for (; t < prm.tiles_total; t += gridDim.x) {
      cp_async_wait_all();
      named_bar_sync(1, NPROD);          // this tile's input (+halo) has landed
      {
        // pull the NEXT tile's input rows into L2 now, so that the refill of the single input buffer at the end of
        // this tile is an L2 hit (31 rows x 948 B = 8 lines per row)
        const int tnx = t + gridDim.x;
        if (tnx < prm.tiles_total && ptid < IN_H * 8) {
          const TileCoord nc = tile_coord(tnx, prm);
          const int rr = ptid >> 3, ln = ptid & 7;
          const int gy = nc.y0 - PB + rr;
          const int gx = max(nc.x0 - PB, 0);
          if (gy >= 0 && gy < h) {
            const float* p = img + (((long long)nc.n * h + gy) * w + gx) * 3 + ln * 32;
            if (p < img + (long long)prm.n * h * w * 3) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
          }
        }
      }

      for (int u = 0; u < units; ++u, ++q) {
        const unsigned s = q % NSTAGE, uu = q / NSTAGE;
        float* sV = sV0 + s * SV_FLOATS;
        mbar_wait(bars + NSTAGE + s, (uu & 1u) ^ 1u);   // consumers released this stage (passes at once on first use)
        const int nchu = min(UC, C - u * UC);
        // An item is (column, two channels).  Channels p and p+3 of a unit read the same colour plane, so the pair
        // shares ONE fetch of the 31-row input column (8 LDS.128); the two left-over channels of the unit form a
        // mixed pair that re-fetches.  8 channels: (0,3) (1,4) (2,5) (6,7);  4 channels: (0,3) (1,2).
        const int nent = nchu >> 1;
        for (int it = ptid; it < IN_W * nent; it += NPROD) {
          const int e = it / IN_W;
          const int xc = it - e * IN_W;
          int cbA, cbB;
          if (nchu == UC) { cbA = (e < 3) ? e : 6; cbB = (e < 3) ? e + 3 : 7; }
          else            { cbA = e;               cbB = e ? 2 : 3; }
          float I[32];
          int colour = -1;
#pragma unroll 1
          for (int half = 0; half < 2; ++half) {
            const int cb = half ? cbB : cbA;
            const int ch = u * UC + cb;
            const int cl = sCol[ch];
            if (cl != colour) {              // warp-uniform except where a warp straddles two entries
              colour = cl;
              const float4* col = reinterpret_cast<const float4*>(sI + (cl * IN_W + xc) * IPITCH);
#pragma unroll
              for (int qd = 0; qd < 8; ++qd) {
                const float4 v = col[qd];
                I[4 * qd + 0] = v.x; I[4 * qd + 1] = v.y; I[4 * qd + 2] = v.z; I[4 * qd + 3] = v.w;
              }
            }
            const float centre = sCentre[ch], nbs = sNb[ch] * (1.0f / 256.0f);
            float a[31];
#pragma unroll
            for (int i = 0; i < 31; ++i) a[i] = vote256(I[i], centre, nbs);
#pragma unroll
            for (int i = 14; i >= 0; --i) a[i] = __fadd_rn(a[i], a[i + 1]);      // suffix sums, rows 0..15
#pragma unroll
            for (int i = 17; i < 31; ++i) a[i] = __fadd_rn(a[i], a[i - 1]);      // prefix sums, rows 16..30
            float* vcol = sV + cb * VPITCH + xc;
            vcol[0] = a[0];
#pragma unroll
            for (int r = 1; r < 16; ++r) vcol[r * (UC * VPITCH)] = __fadd_rn(a[r], a[15 + r]);
          }
        }
        mbar_arrive(bars + s);             // release: this thread's column sums are in sV[s]
      }
      // the input tile is single-buffered (its second copy pays for the TMA staging buffer): refill it once every
      // producer is done with it; the consumers still have up to NSTAGE units queued, so the ring hides the latency
      named_bar_sync(1, NPROD);
      const int tn = t + gridDim.x;
      if (tn < prm.tiles_total) stage_tile(sI, img, tile_coord(tn, prm), h, w, ptid);
    }
  } else {
    // =========================== consumers: pass 2 + whole-sector stores ===========================
    const int ctid = tid - NPROD;
    const int f = ctid & 7;                // channel within the unit == lane within the sector
    const int r = (ctid >> 3) & 15;        // tile row
    const int xbh = ctid >> 7;             // which half of the tile's four 16-column blocks
    const int xb0 = xbh * 2;
    const int lane = ctid & 31;
    const int wi = ctid >> 5;              // consumer warp: rows 4*(wi&3) .. +3, column half wi>>2
    const int rr = r & 3;                  // row within the warp
    // this warp's staging: EO: tile 0 = even pixels' sector, tiles 1, 2 = odd pixels' sectors (alternating);
    // !EO: tiles 0 and 1 together are [32 pixels][4 rows][8]
    float* wst = sStage + wi * WSTG_FLOATS;
    float* stE = wst + rr * UC;
    const bool lo = f < 4;
    float* hp = sHead + (r * 4 + xb0) * 32 + (f & 3);
    unsigned q = 0;
    int k = 0;
    for (int t = blockIdx.x; t < prm.tiles_total; t += gridDim.x, ++k) {
      const TileCoord tc = tile_coord(t, prm);
      const int x0 = tc.x0, y0 = tc.y0;
      const bool interior = (y0 >= PB) && (y0 + PT_H + PA <= h) && (x0 >= PB) && (x0 + PT_W + PA <= w);
      float* sRc = sRc0 + (k & 1) * RC_FLOATS;
      if (!interior) {                      // tile-uniform: all consumers take the same branch
        // sRc is double-buffered by tile parity, but with only two units per tile (C = 12) the ring lets a fast
        // warp run two tiles ahead of the slowest one, which may still be reading this buffer for tile k - 2:
        // wait for every consumer to have left it before overwriting
        named_bar_sync(2, NCONS);
        if (ctid < PT_W) {
          const int gx = min(x0 + ctid, w - 1);
          sRc[ctid] = __fdiv_rn(1.0f, (float)(min(gx + PA, w - 1) - max(gx - PB, 0) + 1));
        } else if (ctid < PT_W + PT_H) {
          const int gy = min(y0 + ctid - PT_W, h - 1);
          sRc[ctid] = __fdiv_rn(1.0f, (float)(min(gy + PA, h - 1) - max(gy - PB, 0) + 1));
        }
        named_bar_sync(2, NCONS);
      }
      const int ty = y0 + 4 * (wi & 3);     // first row / column of this warp's boxes
      const int tx = x0 + 32 * xbh;
      for (int u = 0; u < units; ++u, ++q) {
        const unsigned s = q % NSTAGE, uu = q / NSTAGE;
        const float* sV = sV0 + s * SV_FLOATS;
        const int nchu = min(UC, C - u * UC);
        const bool last = (u == units - 1);
        const bool active = f < nchu;
        const int phase = (u == 0) ? 0 : (last ? 2 : 1);
        // this warp's previous bulk stores must have finished READING its staging before it is overwritten
        if (lane == 0) tma_wait_read();
        __syncwarp();
        mbar_wait(bars + s, uu & 1u);       // producers filled this stage
        const float* vline = sV + (r * UC + f) * VPITCH + xb0 * 16;
        // odd-pixel tiles: unit u completes tile 1 + (u & 1) (lanes 0..3 -> floats 4..7) and starts the other one
        // (lanes 4..7 -> floats 0..3)
        float* ocur = wst + (1 + (u & 1)) * WTILE_FLOATS;
        float* onext = wst + (2 - (u & 1)) * WTILE_FLOATS;
        float* po = (lo ? ocur + 4 + f : onext + (f - 4)) + rr * UC;
#define SHDR_CONSUME(I, P) consume_pair<EO, I, P>(vline, stE, po, hp, sRc, r, xb0, f, active)
        if (interior) {
          if (phase == 0) SHDR_CONSUME(true, 0); else if (phase == 1) SHDR_CONSUME(true, 1); else SHDR_CONSUME(true, 2);
        } else {
          if (phase == 0) SHDR_CONSUME(false, 0); else if (phase == 1) SHDR_CONSUME(false, 1); else SHDR_CONSUME(false, 2);
        }
#undef SHDR_CONSUME
        mbar_arrive(bars + NSTAGE + s);     // this thread is done reading sV[s]
        fence_async_smem();                 // make this thread's staging writes visible to the async (TMA) proxy
        __syncwarp();                       // the warp's part of the unit is staged (and its parked sHead is visible)
        if (lane == 0) {
          if (!EO) {
            tma_store4(&tmap, wst, u * UC, ty, tx, tc.n);                                   // [32 px][4 rows][8 ch]
          } else {
            tma_store4(&tmap, wst, u * UC, ty, tx >> 1, tc.n);                              // even pixels: sector u
            if (u > 0) tma_store4(&tmap, ocur, C + u * UC - 4, ty, tx >> 1, tc.n);   // odd pixels: sector u + C/8
          }
          tma_commit();
        }
      }
    }
    if (lane == 0) tma_wait_all();          // all bulk stores complete before the CTA exits
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Tensor map of the output for the per-unit bulk stores.  Dimensions (innermost first): float, row, pixel pair
// (EO) or pixel (!EO), image -- rows come before pixels so that the staging buffer is [pixel][row][8].
static int make_tmap(CUtensorMap* tm, float* out, int n, int h, int w, int C, bool eo) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    SHDR_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    SHDR_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available");
    encode = (EncodeTiledFn)fn;
  }
  const cuuint64_t pitch = (cuuint64_t)C * 4;
  cuuint64_t gd[4], gs[3];
  cuuint32_t bx[4], es[4] = {1, 1, 1, 1};
  if (eo) {
    gd[0] = 2 * (cuuint64_t)C; gd[1] = h; gd[2] = w / 2; gd[3] = n;
    gs[0] = pitch * w; gs[1] = 2 * pitch; gs[2] = pitch * w * h;
    bx[0] = UC; bx[1] = 4; bx[2] = 16; bx[3] = 1;          // one consumer warp: 4 rows x 16 pixel pairs
  } else {
    gd[0] = C; gd[1] = h; gd[2] = w; gd[3] = n;
    gs[0] = pitch * w; gs[1] = pitch; gs[2] = pitch * w * h;
    bx[0] = UC; bx[1] = 4; bx[2] = 32; bx[3] = 1;          // one consumer warp: 4 rows x 32 pixels
  }
  CUresult rc = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, out, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (output must be 16-byte aligned)", (int)rc);
    return SHDR_ERR_CUDA;
  }
  return SHDR_OK;
}

template <int CT, bool EO>
static int launch_t(const float* img, float* out, const Params& prm, int sms, cudaStream_t st) {
  CUtensorMap tm;
  int rc = make_tmap(&tm, out, prm.n, prm.h, prm.w, prm.C, EO);
  if (rc != SHDR_OK) return rc;
  SHDR_CUDA(cudaFuncSetAttribute(k_hist_pooled_ws<CT, EO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)SMEM_BYTES));
  const int grid = prm.tiles_total < sms ? prm.tiles_total : sms;
  k_hist_pooled_ws<CT, EO><<<grid, THREADS, SMEM_BYTES, st>>>(img, tm, prm);
  SHDR_LAUNCH_CHECK("k_hist_pooled_ws");
  return SHDR_OK;
}
}  // namespace ws

// true when the warp-specialised kernel can run this request
bool hist_pooled_ws_supported(int w, const int* bins, int nbins, int ostride, int ooff) {
  int C = 0;
  for (int i = 0; i < nbins; ++i) {
    if (bins[i] < 4 || (bins[i] & (bins[i] - 1)) != 0) return false;   // power-of-two B >= 4 only
    C += 3 * bins[i];
  }
  return C <= ws::MAXC && ostride == C && ooff == 0 && (w % 2) == 0;
}

int launch_hist_pooled_ws(const float* img, float* out, int n, int h, int w, const int* bins, int nbins, int dev,
                          cudaStream_t st) {
  ws::Params p;
  int C = 0;
  for (int i = 0; i < nbins; ++i) {
    const int B = bins[i];
    for (int b = 0; b < B; ++b)
      for (int c = 0; c < 3; ++c) {
        p.centre[C] = (float)(2 * b + 1) / (float)(2 * B);   // exact for power-of-two B
        p.nbins[C] = (float)B;
        p.col[C] = (unsigned char)c;
        ++C;
      }
  }
  for (int i = C; i < ws::MAXC; ++i) { p.centre[i] = 0.f; p.nbins[i] = 1.f; p.col[i] = 0; }
  p.C = C;
  p.n = n; p.h = h; p.w = w;
  p.tiles_x = (w + ws::PT_W - 1) / ws::PT_W;
  p.tiles_y = (h + ws::PT_H - 1) / ws::PT_H;
  const long long total = (long long)n * p.tiles_x * p.tiles_y;
  SHDR_REQUIRE(total > 0 && total < 0x7fffffffLL, "hist_pooled_ws: %lld tiles out of range", total);
  p.tiles_total = (int)total;
  const int sms = sm_count(dev);
  switch (C) {
    case 84: return ws::launch_t<84, true>(img, out, p, sms, st);    // B = 4, 8, 16
    case 12: return ws::launch_t<12, true>(img, out, p, sms, st);    // B = 4
    case 24: return ws::launch_t<24, false>(img, out, p, sms, st);   // B = 8
    case 48: return ws::launch_t<48, false>(img, out, p, sms, st);   // B = 16
    default: break;
  }
  return (C % 8 == 4) ? ws::launch_t<0, true>(img, out, p, sms, st) : ws::launch_t<0, false>(img, out, p, sms, st);
}

}  // namespace shdr
