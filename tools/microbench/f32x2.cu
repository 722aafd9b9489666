// Microbenchmark: issue throughput of FADD vs FADD2 (add.rn.f32x2) and FFMA vs FFMA2 on sm_100a.
// Each thread runs 8 independent dependency chains; reports warp-instructions per clock per SM.
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack(u64 r, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float a[16]; u64 p[8];
  for (int i = 0; i < 16; ++i) a[i] = seed + i + threadIdx.x;
  for (int i = 0; i < 8; ++i) p[i] = pack(a[2 * i], a[2 * i + 1]);
  u64 c = pack(seed, seed * 0.5f);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = __fadd_rn(a[i], seed);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(c));
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], seed, 1.0f);
    } else if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(c));
    } else if (MODE == 4) {   // mixed: FADD2 + FMNMX (alu pipe) interleaved
#pragma unroll
      for (int i = 0; i < 8; ++i) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(c)); a[i] = fmaxf(a[i], seed); a[i+8] = fminf(a[i+8], seed); }
    } else if (MODE == 6) {   // 8 FADD2 + 8 x (SHF + LOP3) on the ALU pipe
#pragma unroll
      for (int i = 0; i < 8; ++i) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(c)); unsigned x = __float_as_uint(a[i]); x = (x << 13) ^ x; a[i] = __uint_as_float(x); }
    } else if (MODE == 7) {   // 16 FADD + 8 x (SHF + LOP3)
#pragma unroll
      for (int i = 0; i < 8; ++i) { a[i+8] = __fadd_rn(a[i+8], seed); float t; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(t) : "f"(a[i+8]), "f"(seed)); a[i+8] = t; unsigned x = __float_as_uint(a[i]); x = (x << 13) ^ x; a[i] = __uint_as_float(x); }
    } else {                  // mixed scalar: 2 FADD + 2 FMNMX
#pragma unroll
      for (int i = 0; i < 8; ++i) { a[i] = __fadd_rn(a[i], seed); a[i+8] = __fadd_rn(a[i+8], seed); a[i] = fmaxf(a[i], seed); a[i+8] = fminf(a[i+8], seed); }
    }
  }
  float s = 0;
  for (int i = 0; i < 16; ++i) s += a[i];
  for (int i = 0; i < 8; ++i) { float x, y; unpack(p[i], x, y); s += x + y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, int ops_per_iter, int flops_per_iter) {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
  int iters = 20000;
  k<MODE><<<148 * 2, 512>>>(out, 100, 1.0001f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148 * 2, 512>>>(out, iters, 1.0001f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double warp_instr = 148.0 * 2 * 16 * (double)iters * ops_per_iter;
  double clk = ms * 1e-3 * 1.965e9;
  printf("%-28s %.3f ms  warp-instr/clk/SM %.2f  lane-flop/clk/SM %.1f\n", name, ms, warp_instr / clk / 148,
         148.0 * 2 * 512 * (double)iters * flops_per_iter / clk / 148);
  cudaFree(out);
}
int main() {
  run<0>("FADD x16", 16, 16);
  run<1>("FADD2 x8", 8, 16);
  run<2>("FFMA x16", 16, 32);
  run<3>("FFMA2 x8", 8, 32);
  run<4>("FADD2 x8 + FMNMX x16", 24, 16);
  run<5>("FADD x16 + FMNMX x16", 32, 16);
  run<6>("FADD2 x8 + (SHF+LOP3) x8", 24, 16);
  run<7>("FADD x16 + (SHF+LOP3) x8", 32, 16);
  return 0;
}
