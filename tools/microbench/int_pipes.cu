// Microbenchmark (sm_100a): issue throughput of the integer instructions the exact sliding-window pooled kernel
// is built from -- VIADDMNMX(.RELU) (DPX add+min+relu), IADD3, I2FP.F32.U32, F2I, VIMNMX -- alone and in the
// producer / consumer mixes.  Reports warp-instructions per clock per SM (clock from SM cycle counters).
#include <cuda_runtime.h>
#include <stdio.h>

template <int MODE>
__global__ void k(int* out, int iters, int seed, long long* clk) {
  int x[16];
  int kk[16];
  for (int i = 0; i < 16; ++i) { x[i] = seed * (i + 1) + threadIdx.x; kk[i] = (seed ^ i) - 7; }
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {          // VIADDMNMX.RELU
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = __viaddmin_s32_relu(x[i], kk[i], 1 << 24);
    } else if (MODE == 1) {   // IADD3
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("add.s32 %0, %0, %1;" : "+r"(x[i]) : "r"(kk[i]));
    } else if (MODE == 2) {   // I2FP.F32.U32
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = __float_as_int(__uint2float_rn((unsigned)x[i]));
    } else if (MODE == 3) {   // F2I
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = __float2int_rn(__int_as_float(x[i] | 0x3f000000));
    } else if (MODE == 4) {   // producer mix: per G 2 x VIADDMNMX.RELU + IADD3(a + b - c); 8 G per iteration + 8 diffs
      int tn = x[15], to = x[14];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int gn = __viaddmin_s32_relu(tn, kk[i], 1 << 24);
        const int go = __viaddmin_s32_relu(to, kk[i], 1 << 24);
        x[i] = x[i] + gn - go;
      }
#pragma unroll
      for (int i = 0; i < 7; ++i) out[(threadIdx.x + i * 32) & 1023] = __viaddmin_s32(x[i], -x[i + 1], (1 << 28) - 1);
      x[15] += 12345; x[14] -= 321;
    } else if (MODE == 5) {   // consumer mix: IADD3(a + b - c) + I2FP + FMUL per output
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        x[i] = x[i] + kk[i] - kk[(i + 1) & 15];
        const float f = __uint2float_rn((unsigned)x[i]) * 2.3283064365386963e-10f;
        kk[i] ^= __float_as_int(f);
      }
    } else if (MODE == 6) {   // VIMNMX
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = max(x[i], kk[i]) - 1;
    } else if (MODE == 7) {   // IADD3 as a + b - c (three-input form)
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = x[i] + kk[i] - kk[(i + 3) & 15];
    }
  }
  const long long t1 = clock64();
  int s = 0;
  for (int i = 0; i < 16; ++i) s += x[i] + kk[i];
  out[1024 + blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <int MODE> void run(const char* name, double ops_per_iter) {
  int* out; long long* clk; cudaMalloc(&out, (1024 + 148 * 2 * 512) * 4); cudaMalloc(&clk, 8);
  const int iters = 20000;
  k<MODE><<<148 * 2, 512>>>(out, 100, 3, clk);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148 * 2, 512>>>(out, iters, 3, clk);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
  // per SM: 2 blocks x 16 warps
  printf("%-44s %.3f ms  %lld clk  warp-instr/clk/SM %.2f (listed ops only)\n", name, ms, c,
         32.0 * iters * ops_per_iter / (double)c);
  cudaFree(out); cudaFree(clk);
}

int main() {
  run<0>("VIADDMNMX.RELU x16", 16);
  run<1>("IADD (2-input) x16", 16);
  run<7>("IADD3 a+b-c x16", 16);
  run<6>("VIMNMX + IADD x16", 32);
  run<2>("I2FP.F32.U32 x16", 16);
  run<3>("LOP + F2I x16", 32);
  run<4>("producer mix 16 DPX + 8 IADD3 + 7 DPX + 7 STS", 16 + 8 + 7 + 7);
  run<5>("consumer mix 16 x (IADD3 + I2FP + FMUL + LOP)", 64);
  return 0;
}
