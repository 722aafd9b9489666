// Unit test + microbenchmark (sm_100a) of the tcgen05 building blocks the fused conv1 kernel relies on:
//   * K-major NO-swizzle shared-memory descriptors with a non-trivial LBO / SBO (the im2col view of the feature tile:
//     core matrices 176 B apart in K, 4224 B apart in M) and the packed weight layout (LBO 1024, SBO 128),
//   * instruction descriptor bf16 x bf16 -> fp32, M = 128, N = 64, accumulation over several K = 16 steps,
//   * tensor-memory allocation, tcgen05.commit -> mbarrier, tcgen05.ld 32x32b.
// Prints the max abs error against a host GEMM, then the MMA issue rate.  (--swap tries the other reading of
// (LBO, SBO): on the B200 it faults with an illegal shared-memory access, i.e. lbo = K step, sbo = M/N step is right.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I singlehdr-tf2_b200/csrc tools/microbench/umma_nosw.cu -o tools/_build/umma_nosw
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "umma.cuh"

using namespace shdr::umma;

constexpr int M = 128, N = 64, KSTEPS = 3, K = 16 * KSTEPS;
constexpr int A_LBO = 176, A_SBO = 4224, A_BYTES = 16 * A_SBO;          // 67584
constexpr int B_LBO = 1024, B_SBO = 128, B_BYTES = KSTEPS * 2048;       // 6144

// element (m, k) of A / (n, k) of B -> byte offset in the staged image
static int a_off(int m, int k) { return (m / 8) * A_SBO + (k / 8) * A_LBO + (m % 8) * 16 + (k % 8) * 2; }
static int b_off(int n, int k) { return (k / 16) * 2048 + ((k / 8) % 2) * B_LBO + (n / 8) * B_SBO + (n % 8) * 16 + (k % 8) * 2; }

__global__ void __launch_bounds__(128)
k_test(const uint4* __restrict__ a_img, const uint4* __restrict__ b_img, float* __restrict__ out, int swap, int reps,
       long long* clk, int commit_every = 0, int two_acc = 0) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sA = smem;
  unsigned char* sB = smem + A_BYTES;
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2;   // sink of the intermediate commits
  __shared__ uint32_t tbase;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < A_BYTES / 16; i += 128) reinterpret_cast<uint4*>(sA)[i] = a_img[i];
  for (int i = tid; i < B_BYTES / 16; i += 128) reinterpret_cast<uint4*>(sB)[i] = b_img[i];
  if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); mbar_init_fence(); }
  if (warp == 0) tmem_alloc<128>(&tbase);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tbase;
  const uint32_t idesc = idesc_bf16_f32(M, N);
  long long t0 = 0, t1 = 0;
  if (warp == 0) {                       // whole warp, one elected lane issues (descriptors stay on the uniform datapath)
    const uint64_t ad0 = swap ? smem_desc_nosw(smem_u32(sA), A_SBO, A_LBO) : smem_desc_nosw(smem_u32(sA), A_LBO, A_SBO);
    const uint64_t bd0 = swap ? smem_desc_nosw(smem_u32(sB), B_SBO, B_LBO) : smem_desc_nosw(smem_u32(sB), B_LBO, B_SBO);
    t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < reps; ++r) {
      if (elect_one()) {
#pragma unroll
        for (int s = 0; s < KSTEPS; ++s)
          mma_ss2(tm + ((two_acc && (r & 1)) ? 64u : 0u), (uint32_t)ad0 + ((s * 2 * A_LBO) >> 4), (uint32_t)(ad0 >> 32),
                  (uint32_t)bd0 + ((s * 2048) >> 4), (uint32_t)(bd0 >> 32), idesc, (r | s) ? 1u : 0u);
        if (commit_every && ((r + 1) & (commit_every - 1)) == 0) mma_commit(&bar2);
      }
      __syncwarp();
    }
    if (elect_one()) mma_commit(&bar);
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  if (tid == 0) { t1 = clock64(); clk[0] = t1 - t0; }
  fence_after_sync();
  float v[32];
  for (int half = 0; half < 2; ++half) {
    tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + half * 32, v);
    for (int i = 0; i < 32; ++i) out[(warp * 32 + (tid & 31)) * N + half * 32 + i] = v[i];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_free<128>(tm);
}

int main(int argc, char** argv) {
  const int try_swap = argc > 1 && !strcmp(argv[1], "--swap");
  unsigned char* a = (unsigned char*)calloc(A_BYTES, 1);
  unsigned char* b = (unsigned char*)calloc(B_BYTES, 1);
  float* fa = (float*)malloc(M * K * 4);
  float* fb = (float*)malloc(N * K * 4);
  srand(1);
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) {
      __nv_bfloat16 h = __float2bfloat16((float)(rand() % 2001 - 1000) / 500.0f);
      fa[m * K + k] = __bfloat162float(h);
      memcpy(a + a_off(m, k), &h, 2);
    }
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      __nv_bfloat16 h = __float2bfloat16((float)(rand() % 2001 - 1000) / 700.0f);
      fb[n * K + k] = __bfloat162float(h);
      memcpy(b + b_off(n, k), &h, 2);
    }
  uint4 *da, *db;
  float* dout;
  long long* dclk;
  cudaMalloc(&da, A_BYTES); cudaMalloc(&db, B_BYTES); cudaMalloc(&dout, M * N * 4); cudaMalloc(&dclk, 8);
  cudaMemcpy(da, a, A_BYTES, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b, B_BYTES, cudaMemcpyHostToDevice);
  const int smem = A_BYTES + B_BYTES;
  cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  float* out = (float*)malloc(M * N * 4);
  int ok = 0;
  for (int swap = 0; swap < 1 + try_swap; ++swap) {
    cudaMemset(dout, 0, M * N * 4);
    k_test<<<1, 128, smem>>>(da, db, dout, swap, 1, dclk);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("swap=%d: CUDA error %s\n", swap, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(out, dout, M * N * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double r = 0;
        for (int k = 0; k < K; ++k) r += (double)fa[m * K + k] * fb[n * K + k];
        maxerr = fmax(maxerr, fabs(r - out[m * N + n]));
        maxref = fmax(maxref, fabs(r));
      }
    printf("descriptor reading %s: max |err| = %.3e (max |ref| = %.2f) %s\n",
           swap ? "(lbo<->sbo swapped)" : "(lbo = K step, sbo = M/N step)", maxerr, maxref, maxerr < 1e-3 ? "OK" : "WRONG");
    if (maxerr < 1e-3) ok |= 1 << swap;
  }
  if (ok) {
    const int swap = (ok & 1) ? 0 : 1;
    for (int reps : {64, 512}) {
      k_test<<<1, 128, smem>>>(da, db, dout, swap, reps, dclk);
      cudaDeviceSynchronize();
      long long c;
      cudaMemcpy(&c, dclk, 8, cudaMemcpyDeviceToHost);
      printf("%d MMAs (128x64x16, operands in shared memory): %lld clk = %.1f clk / MMA\n", reps * KSTEPS, c,
             (double)c / (reps * KSTEPS));
    }
    // does tcgen05.commit cost tensor-pipe time?  one commit (to a sink barrier) after every ce x 3 MMAs
    for (int ce : {1, 2, 4, 8}) {
      k_test<<<1, 128, smem>>>(da, db, dout, swap, 512, dclk, ce, 0);
      cudaDeviceSynchronize();
      long long c;
      cudaMemcpy(&c, dclk, 8, cudaMemcpyDeviceToHost);
      printf("commit after every %2d MMAs: %.1f clk / MMA\n", ce * KSTEPS, (double)c / (512 * KSTEPS));
    }
    k_test<<<1, 128, smem>>>(da, db, dout, swap, 512, dclk, 0, 1);
    cudaDeviceSynchronize();
    long long c2;
    cudaMemcpy(&c2, dclk, 8, cudaMemcpyDeviceToHost);
    printf("two accumulators alternating every 3 MMAs, no commits: %.1f clk / MMA\n", (double)c2 / (512 * KSTEPS));
  }
  printf("RESULT %s\n", ok == 1 ? "PASS" : (ok ? "PASS-SWAPPED" : "FAIL"));
  return ok ? 0 : 2;
}
