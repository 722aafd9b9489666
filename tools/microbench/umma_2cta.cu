// Unit test + microbenchmark (sm_100a) of the CTA-pair form of tcgen05.mma the fused conv1 kernel uses:
//   cluster of 2 CTAs; M = 256 (each CTA: its own 128 rows of A, its own 128 rows of D in its own tensor memory),
//   N = 64 with each CTA holding 32 of the 64 rows of B (K-major, no swizzle); issue by the even CTA; tcgen05.commit
//   multicast to both CTAs; a remote mbarrier arrive (peer -> leader) announcing the peer's operands.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I singlehdr-tf2_b200/csrc tools/microbench/umma_2cta.cu -o tools/_build/umma_2cta
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "umma.cuh"

using namespace shdr::umma;

constexpr int M = 256, N = 64, KSTEPS = 3, K = 16 * KSTEPS;
constexpr int A_LBO = 336, A_SBO = 2720, A_BYTES = 16 * A_SBO;          // per CTA: 128 rows
constexpr int B_LBO = 512, B_SBO = 128, B_BYTES = KSTEPS * 1024;        // per CTA: 32 rows of B

static int a_off(int m, int k) { return (m / 8) * A_SBO + (k / 8) * A_LBO + (m % 8) * 16 + (k % 8) * 2; }            // m in 0..127
static int b_off(int n, int k) { return (k / 16) * 1024 + ((k / 8) % 2) * B_LBO + (n / 8) * B_SBO + (n % 8) * 16 + (k % 8) * 2; }   // n in 0..31

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
k_test(const uint4* __restrict__ a_img, const uint4* __restrict__ b_img, float* __restrict__ out, int reps, long long* clk) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sA = smem;
  unsigned char* sB = smem + A_BYTES;
  __shared__ uint64_t done, ready;
  __shared__ uint32_t tbase;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const uint4* a_src = a_img + rank * (A_BYTES / 16);
  const uint4* b_src = b_img + rank * (B_BYTES / 16);
  for (int i = tid; i < A_BYTES / 16; i += 128) reinterpret_cast<uint4*>(sA)[i] = a_src[i];
  for (int i = tid; i < B_BYTES / 16; i += 128) reinterpret_cast<uint4*>(sB)[i] = b_src[i];
  if (tid == 0) { mbar_init(&done, 1); mbar_init(&ready, 1); mbar_init_fence(); }
  if (warp == 0) tmem_alloc_pair<64>(&tbase);
  fence_async_smem();
  fence_before_sync();
  cluster_sync_all();
  fence_after_sync();
  const uint32_t tm = tbase;
  long long t0 = 0, t1 = 0;
  if (rank == 1 && tid == 0) mbar_arrive_cluster(map_to_rank(&ready, 0));   // "my operands are in place"
  if (rank == 0 && warp == 0) {
    mbar_wait_cluster(&ready, 0);
    fence_after_sync();
    const uint64_t ad0 = smem_desc_nosw(smem_u32(sA), A_LBO, A_SBO);
    const uint64_t bd0 = smem_desc_nosw(smem_u32(sB), B_LBO, B_SBO);
    const uint32_t idesc = idesc_bf16_f32(M, N);
    t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < reps; ++r) {
      if (elect_one()) {
#pragma unroll
        for (int s = 0; s < KSTEPS; ++s)
          mma_ss2_pair(tm, (uint32_t)ad0 + ((s * 2 * A_LBO) >> 4), (uint32_t)(ad0 >> 32),
                       (uint32_t)bd0 + ((s * 1024) >> 4), (uint32_t)(bd0 >> 32), idesc, (r | s) ? 1u : 0u);
      }
      __syncwarp();
    }
    if (elect_one()) mma_commit_pair(&done);
    __syncwarp();
  }
  mbar_wait(&done, 0);                      // both CTAs: the multicast commit arrives on each CTA's own barrier
  if (rank == 0 && tid == 0) { t1 = clock64(); clk[0] = t1 - t0; }
  fence_after_sync();
  float v[32];
  for (int half = 0; half < 2; ++half) {
    tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + half * 32, v);
    for (int i = 0; i < 32; ++i) out[(rank * 128 + warp * 32 + (tid & 31)) * N + half * 32 + i] = v[i];
  }
  fence_before_sync();
  cluster_sync_all();
  if (warp == 0) tmem_free_pair<64>(tm);
}

int main() {
  unsigned char* a = (unsigned char*)calloc(2 * A_BYTES, 1);
  unsigned char* b = (unsigned char*)calloc(2 * B_BYTES, 1);
  float* fa = (float*)malloc(M * K * 4);
  float* fb = (float*)malloc(N * K * 4);
  srand(2);
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) {
      __nv_bfloat16 h = __float2bfloat16((float)(rand() % 2001 - 1000) / 500.0f);
      fa[m * K + k] = __bfloat162float(h);
      memcpy(a + (m / 128) * A_BYTES + a_off(m % 128, k), &h, 2);
    }
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      __nv_bfloat16 h = __float2bfloat16((float)(rand() % 2001 - 1000) / 700.0f);
      fb[n * K + k] = __bfloat162float(h);
      memcpy(b + (n / 32) * B_BYTES + b_off(n % 32, k), &h, 2);
    }
  uint4 *da, *db;
  float* dout;
  long long* dclk;
  cudaMalloc(&da, 2 * A_BYTES); cudaMalloc(&db, 2 * B_BYTES); cudaMalloc(&dout, M * N * 4); cudaMalloc(&dclk, 8);
  cudaMemcpy(da, a, 2 * A_BYTES, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b, 2 * B_BYTES, cudaMemcpyHostToDevice);
  const int smem = A_BYTES + B_BYTES;
  cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  float* out = (float*)malloc(M * N * 4);
  cudaMemset(dout, 0, M * N * 4);
  k_test<<<2, 128, smem>>>(da, db, dout, 1, dclk);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(out, dout, M * N * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  int bad_m = -1, bad_n = -1;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double r = 0;
      for (int k = 0; k < K; ++k) r += (double)fa[m * K + k] * fb[n * K + k];
      if (fabs(r - out[m * N + n]) > maxerr) { maxerr = fabs(r - out[m * N + n]); bad_m = m; bad_n = n; }
      maxref = fmax(maxref, fabs(r));
    }
  printf("CTA-pair MMA 256x64x48: max |err| = %.3e at (m=%d, n=%d) (max |ref| = %.2f) %s\n", maxerr, bad_m, bad_n, maxref,
         maxerr < 1e-3 ? "OK" : "WRONG");
  if (maxerr < 1e-3) {
    for (int reps : {64, 512}) {
      k_test<<<2, 128, smem>>>(da, db, dout, reps, dclk);
      cudaDeviceSynchronize();
      long long c;
      cudaMemcpy(&c, dclk, 8, cudaMemcpyDeviceToHost);
      printf("%d pair-MMAs (256x64x16): %lld clk = %.1f clk / MMA\n", reps * KSTEPS, c, (double)c / (reps * KSTEPS));
    }
  }
  printf("RESULT %s\n", maxerr < 1e-3 ? "PASS" : "FAIL");
  return maxerr < 1e-3 ? 0 : 2;
}
