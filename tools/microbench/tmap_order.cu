// Does cuTensorMapEncodeTiled accept a dimension order whose strides are not increasing (row before pixel-pair)?
// And does a TMA store with that map write the right bytes?  Output viewed as (float, row, pair, image).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <vector>
constexpr int H = 64, W = 128, N = 2, CH = 84;
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap tmap, int sector, int y0, int pair0, int n) {
  extern __shared__ __align__(128) float sm[];   // [32 pairs][16 rows][8 floats]
  for (int i = threadIdx.x; i < 32 * 16 * 8; i += blockDim.x) sm[i] = (float)i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(sm);
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
                 ::"l"(&tmap), "r"(sector * 8), "r"(y0), "r"(pair0), "r"(n), "r"(s) : "memory");
    asm volatile("cp.async.bulk.commit_group;\n\tcp.async.bulk.wait_group 0;" ::: "memory");
  }
}
int main() {
  const size_t nfl = (size_t)N * H * W * CH;
  float* out; cudaMalloc(&out, nfl * 4); cudaMemset(out, 0, nfl * 4);
  void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  CUtensorMap tmap;
  cuuint64_t gd[4] = {2 * CH, H, W / 2, N};
  cuuint64_t gs[3] = {(cuuint64_t)W * CH * 4, 2 * CH * 4, (cuuint64_t)H * W * CH * 4};
  cuuint32_t bx[4] = {8, 16, 32, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult rc = ((EncodeFn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, out, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode (float,row,pair,image) rc=%d\n", (int)rc);
  if (rc != CUDA_SUCCESS) return 0;
  // store sector 13 of the tile at rows 56.., pairs 48.. of image 1 (rows 56..71 -> clipped at 64, pairs 48..79 -> clipped at 64)
  k<<<1, 256, 32 * 16 * 8 * 4>>>(tmap, 13, 56, 48, 1);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  std::vector<float> hst(nfl); cudaMemcpy(hst.data(), out, nfl * 4, cudaMemcpyDeviceToHost);
  size_t bad = 0, written = 0;
  for (int n = 0; n < N; ++n) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) for (int c = 0; c < CH; ++c) {
    const float v = hst[(((size_t)n * H + y) * W + x) * CH + c];
    const int pf = (x & 1) * CH + c;             // float index within the pixel pair
    const int pair = x / 2;
    float want = 0.f;
    if (n == 1 && y >= 56 && pair >= 48 && pf >= 104 && pf < 112) want = (float)((((pair - 48) * 16) + (y - 56)) * 8 + (pf - 104));
    if (v != want) ++bad;
    if (v != 0.f) ++written;
  }
  printf("written %zu (expect %d), mismatches %zu\n", written, 8 * 16 * 8 - 1, bad);
  return 0;
}
