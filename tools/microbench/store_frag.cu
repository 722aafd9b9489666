// Microbenchmark: store throughput vs fragment size.  Output [pixels][84 floats]; a CTA owns a 16x64 tile and
// writes it in 84/G passes of G channels each (G*4-byte fragments at a 336-byte pitch).
//   map 0: lanes = channels-in-group fastest, then adjacent pixels (one STG.32 per element)
//   map 1: lanes = channels-in-group fastest, then rows; 16 stores per thread along x (current kernel's mapping)
#include <cuda_runtime.h>
#include <stdio.h>
constexpr int H = 512, W = 512, N = 32, CH = 84, TH = 16, TW = 64;
__global__ void __launch_bounds__(256) k(float* out, int G, int map, int order) {
  const int tiles_x = W / TW, tiles_y = H / TH;
  const int tid = threadIdx.x;
  for (int t = blockIdx.x; t < N * tiles_x * tiles_y; t += gridDim.x) {
    const int n = t / (tiles_x * tiles_y);
    const int r0 = t % (tiles_x * tiles_y);
    const int y0 = (r0 / tiles_x) * TH, x0 = (r0 % tiles_x) * TW;
    float* base = out + (((long long)n * H + y0) * W + x0) * CH;
    const float val = (float)t;
    for (int g = 0; g < CH / G; ++g) {
      float* gb = base + g * G;
      if (map == 0) {
        for (int it = tid; it < G * TH * TW; it += 256) {
          const int ch = it % G, p = it / G, r = p / TW, x = p % TW;
          gb[((long long)r * W + x) * CH + ch] = val;
        }
      } else {
        for (int it = tid; it < G * TH * 4; it += 256) {
          const int xb = it / (G * TH), line = it % (G * TH), r = line / G, ch = line % G;
          float* o = gb + ((long long)r * W + xb * 16) * CH + ch;
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j * CH] = val;
        }
      }
      if (order) __syncthreads();
    }
  }
}
int main() {
  const size_t bytes = (size_t)N * H * W * CH * 4;
  float* out; cudaMalloc(&out, bytes);
  int Gs[7] = {4, 12, 21, 28, 42, 84, 84};
  for (int map = 0; map < 2; ++map)
    for (int gi = 0; gi < 6; ++gi) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        k<<<148 * 4, 256>>>(out, Gs[gi], map, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("map %d  G=%2d (%3d B fragments)  %.3f ms  %.0f GB/s\n", map, Gs[gi], Gs[gi] * 4, ms, bytes / ms / 1e6);
    }
  return 0;
}
