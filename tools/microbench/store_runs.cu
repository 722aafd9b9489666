// Microbenchmark: cost of sector-aligned store runs of different lengths on B200.
// Output [32*512*512 pixels][84 floats]; CTA owns a 16x64 tile (as the pooled kernel) and writes every pixel PAIR
// (672 B = 21 sectors) in passes of RUN sectors each (aligned), lanes = floats of the run, then rows of the tile.
// RUN = 1 (32 B), 3 (96 B), 7 (224 B), 21 (672 B: whole pair).  Also varies the number of CTAs (SM-side vs
// memory-side limit) and the store width (STG.32 vs STG.128).
#include <cuda_runtime.h>
#include <stdio.h>
constexpr int H = 512, W = 512, N = 32, CH = 84, TH = 16, TW = 64;
template <int RUN, int VEC>
__global__ void __launch_bounds__(256) k(float* out) {
  constexpr int RF = RUN * 8 / VEC;        // store slots per run
  const int tiles_x = W / TW, tiles_y = H / TH, tid = threadIdx.x;
  for (int t = blockIdx.x; t < N * tiles_x * tiles_y; t += gridDim.x) {
    const int n = t / (tiles_x * tiles_y), r0 = t % (tiles_x * tiles_y);
    const int y0 = (r0 / tiles_x) * TH, x0 = (r0 % tiles_x) * TW;
    float* base = out + (((long long)n * H + y0) * W + x0) * CH;
    const float val = (float)t;
    for (int g = 0; g < 21 / RUN; ++g) {
      for (int it = tid; it < RF * TH * (TW / 2); it += 256) {
        const int f = it % RF, pp = it / RF, r = pp % TH, xp = pp / TH;     // lanes: run slots, then rows
        float* o = base + ((long long)r * W + 2 * xp) * CH + g * RUN * 8 + f * VEC;
        if (VEC == 1) __stcs(o, val); else __stcs(reinterpret_cast<float4*>(o), make_float4(val, val, val, val));
      }
    }
  }
}
template <int RUN, int VEC> void run(float* out, int ctas) {
  const size_t bytes = (size_t)N * H * W * CH * 4;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); k<RUN, VEC><<<ctas, 256>>>(out); cudaEventRecord(e1); cudaEventSynchronize(e1); }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("run %2d sectors (%3d B)  STG.%-3d  CTAs %4d  %.3f ms  %.0f GB/s\n", RUN, RUN * 32, VEC * 32, ctas, ms, bytes / ms / 1e6);
}
int main() {
  float* out; cudaMalloc(&out, (size_t)N * H * W * CH * 4);
  for (int ctas : {148 * 4, 148, 37}) {
    run<1, 1>(out, ctas); run<3, 1>(out, ctas); run<7, 1>(out, ctas); run<21, 1>(out, ctas);
    run<1, 4>(out, ctas); run<3, 4>(out, ctas); run<7, 4>(out, ctas); run<21, 4>(out, ctas);
  }
  return 0;
}
