// Microbenchmark: how fast can B200 absorb the pooled kernel's output pattern?
// Output [n*h*w pixels][84 floats]; a CTA owns a 16x64-pixel tile and writes it group by group
// (7 groups x 12 channels = 48-byte fragments at a 336-byte pixel pitch), like k_hist_pooled.
//   mode 0: STG.32, lanes = 12 channels x rows (current kernel mapping), 16 stores per thread along x
//   mode 1: STG.32, lanes = 12 channels x adjacent pixels
//   mode 2: STG.128, 3 lanes per 48-byte fragment, adjacent pixels
//   mode 3: TMA tensor store of a [16][64][12] box from shared memory
//   mode 4: full contiguous float4 stores (upper bound; tile rows written as whole 21.5 KB runs)
//   mode 5: STG.32, all 84 channels of a pixel by adjacent lanes (what a 84-channel-at-once kernel could do)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>

constexpr int H = 512, W = 512, N = 32, CH = 84, TH = 16, TW = 64;

__global__ void __launch_bounds__(256) k_modes(float* out, int mode, const __grid_constant__ CUtensorMap tmap,
                                               const __grid_constant__ CUtensorMap tmap_pair, int run) {
  extern __shared__ __align__(128) float sm[];
  const int tiles_x = W / TW, tiles_y = H / TH;
  const int tid = threadIdx.x;
  for (int t = blockIdx.x; t < N * tiles_x * tiles_y; t += gridDim.x) {
    const int n = t / (tiles_x * tiles_y);
    const int r0 = t % (tiles_x * tiles_y);
    const int y0 = (r0 / tiles_x) * TH, x0 = (r0 % tiles_x) * TW;
    float* base = out + (((long long)n * H + y0) * W + x0) * CH;
    const float val = (float)t;
    if (mode == 4) {
      for (int r = 0; r < TH; ++r) {
        float4* row = reinterpret_cast<float4*>(base + (long long)r * W * CH);
        for (int i = tid; i < TW * CH / 4; i += 256) __stcs(row + i, make_float4(val, val, val, val));
      }
      continue;
    }
    if (mode == 5) {
      for (int i = tid; i < TH * TW * CH; i += 256) {
        const int r = i / (TW * CH);
        base[(long long)r * W * CH + (i - r * TW * CH)] = val;
      }
      continue;
    }
    if (mode == 6 || mode == 7) {   // prefill the tile with full-line stores so the fragment writes hit in L2
      for (int r = 0; r < TH; ++r) {
        float4* row = reinterpret_cast<float4*>(base + (long long)r * W * CH);
        for (int i = tid; i < TW * CH / 4; i += 256) row[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      __syncthreads();
    }
    if (mode == 11) {
      // TMA stores of whole sectors: output viewed as [rows][w/2 pixel pairs][168 floats]; box = (8*run floats,
      // 32 pairs, 16 rows): `run` aligned sectors of every pixel pair of the tile per bulk store, 21/run passes
      for (int g = 0; g < 21 / run; ++g) {
        for (int i = tid; i < TH * (TW / 2) * 8 * run; i += 256) sm[i] = val;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
          const unsigned s = (unsigned)__cvta_generic_to_shared(sm);
          asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                       ::"l"(&tmap_pair), "r"(g * 8 * run), "r"(x0 / 2), "r"(n * H + y0), "r"(s) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncthreads();
      }
      continue;
    }
    if (mode == 10) {
      // 21 passes; pass k writes ONE aligned 32-B sector (sector k of 21) of every pixel pair;
      // lanes = 8 floats x 4 rows, each thread walks 8 pixel pairs along x (the planned 8-channel-unit consumer)
      for (int g = 0; g < 21; ++g) {
        for (int it = tid; it < 8 * TH * 4; it += 256) {
          const int xb = it / (8 * TH), line = it % (8 * TH), r = line / 8, f = line % 8;
          float* o = base + ((long long)r * W + xb * 16) * CH + g * 8 + f;
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j * 2 * CH] = val;
        }
      }
      continue;
    }
    if (mode == 8 || mode == 9) {
      // 7 passes; pass k writes sectors [3k, 3k+3) (96 B, 32-B aligned) of every PIXEL PAIR (672 B = 21 sectors)
      for (int g = 0; g < 7; ++g) {
        if (mode == 8) {        // lanes = 24 floats of a fragment, then adjacent pairs
          for (int it = tid; it < 24 * TH * (TW / 2); it += 256) {
            const int f = it % 24, pp = it / 24, r = pp / (TW / 2), xp = pp % (TW / 2);
            base[((long long)r * W + 2 * xp) * CH + g * 24 + f] = val;
          }
        } else {                // lanes = 24 floats x rows, 8 stores per thread along x (pairs)
          for (int it = tid; it < 24 * TH * 4; it += 256) {
            const int xb = it / (24 * TH), line = it % (24 * TH), r = line / 24, f = line % 24;
            float* o = base + ((long long)r * W + xb * 16) * CH + g * 24 + f;
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j * 2 * CH] = val;
          }
        }
      }
      continue;
    }
    for (int g = 0; g < 7; ++g) {
      float* gb = base + g * 12;
      if (mode == 0 || mode == 6) {
        for (int it = tid; it < 12 * TH * 4; it += 256) {
          const int xb = it / (12 * TH), line = it % (12 * TH), r = line / 12, ch = line % 12;
          float* o = gb + ((long long)r * W + xb * 16) * CH + ch;
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j * CH] = val;
        }
      } else if (mode == 1) {
        for (int it = tid; it < 12 * TH * TW; it += 256) {
          const int ch = it % 12, p = it / 12, r = p / TW, x = p % TW;
          gb[((long long)r * W + x) * CH + ch] = val;
        }
      } else if (mode == 2 || mode == 7) {
        for (int it = tid; it < 3 * TH * TW; it += 256) {
          const int q = it % 3, p = it / 3, r = p / TW, x = p % TW;
          *reinterpret_cast<float4*>(gb + ((long long)r * W + x) * CH + q * 4) = make_float4(val, val, val, val);
        }
      } else if (mode == 3) {
        for (int i = tid; i < TH * TW * 12; i += 256) sm[i] = val;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
          const unsigned s = (unsigned)__cvta_generic_to_shared(sm);
          asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                       ::"l"(&tmap), "r"(g * 12), "r"(x0), "r"(n * H + y0), "r"(s) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncthreads();
      }
    }
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const size_t bytes = (size_t)N * H * W * CH * 4;
  float* out; cudaMalloc(&out, bytes);
  cudaMemset(out, 0, bytes);
  CUtensorMap tmap;
  void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  cuuint64_t gdim[3] = {CH, W, (cuuint64_t)N * H};
  cuuint64_t gstr[2] = {CH * 4, (cuuint64_t)W * CH * 4};
  cuuint32_t box[3] = {12, TW, TH};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult rc = ((EncodeFn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, gdim, gstr, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("tensor map encode rc=%d\n", (int)rc);
  CUtensorMap tmap_pair[3];
  int runs[3] = {1, 3, 7};
  for (int i = 0; i < 3; ++i) {
    cuuint64_t gd[3] = {2 * CH, W / 2, (cuuint64_t)N * H};
    cuuint64_t gs[2] = {2 * CH * 4, (cuuint64_t)W * CH * 4};
    cuuint32_t bx[3] = {(cuuint32_t)(8 * runs[i]), TW / 2, TH};
    CUresult r2 = ((EncodeFn)fn)(&tmap_pair[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, gd, gs, bx, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                 CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("pair tensor map (run %d) rc=%d\n", runs[i], (int)r2);
  }
  cudaFuncSetAttribute(k_modes, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
  const char* names[11] = {"STG.32 ch x rows (current)", "STG.32 ch x adjacent px", "STG.128 3 lanes/fragment", "TMA box [16][64][12]",
                          "full-line float4 (bound)", "STG.32 84 ch of adjacent px", "prefill full lines + mode 0", "prefill full lines + mode 2", "aligned 96B/pixel-pair, adj pairs", "aligned 96B/pixel-pair, ch x rows", "aligned 32B sectors, 8ch x 4 rows"};
  for (int mode = 0; mode < 11; ++mode) {
    if (mode == 3 && rc != CUDA_SUCCESS) continue;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      k_modes<<<148 * 4, 256, TH * TW * 12 * 4>>>(out, mode, tmap, tmap_pair[0], 1);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    printf("mode %d %-32s %.3f ms  %.0f GB/s  %s\n", mode, names[mode], ms, bytes / ms / 1e6, cudaGetErrorString(err));
  }
  for (int i = 0; i < 3; ++i)
    for (int ctas : {148, 296}) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        k_modes<<<ctas, 256, TH * (TW / 2) * 8 * runs[i] * 4>>>(out, 11, tmap, tmap_pair[i], runs[i]);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("mode 11 TMA aligned %3d-byte runs per pixel pair, %3d CTAs  %.3f ms  %.0f GB/s  %s\n", 32 * runs[i], ctas, ms,
             bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
  float hv[4]; cudaMemcpy(hv, out + 84 * 100 + 13, 4, cudaMemcpyDeviceToHost); printf("sample %.1f\n", hv[0]);
  return 0;
}
