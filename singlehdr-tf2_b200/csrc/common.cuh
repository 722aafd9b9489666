// common.cuh -- shared host/device helpers for libshdr (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>

#include "../../include/shdr.h"

namespace shdr {

// ---- error plumbing (thread-local message, no exceptions across the ABI) ----
void set_error(const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what);
extern std::atomic<long long> g_launches;

#define SHDR_CUDA(expr)                                                        \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess) return ::shdr::cuda_fail(_e, #expr);                \
  } while (0)

#define SHDR_REQUIRE(cond, ...)                                                \
  do {                                                                         \
    if (!(cond)) { ::shdr::set_error(__VA_ARGS__); return SHDR_ERR_INVALID; }  \
  } while (0)

// after a <<<>>> launch
#define SHDR_LAUNCH_CHECK(name)                                                \
  do {                                                                         \
    ::shdr::g_launches.fetch_add(1, std::memory_order_relaxed);                \
    cudaError_t _e = cudaGetLastError();                                       \
    if (_e != cudaSuccess) return ::shdr::cuda_fail(_e, name);                 \
  } while (0)

// Sets the device that owns `p` for the lifetime of the guard.
struct DeviceGuard {
  int prev = -1;
  int dev = -1;
  int status = SHDR_OK;
  explicit DeviceGuard(const void* p);
  explicit DeviceGuard(int device);
  ~DeviceGuard();
};

int sm_count(int dev);

// EMoR table on the current device ([0..1023] g0, then hinv TRANSPOSED [11][1024]); error if unset.
int emor_device_table(int dev, const float** g0, const float** hinv);

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

#ifdef __CUDACC__
// ---- device helpers ---------------------------------------------------------
// streaming (evict-first) 128-bit global accesses: every tensor on this path is
// touched exactly once, so keep it from displacing the small reused data in L2/L1.
__device__ __forceinline__ float4 ld_stream4(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream4(float4* p, float4 v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream1(float* p, float v) { __stcs(p, v); }

// One soft-histogram vote, op-for-op what linearization_net.py:345-346 does:
//   d = |v - centre| ; h = (d < thr) ? 1 - d*B : 0      (sub, abs, less, mul, sub, select)
// __fmul_rn/__fsub_rn forbid FMA contraction so a non-power-of-two B rounds like TF.
__device__ __forceinline__ float hist_vote(float v, float centre, float thr, float nbins) {
  float d = fabsf(__fsub_rn(v, centre));
  float h = __fsub_rn(1.0f, __fmul_rn(d, nbins));
  return d < thr ? h : 0.0f;
}

// Power-of-two B only: d*B and 1/B are exact, so (d < 1/B) <=> (1 - d*B > 0) and the
// single-rounding FMA equals mul-then-sub bit for bit; NaN -> 0 on both forms.
__device__ __forceinline__ float hist_vote_pow2(float v, float centre, float nbins) {
  return fmaxf(fmaf(-fabsf(__fsub_rn(v, centre)), nbins, 1.0f), 0.0f);
}

// REFLECT index for a pad of 1: -1 -> 1, n -> n-2   (n >= 2)
__device__ __forceinline__ int reflect1(int i, int n) {
  return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i);
}
#endif

}  // namespace shdr
