// api.cu -- C-ABI plumbing of libshdr: errors, device guard, EMoR table, DLPack entry points,
// and the TF-free / torch-free device helpers used by the tests, the bench and the host-buffer API.
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace shdr {

// ------------------------------------------------------------------ errors
static thread_local char t_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  cudaGetLastError();   // clear the (non-sticky) error so that it is not reported again by a later call
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return SHDR_ERR_CUDA;
}

// ------------------------------------------------------------------ device guard
DeviceGuard::DeviceGuard(const void* p) {
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, p);
  if (e != cudaSuccess) { status = cuda_fail(e, "cudaPointerGetAttributes"); return; }
  if (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged) {
    set_error("pointer %p is not device memory (cudaMemoryType %d): this library has no CPU path", p, (int)at.type);
    status = SHDR_ERR_INVALID;
    return;
  }
  dev = at.device;
  if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
  if (prev != dev) {
    e = cudaSetDevice(dev);
    if (e != cudaSuccess) status = cuda_fail(e, "cudaSetDevice");
  }
}
DeviceGuard::DeviceGuard(int device) : dev(device) {
  if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
  if (prev != dev) {
    cudaError_t e = cudaSetDevice(dev);
    if (e != cudaSuccess) status = cuda_fail(e, "cudaSetDevice");
  }
}
DeviceGuard::~DeviceGuard() {
  if (prev >= 0 && prev != dev) cudaSetDevice(prev);
}

int sm_count(int dev) {
  static std::mutex mu;
  static std::vector<int> cache;
  std::lock_guard<std::mutex> lk(mu);
  if ((int)cache.size() <= dev) cache.resize(dev + 1, 0);
  if (cache[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cache[dev] = v;
  }
  return cache[dev];
}

// ------------------------------------------------------------------ EMoR table
static std::mutex g_tab_mu;
static std::vector<float> g_tab_host;          // g0[1024] then hinv[1024][11]
static unsigned g_tab_version = 0;
struct DevTab { float* ptr = nullptr; unsigned version = 0; };
static std::vector<DevTab> g_tab_dev;

int emor_device_table(int dev, const float** g0, const float** hinv) {
  std::lock_guard<std::mutex> lk(g_tab_mu);
  if (g_tab_host.empty()) {
    set_error("EMoR table not set: call shdr_set_emor_table (parse_invemor) first");
    return SHDR_ERR_NOTABLE;
  }
  if ((int)g_tab_dev.size() <= dev) g_tab_dev.resize(dev + 1);
  DevTab& t = g_tab_dev[dev];
  const size_t bytes = g_tab_host.size() * sizeof(float);
  if (t.version != g_tab_version) {
    // a changed table goes to a NEW buffer: kernels already enqueued on other streams keep reading the old one
    // (it is leaked on purpose -- 49 KB per table change, which happens once per process in practice)
    SHDR_CUDA(cudaMalloc((void**)&t.ptr, bytes));
    // one-time 49 KB upload per device (g0, then hinv TRANSPOSED to [11][1024] so that the curve kernel's
    // per-sample reads are coalesced); synchronous so the host vector may change afterwards
    std::vector<float> dev_img(g_tab_host.size());
    memcpy(dev_img.data(), g_tab_host.data(), SHDR_EMOR_SAMPLES * sizeof(float));
    for (int s = 0; s < SHDR_EMOR_SAMPLES; ++s)
      for (int j = 0; j < SHDR_EMOR_NCOMP; ++j)
        dev_img[SHDR_EMOR_SAMPLES + (size_t)j * SHDR_EMOR_SAMPLES + s] =
            g_tab_host[SHDR_EMOR_SAMPLES + (size_t)s * SHDR_EMOR_NCOMP + j];
    SHDR_CUDA(cudaMemcpy(t.ptr, dev_img.data(), bytes, cudaMemcpyHostToDevice));
    t.version = g_tab_version;
  }
  *g0 = t.ptr;
  *hinv = t.ptr + SHDR_EMOR_SAMPLES;
  return SHDR_OK;
}

}  // namespace shdr

// ------------------------------------------------------------------ DLPack (v0.x ABI)
extern "C" {
typedef struct { int32_t device_type; int32_t device_id; } DLDevice;
typedef struct { uint8_t code; uint8_t bits; uint16_t lanes; } DLDataType;
typedef struct {
  void* data; DLDevice device; int32_t ndim; DLDataType dtype;
  int64_t* shape; int64_t* strides; uint64_t byte_offset;
} DLTensor;
struct DLManagedTensor {
  DLTensor dl_tensor;
  void* manager_ctx;
  void (*deleter)(struct DLManagedTensor*);
};
}

namespace shdr {
enum { kDLCUDA = 2, kDLCUDAManaged = 13, kDLFloat = 2, kDLBfloat = 4 };

struct View { const float* p = nullptr; int dev = 0; int ndim = 0; int64_t shape[8] = {0}; int64_t numel = 1; };

static int dl_view(const DLManagedTensor* m, const char* who, View* v) {
  SHDR_REQUIRE(m != nullptr, "%s: NULL DLManagedTensor", who);
  const DLTensor& t = m->dl_tensor;
  SHDR_REQUIRE(t.device.device_type == kDLCUDA || t.device.device_type == kDLCUDAManaged,
               "%s: tensor is on DLPack device type %d, need kDLCUDA (no CPU path)", who, t.device.device_type);
  SHDR_REQUIRE(t.dtype.code == kDLFloat && t.dtype.bits == 32 && t.dtype.lanes == 1,
               "%s: dtype (code %d, bits %d, lanes %d) is not float32", who, t.dtype.code, t.dtype.bits, t.dtype.lanes);
  SHDR_REQUIRE(t.ndim >= 1 && t.ndim <= 8, "%s: ndim=%d", who, t.ndim);
  int64_t numel = 1;
  for (int i = 0; i < t.ndim; ++i) { SHDR_REQUIRE(t.shape[i] >= 0, "%s: negative dim", who); numel *= t.shape[i]; }
  if (t.strides != nullptr && numel > 0) {
    int64_t expect = 1;
    for (int i = t.ndim - 1; i >= 0; --i) {
      SHDR_REQUIRE(t.shape[i] == 1 || t.strides[i] == expect, "%s: tensor is not compact row-major (dim %d stride %lld)",
                   who, i, (long long)t.strides[i]);
      expect *= t.shape[i];
    }
  }
  v->p = reinterpret_cast<const float*>(static_cast<const char*>(t.data) + t.byte_offset);
  v->dev = t.device.device_id;
  v->ndim = t.ndim;
  for (int i = 0; i < t.ndim; ++i) v->shape[i] = t.shape[i];
  v->numel = numel;
  return SHDR_OK;
}

struct OwnedTensor {
  DLManagedTensor m;
  int64_t shape[8];
  int dev;
  cudaEvent_t ready;      // recorded after the producing work; nullptr until the first mark
  bool pooled;            // allocated with cudaMallocAsync (stream-ordered pool) rather than cudaMalloc
};
// The deleter runs when the DLPack consumer drops the tensor.  The memory goes back to the device's stream-ordered
// pool with cudaFreeAsync on the legacy default stream: no device-wide synchronisation (a plain cudaFree costs
// milliseconds per op), and ordered after everything already enqueued on the legacy stream and on every blocking
// stream.  Consumers with private non-blocking streams hold their own reference until their kernels have run
// (TensorFlow's GPU device keeps input buffers referenced until the consuming kernels complete), which is the
// contract every DLPack producer with a memory pool relies on.
static void owned_deleter(DLManagedTensor* m) {
  OwnedTensor* o = static_cast<OwnedTensor*>(m->manager_ctx);
  DeviceGuard g(o->dev);
  if (o->ready) cudaEventDestroy(o->ready);
  if (o->m.dl_tensor.data) {
    if (o->pooled) cudaFreeAsync(o->m.dl_tensor.data, (cudaStream_t)0);
    else cudaFree(o->m.dl_tensor.data);
  }
  delete o;
}
static OwnedTensor* owned_of(DLManagedTensor* t) {
  return (t && t->deleter == owned_deleter) ? static_cast<OwnedTensor*>(t->manager_ctx) : nullptr;
}
static int dl_alloc(const int64_t* shape, int ndim, int dev, cudaStream_t st, DLManagedTensor** out,
                    bool stream_ordered = true, int half_kind = 0 /* 1: bfloat16, 2: fp16 */) {
  SHDR_REQUIRE(out != nullptr && ndim >= 1 && ndim <= 8, "dl_alloc: bad arguments");
  int64_t numel = 1;
  for (int i = 0; i < ndim; ++i) { SHDR_REQUIRE(shape[i] >= 0, "dl_alloc: negative dim"); numel *= shape[i]; }
  OwnedTensor* o = new OwnedTensor();
  memset(&o->m, 0, sizeof(o->m));
  o->dev = dev;
  o->ready = nullptr;
  o->pooled = stream_ordered;
  void* p = nullptr;
  {
    DeviceGuard g(dev);
    if (g.status != SHDR_OK) { delete o; return g.status; }
    // stream-ordered allocation from the device's default pool: no implicit device synchronisation, and blocks freed
    // by cudaFree are reused without going back to the driver
    const size_t bytes = (size_t)(numel > 0 ? numel : 1) * (half_kind ? 2 : sizeof(float));
    cudaError_t e = stream_ordered ? cudaMallocAsync(&p, bytes, st) : cudaMalloc(&p, bytes);
    if (e != cudaSuccess) { delete o; return cuda_fail(e, "cudaMalloc(output tensor)"); }
  }
  for (int i = 0; i < ndim; ++i) o->shape[i] = shape[i];
  DLTensor& t = o->m.dl_tensor;
  t.data = p;
  t.device.device_type = kDLCUDA;
  t.device.device_id = dev;
  t.ndim = ndim;
  t.dtype.code = half_kind == 1 ? kDLBfloat : kDLFloat; t.dtype.bits = half_kind ? 16 : 32; t.dtype.lanes = 1;
  t.shape = o->shape;
  t.strides = nullptr;
  t.byte_offset = 0;
  o->m.manager_ctx = o;
  o->m.deleter = owned_deleter;
  *out = &o->m;
  return SHDR_OK;
}
static int dl_mark_ready(DLManagedTensor* t, cudaStream_t st) {
  OwnedTensor* o = owned_of(t);
  SHDR_REQUIRE(o != nullptr, "dl_mark_ready: not a tensor allocated by this library");
  DeviceGuard g(o->dev);
  if (g.status != SHDR_OK) return g.status;
  if (!o->ready) SHDR_CUDA(cudaEventCreateWithFlags(&o->ready, cudaEventDisableTiming));
  SHDR_CUDA(cudaEventRecord(o->ready, st));
  return SHDR_OK;
}

}  // namespace shdr

using namespace shdr;

// ================================================================== C ABI
extern "C" int shdr_version(void) { return SHDR_VERSION; }
extern "C" const char* shdr_last_error(void) { return t_err; }
extern "C" long long shdr_launch_count(void) { return g_launches.load(); }

extern "C" int shdr_device_count(int* count) {
  SHDR_REQUIRE(count != nullptr, "device_count: NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
  *count = n;
  return SHDR_OK;
}

extern "C" int shdr_set_emor_table(const float* g0_host, const float* hinv_host, int s, int ncomp) {
  SHDR_REQUIRE(g0_host && hinv_host, "set_emor_table: NULL pointer");
  SHDR_REQUIRE(s == SHDR_EMOR_SAMPLES && ncomp == SHDR_EMOR_NCOMP,
               "set_emor_table: this build needs s=%d, ncomp=%d (got %d, %d)", SHDR_EMOR_SAMPLES, SHDR_EMOR_NCOMP, s, ncomp);
  std::lock_guard<std::mutex> lk(g_tab_mu);
  const size_t n = (size_t)s * (1 + ncomp);
  if (g_tab_host.size() == n && memcmp(g_tab_host.data(), g0_host, (size_t)s * sizeof(float)) == 0 &&
      memcmp(g_tab_host.data() + s, hinv_host, (size_t)s * ncomp * sizeof(float)) == 0)
    return SHDR_OK;                      // same table again (the reference re-parses it on every call): nothing to do
  g_tab_host.resize(n);
  memcpy(g_tab_host.data(), g0_host, (size_t)s * sizeof(float));
  memcpy(g_tab_host.data() + s, hinv_host, (size_t)s * ncomp * sizeof(float));
  ++g_tab_version;
  return SHDR_OK;
}

// ---- DLPack entry points
static int img_dims(const View& v, const char* who, int* n, int* h, int* w, int* c) {
  SHDR_REQUIRE(v.ndim == 4, "%s: img must be [n,h,w,c] (ndim=%d)", who, v.ndim);
  for (int i = 0; i < 4; ++i) SHDR_REQUIRE(v.shape[i] <= 0x7fffffff, "%s: dim %d too large", who, i);
  *n = (int)v.shape[0]; *h = (int)v.shape[1]; *w = (int)v.shape[2]; *c = (int)v.shape[3];
  return SHDR_OK;
}

#define DL_TRY(expr) do { int _rc = (expr); if (_rc != SHDR_OK) return _rc; } while (0)
#define DL_RUN(outp, expr) do { int _rc = (expr); if (_rc != SHDR_OK) { shdr_dl_release(*(outp)); *(outp) = nullptr; return _rc; } } while (0)

extern "C" void shdr_dl_release(struct DLManagedTensor* t) { if (t && t->deleter) t->deleter(t); }

// PyCapsule destructor for "dltensor" capsules made by the Python side.  The CPython symbols are
// resolved from the running interpreter at first use, so libshdr has no link-time Python dependency.
#include <dlfcn.h>
static void capsule_destructor(void* pyobj) {
  typedef int (*is_valid_t)(void*, const char*);
  typedef void* (*get_ptr_t)(void*, const char*);
  static is_valid_t is_valid = (is_valid_t)dlsym(RTLD_DEFAULT, "PyCapsule_IsValid");
  static get_ptr_t get_ptr = (get_ptr_t)dlsym(RTLD_DEFAULT, "PyCapsule_GetPointer");
  if (!is_valid || !get_ptr) return;
  if (is_valid(pyobj, "dltensor")) {           // never consumed: we still own the tensor
    DLManagedTensor* m = (DLManagedTensor*)get_ptr(pyobj, "dltensor");
    if (m && m->deleter) m->deleter(m);
  }
}
extern "C" void* shdr_dl_capsule_destructor(void) { return (void*)&capsule_destructor; }

extern "C" int shdr_dl_mark_ready(struct DLManagedTensor* t, void* stream) {
  return dl_mark_ready(t, (cudaStream_t)stream);
}
extern "C" int shdr_dl_wait_ready(struct DLManagedTensor* t, void* consumer_stream, int on_host) {
  OwnedTensor* o = owned_of(t);
  SHDR_REQUIRE(o != nullptr, "dl_wait_ready: not a tensor allocated by this library");
  DeviceGuard g(o->dev);
  if (g.status != SHDR_OK) return g.status;
  if (!o->ready) {                       // producer unknown (caller enqueued work without marking): be conservative
    SHDR_CUDA(cudaDeviceSynchronize());
    return SHDR_OK;
  }
  if (on_host) SHDR_CUDA(cudaEventSynchronize(o->ready));
  else SHDR_CUDA(cudaStreamWaitEvent((cudaStream_t)consumer_stream, o->ready, 0));
  return SHDR_OK;
}

extern "C" int shdr_dl_alloc_f32(const int64_t* shape, int ndim, int device, struct DLManagedTensor** out) {
  SHDR_REQUIRE(shape != nullptr, "dl_alloc: NULL shape");
  return dl_alloc(shape, ndim, device, (cudaStream_t)0, out, false);   // the caller's stream is unknown: plain cudaMalloc
}

extern "C" int shdr_dl_frontend(const struct DLManagedTensor* img, int pool_k, void* stream, struct DLManagedTensor** out) {
  View v; int n, h, w, c;
  DL_TRY(dl_view(img, "frontend", &v));
  DL_TRY(img_dims(v, "frontend", &n, &h, &w, &c));
  SHDR_REQUIRE(c == 3, "frontend: img must have 3 channels (got %d)", c);
  int64_t shp[4] = {n, h, w, SHDR_FRONTEND_CH};
  DL_TRY(dl_alloc(shp, 4, v.dev, (cudaStream_t)stream, out));
  DL_RUN(out, shdr_frontend_f32(v.p, (float*)(*out)->dl_tensor.data, n, h, w, pool_k, stream));
  DL_RUN(out, dl_mark_ready(*out, (cudaStream_t)stream));
  return SHDR_OK;
}

extern "C" int shdr_dl_frontend_bf16(const struct DLManagedTensor* img, void* stream, struct DLManagedTensor** out) {
  View v; int n, h, w, c;
  DL_TRY(dl_view(img, "frontend_bf16", &v));
  DL_TRY(img_dims(v, "frontend_bf16", &n, &h, &w, &c));
  SHDR_REQUIRE(c == 3, "frontend_bf16: img must have 3 channels (got %d)", c);
  int64_t shp[4] = {n, h, w, SHDR_FRONTEND_CH};
  DL_TRY(dl_alloc(shp, 4, v.dev, (cudaStream_t)stream, out, true, 1));
  DL_RUN(out, shdr_frontend_bf16(v.p, (*out)->dl_tensor.data, n, h, w, stream));
  DL_RUN(out, dl_mark_ready(*out, (cudaStream_t)stream));
  return SHDR_OK;
}

extern "C" int shdr_dl_frontend_f16(const struct DLManagedTensor* img, void* stream, struct DLManagedTensor** out) {
  View v; int n, h, w, c;
  DL_TRY(dl_view(img, "frontend_f16", &v));
  DL_TRY(img_dims(v, "frontend_f16", &n, &h, &w, &c));
  SHDR_REQUIRE(c == 3, "frontend_f16: img must have 3 channels (got %d)", c);
  int64_t shp[4] = {n, h, w, SHDR_FRONTEND_CH};
  DL_TRY(dl_alloc(shp, 4, v.dev, (cudaStream_t)stream, out, true, 2));
  DL_RUN(out, shdr_frontend_f16(v.p, (*out)->dl_tensor.data, n, h, w, stream));
  DL_RUN(out, dl_mark_ready(*out, (cudaStream_t)stream));
  return SHDR_OK;
}

extern "C" int shdr_dl_sobel6(const struct DLManagedTensor* img, void* stream, struct DLManagedTensor** out) {
  View v; int n, h, w, c;
  DL_TRY(dl_view(img, "sobel6", &v));
  DL_TRY(img_dims(v, "sobel6", &n, &h, &w, &c));
  int64_t shp[4] = {n, h, w, 2 * (int64_t)c};
  DL_TRY(dl_alloc(shp, 4, v.dev, (cudaStream_t)stream, out));
  DL_RUN(out, shdr_sobel6_f32(v.p, (float*)(*out)->dl_tensor.data, n, h, w, c, 2 * c, 0, stream));
  DL_RUN(out, dl_mark_ready(*out, (cudaStream_t)stream));
  return SHDR_OK;
}

extern "C" int shdr_dl_soft_hist(const struct DLManagedTensor* img, int bins, int pool_k, void* stream,
                                 struct DLManagedTensor** out) {
  View v; int n, h, w, c;
  DL_TRY(dl_view(img, "soft_hist", &v));
  DL_TRY(img_dims(v, "soft_hist", &n, &h, &w, &c));
  SHDR_REQUIRE(bins >= 1 && bins <= 4096, "soft_hist: bins=%d (need 1..4096)", bins);
  int64_t shp[4] = {n, h, w, (int64_t)c * bins};
  DL_TRY(dl_alloc(shp, 4, v.dev, (cudaStream_t)stream, out));
  DL_RUN(out, shdr_soft_hist_f32(v.p, (float*)(*out)->dl_tensor.data, n, h, w, c, bins, pool_k, c * bins, 0, stream));
  DL_RUN(out, dl_mark_ready(*out, (cudaStream_t)stream));
  return SHDR_OK;
}

extern "C" int shdr_dl_invcrf_build(const struct DLManagedTensor* w, int monotone, void* stream,
                                    struct DLManagedTensor** out) {
  View v;
  DL_TRY(dl_view(w, "invcrf_build", &v));
  SHDR_REQUIRE(v.ndim == 2 && v.shape[1] == SHDR_EMOR_NCOMP, "invcrf_build: w must be [b,%d]", SHDR_EMOR_NCOMP);
  int64_t shp[2] = {v.shape[0], SHDR_EMOR_SAMPLES};
  DL_TRY(dl_alloc(shp, 2, v.dev, (cudaStream_t)stream, out));
  DL_RUN(out, shdr_invcrf_build_f32(v.p, (float*)(*out)->dl_tensor.data, (int)v.shape[0], monotone, stream));
  DL_RUN(out, dl_mark_ready(*out, (cudaStream_t)stream));
  return SHDR_OK;
}

extern "C" int shdr_dl_increase(const struct DLManagedTensor* rf, void* stream, struct DLManagedTensor** out) {
  View v;
  DL_TRY(dl_view(rf, "increase", &v));
  SHDR_REQUIRE(v.ndim == 2, "increase: rf must be [b,k]");
  int64_t shp[2] = {v.shape[0], v.shape[1]};
  DL_TRY(dl_alloc(shp, 2, v.dev, (cudaStream_t)stream, out));
  DL_RUN(out, shdr_increase_f32(v.p, (float*)(*out)->dl_tensor.data, (int)v.shape[0], (int)v.shape[1], stream));
  DL_RUN(out, dl_mark_ready(*out, (cudaStream_t)stream));
  return SHDR_OK;
}

extern "C" int shdr_dl_apply_rf(const struct DLManagedTensor* x, const struct DLManagedTensor* rf, void* stream,
                                struct DLManagedTensor** out) {
  View vx, vr;
  DL_TRY(dl_view(x, "apply_rf(x)", &vx));
  DL_TRY(dl_view(rf, "apply_rf(rf)", &vr));
  SHDR_REQUIRE(vr.ndim == 2, "apply_rf: rf must be [b,k]");
  SHDR_REQUIRE(vx.ndim >= 1 && vx.shape[0] == vr.shape[0], "apply_rf: batch of x (%lld) != batch of rf (%lld)",
               (long long)vx.shape[0], (long long)vr.shape[0]);
  SHDR_REQUIRE(vx.dev == vr.dev, "apply_rf: x is on device %d, rf on device %d", vx.dev, vr.dev);
  SHDR_REQUIRE(vr.shape[1] >= 1 && vr.shape[1] <= 0x7fffffff, "apply_rf: k out of range");
  const int b = (int)vx.shape[0];
  const long long per = b > 0 ? vx.numel / b : 0;
  DL_TRY(dl_alloc(vx.shape, vx.ndim, vx.dev, (cudaStream_t)stream, out));
  DL_RUN(out, shdr_apply_rf_f32(vx.p, vr.p, (float*)(*out)->dl_tensor.data, b, per, (int)vr.shape[1], stream));
  DL_RUN(out, dl_mark_ready(*out, (cudaStream_t)stream));
  return SHDR_OK;
}

// ---- device helpers
extern "C" int shdr_malloc(void** p, size_t bytes, int device) {
  SHDR_REQUIRE(p != nullptr, "malloc: NULL");
  DeviceGuard g(device);
  if (g.status != SHDR_OK) return g.status;
  SHDR_CUDA(cudaMalloc(p, bytes ? bytes : 1));
  return SHDR_OK;
}
extern "C" int shdr_free(void* p, int device) {
  if (!p) return SHDR_OK;
  DeviceGuard g(device);
  if (g.status != SHDR_OK) return g.status;
  SHDR_CUDA(cudaFree(p));
  return SHDR_OK;
}
extern "C" int shdr_malloc_host(void** p, size_t bytes) {
  SHDR_REQUIRE(p != nullptr, "malloc_host: NULL");
  SHDR_CUDA(cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocPortable));
  return SHDR_OK;
}
extern "C" int shdr_free_host(void* p) {
  if (!p) return SHDR_OK;
  SHDR_CUDA(cudaFreeHost(p));
  return SHDR_OK;
}
extern "C" int shdr_memset(void* p, int value, size_t bytes, int device, void* stream) {
  DeviceGuard g(device);
  if (g.status != SHDR_OK) return g.status;
  SHDR_CUDA(cudaMemsetAsync(p, value, bytes, (cudaStream_t)stream));
  return SHDR_OK;
}
extern "C" int shdr_h2d(void* dst, const void* src, size_t bytes, int device, void* stream) {
  DeviceGuard g(device);
  if (g.status != SHDR_OK) return g.status;
  SHDR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return SHDR_OK;
}
extern "C" int shdr_d2h(void* dst, const void* src, size_t bytes, int device, void* stream) {
  DeviceGuard g(device);
  if (g.status != SHDR_OK) return g.status;
  SHDR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return SHDR_OK;
}
extern "C" int shdr_sync(int device) {
  DeviceGuard g(device);
  if (g.status != SHDR_OK) return g.status;
  SHDR_CUDA(cudaDeviceSynchronize());
  return SHDR_OK;
}
extern "C" int shdr_stream_create(void** stream, int device) {
  SHDR_REQUIRE(stream != nullptr, "stream_create: NULL");
  DeviceGuard g(device);
  if (g.status != SHDR_OK) return g.status;
  cudaStream_t s;
  SHDR_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  *stream = (void*)s;
  return SHDR_OK;
}
extern "C" int shdr_stream_destroy(void* stream, int device) {
  DeviceGuard g(device);
  if (g.status != SHDR_OK) return g.status;
  SHDR_CUDA(cudaStreamDestroy((cudaStream_t)stream));
  return SHDR_OK;
}
extern "C" int shdr_stream_sync(void* stream, int device) {
  DeviceGuard g(device);
  if (g.status != SHDR_OK) return g.status;
  SHDR_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return SHDR_OK;
}
extern "C" int shdr_event_create(void** event, int device) {
  SHDR_REQUIRE(event != nullptr, "event_create: NULL");
  DeviceGuard g(device);
  if (g.status != SHDR_OK) return g.status;
  cudaEvent_t e;
  SHDR_CUDA(cudaEventCreate(&e));
  *event = (void*)e;
  return SHDR_OK;
}
extern "C" int shdr_event_destroy(void* event, int device) {
  DeviceGuard g(device);
  if (g.status != SHDR_OK) return g.status;
  SHDR_CUDA(cudaEventDestroy((cudaEvent_t)event));
  return SHDR_OK;
}
extern "C" int shdr_event_record(void* event, void* stream, int device) {
  DeviceGuard g(device);
  if (g.status != SHDR_OK) return g.status;
  SHDR_CUDA(cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream));
  return SHDR_OK;
}
extern "C" int shdr_event_elapsed_ms(void* start, void* stop, float* ms) {
  SHDR_REQUIRE(ms != nullptr, "event_elapsed_ms: NULL");
  SHDR_CUDA(cudaEventSynchronize((cudaEvent_t)stop));
  SHDR_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
  return SHDR_OK;
}
extern "C" int shdr_stream_wait_event(void* stream, void* event, int device) {
  DeviceGuard g(device);
  if (g.status != SHDR_OK) return g.status;
  SHDR_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)event, 0));
  return SHDR_OK;
}
