/*
 * shdr.h -- C ABI of libshdr.so: the B200-native (sm_100a) Linearization-Net
 * per-pixel path of SingleHDR-tf2.
 *
 * The reference (ShinYwings/SingleHDR-tf2) has no FFI/plugin layer: its
 * boundary for this path is the Python call surface of a Keras Model and a
 * utility module.  Each entry point below is the native replacement for one
 * of those Python functions; the Python shims that keep the reference's names
 * and argument meaning live in singlehdr-tf2_b200/layers.py, and the
 * reference-side binding a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / TF types.
 *   - Every tensor is float32, NHWC, compact row-major.
 *   - Device-pointer entry points are ASYNCHRONOUS on `stream` (a cudaStream_t
 *     cast to void*; NULL = legacy default stream) on the device that owns
 *     `out`; the caller owns every buffer and nothing is retained after return.
 *   - Return value: SHDR_OK (0) or a negative SHDR_ERR_* code; the message is
 *     available from shdr_last_error() (thread-local).
 *   - Re-entrant and thread-safe; the only global state is the EMoR table.
 *   - There is NO CPU fallback: without a CUDA device every compute entry
 *     point returns SHDR_ERR_CUDA.
 */
#ifndef SHDR_H_
#define SHDR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SHDR_VERSION 200            /* 0.2.0 */

#define SHDR_OK               0
#define SHDR_ERR_INVALID     -1     /* bad argument (shape, NULL pointer, alignment ...) */
#define SHDR_ERR_CUDA        -2     /* CUDA runtime error (message has cudaGetErrorString) */
#define SHDR_ERR_UNSUPPORTED -3     /* valid request this build does not implement */
#define SHDR_ERR_NOTABLE     -4     /* EMoR table needed but shdr_set_emor_table not called */

#define SHDR_EMOR_SAMPLES   1024    /* linearization_net.py:181 */
#define SHDR_EMOR_NCOMP       11    /* linearization_net.py:182,225 */
#define SHDR_FRONTEND_CH      93    /* 3 img + 6 edge + 12 + 24 + 48, linearization_net.py:322 */
#define SHDR_HIST_CH          84    /* 12 + 24 + 48 */

int         shdr_version(void);
const char* shdr_last_error(void);
/* number of visible CUDA devices (0 and SHDR_OK when the driver is absent) */
int         shdr_device_count(int* count);

/* ---- EMoR inverse table ------------------------------------------------
 * Replaces the per-call text parse inside AEInvcrfDecodeNet.invcrf_pca_w_2_invcrf
 * (linearization_net.py:233 -> parse_invemor :217-227).  g0_host[s], hinv_host is
 * [s][ncomp] row-major (what parse_invemor's np.stack(..., axis=-1) returns).
 * Host copy is kept; each device gets its copy on first use.  s must be 1024 and
 * ncomp 11 in this build. */
int shdr_set_emor_table(const float* g0_host, const float* hinv_host, int s, int ncomp);

/* ---- (A) feature front end ---------------------------------------------- */

/* tf.image.sobel_edges(img) reshaped to [n,h,w,2c]  (linearization_net.py:312-314):
 * REFLECT pad 1, channel = c*2 + k, k=0 dy, k=1 dx.  h,w >= 2.
 * Output pixel p, channel j is written to out[p*out_ch_stride + out_ch_off + j]
 * (out_ch_stride = 2c, out_ch_off = 0 for a stand-alone tensor). */
int shdr_sobel6_f32(const float* img, float* out, int n, int h, int w, int c,
                    int out_ch_stride, int out_ch_off, void* stream);

/* model.histogram_layer(img, bins)  (linearization_net.py:336-350), optionally fused
 * with the 16x16 / stride 1 / 'same' average pool of :351 (pool_k = 0 or 16; the
 * pooled form needs c == 3).  Output channel = (bin-1)*c + ch, written with the same
 * out_ch_stride / out_ch_off rule as above (stride = c*bins for stand-alone). */
int shdr_soft_hist_f32(const float* img, float* out, int n, int h, int w, int c,
                       int bins, int pool_k, int out_ch_stride, int out_ch_off,
                       void* stream);

/* concat(histogram_layer(img,4), (img,8), (img,16)) -> out[n,h,w,84], one launch,
 * pool_k = 0 | 16.  img is [n,h,w,3]. */
int shdr_hist_multi_f32(const float* img, float* out, int n, int h, int w,
                        int pool_k, void* stream);

/* The whole front end of linearization_net.model.call (:312-322):
 * out[n,h,w,93] = concat(img, edge6, hist4, hist8, hist16); pool_k = 0 (as shipped
 * by the reference) or 16 (histograms pooled). */
int shdr_frontend_f32(const float* img, float* out, int n, int h, int w,
                      int pool_k, void* stream);

/* shdr_frontend_f32(pool_k = 0) with the 93-channel tensor rounded to bfloat16 (round to nearest even) on the way
 * out: out_bf16 is [n,h,w,93] of 16-bit values (186 B/px instead of 372).  The reduced-precision output flag of
 * SURVEY.md 8(f) rank 2 for a consumer (crfFeatureNet.conv1, linearization_net.py:91,107) that runs in bf16; it
 * changes numerics and is therefore never what the parity-gated default path or bench.py's headline uses. */
int shdr_frontend_bf16(const float* img, void* out_bf16, int n, int h, int w, void* stream);
/* the same with IEEE half precision (fp16) instead of bfloat16: 11 instead of 8 significand bits for a consumer that
 * runs in TensorFlow's mixed_float16 policy; every feature is within +-4, far inside fp16's range. */
int shdr_frontend_f16(const float* img, void* out_f16, int n, int h, int w, void* stream);

/* ---- front end fused into the input of crfFeatureNet.conv1 --------------
 * Replaces  tf.concat([img, edge6, hist4, hist8, hist16], -1)   (linearization_net.py:322)
 *        ->  crfFeatureNet.conv1 = Conv2D(64, (7,7), strides (2,2), padding 'SAME', bias)
 *            (linearization_net.py:91,107)
 * and optionally the inference-mode norm1 + act1 behind it (:108-109) as a per-channel scale / shift / ReLU.
 * The 93-channel tensor never reaches HBM: each CTA builds the fp16 feature tile of a 16 x 8 block of output
 * pixels in shared memory and contracts it with the 7x7x93x64 kernel on the tensor cores (tcgen05, fp32
 * accumulation).  Half-precision (IEEE fp16) operands change numerics, so this is a separate entry point and never the
 * parity-gated fp32 default: against the fp32 convolution of the fp32 features the error is ~3e-4 of the output's
 * scale (fp16 rather than bf16: every operand is O(1), and 11 instead of 8 significand bits come at the same
 * tensor-core rate; values beyond +-65504 saturate); against the same convolution of fp16-rounded features and
 * weights it is fp32 summation-order noise.
 *
 * shdr_conv1_pack_weights_f32: kernel_hwio is conv1's variable [7][7][93][64] (device, fp32, the layout Keras
 * stores); packed receives shdr_conv1_packed_bytes() bytes: fp16, K in three passes of 32 channels, every tap an
 * operand image the tensor cores read as is -- once for the single-CTA kernel and once split by output-channel half
 * for the CTA-pair kernel (cta_group::2), which is used whenever the input has at least 4 output tiles of 16 x 8.
 * Re-pack whenever the weights change.
 * shdr_frontend_conv1_f32: img [n,h,w,3] -> out [n, ceil(h/2), ceil(w/2), 64] fp32 =
 *   act((conv) * scale[o] + shift[o]);  scale NULL = 1, shift NULL = 0 (shift = the conv bias for a plain
 *   conv1; a folded batch norm gives both), relu != 0 applies max(., 0).  h, w >= 2. */
size_t shdr_conv1_packed_bytes(void);
int shdr_conv1_pack_weights_f32(const float* kernel_hwio, void* packed, void* stream);
int shdr_frontend_conv1_f32(const float* img, const void* packed, const float* scale,
                            const float* shift, int relu, float* out, int n, int h, int w,
                            void* stream);

/* ---- (B) inverse-CRF stage ----------------------------------------------- */

/* AEInvcrfDecodeNet.invcrf_pca_w_2_invcrf (:231-253): curve[b,1024] = g0 + hinv.w[b,11];
 * monotone != 0 additionally applies model._increase (:368-392). */
int shdr_invcrf_build_f32(const float* w, float* curve, int b, int monotone, void* stream);

/* model._increase(rf) (:368-392) on rf[b,k], 2 <= k <= 49152 (the curve lives in shared memory). */
int shdr_increase_f32(const float* rf, float* out, int b, int k, void* stream);

/* tf_utils.apply_rf(x, rf) (tf_utils.py:95-105): x[b, elems_per_item], rf[b,k]. */
int shdr_apply_rf_f32(const float* x, const float* rf, float* y, int b,
                      long long elems_per_item, int k, void* stream);

/* PCA build + _increase + apply_rf back to back on `stream` (what the inference
 * graph does at test_real_refinement.py:94-95).  curve_out[b,1024] is a required
 * scratch/output buffer. */
int shdr_linearize_f32(const float* x, const float* w, float* y, float* curve_out,
                       int b, long long elems_per_item, void* stream);

/* The steps either side of apply_rf in the inference graph, in the same pass
 * (test_real_refinement.py:91-101): x is [b, pixels_per_item, 3] (RGB pixels);
 *   c = clip_by_value(x, 0, 1)                           if clip != 0   (:91)
 *   y = apply_rf(c, rf)                                                  (:95)
 *   alpha = min(1, max(0, max_c(y) - 1 + thr) / thr), tiled over the 3 channels   (:98-101)
 * clipped_out (C_pred, [b,pixels,3]) and alpha_out ([b,pixels,3]) may be NULL. */
int shdr_apply_rf_ex_f32(const float* x, const float* rf, float* y, float* clipped_out,
                         float* alpha_out, int b, long long pixels_per_item, int k,
                         int clip, float thr, void* stream);
/* shdr_linearize_f32 with the same fused neighbours (curve from the PCA weights w[b,11]). */
int shdr_linearize_ex_f32(const float* x, const float* w, float* y, float* curve_out,
                          float* clipped_out, float* alpha_out, int b,
                          long long pixels_per_item, int clip, float thr, void* stream);

/* Synthetic-LDR generator: the per-pixel part of _preprocessing (train.py:28-51, joint_training.py:26-47), one pass:
 *   x = relu(hdr*t[b] + noise_s * (sigma_s[b,c] * hdr*t[b]) + sigma_c[b,c] * noise_c);   c = clip(x, 0, 1);
 *   l = apply_rf(c, crf[b]);   q = round(l * 255)  (half to even, as float)
 * hdr / noise_s / noise_c are [b, pixels_per_item, 3]; the unit-normal noise samples are inputs (the random generator
 * stays with the caller); t is [b], sigma_s / sigma_c are [b,3], crf is [b,k] (forward CRFs; the reference draws them
 * from dorfCurves.txt, which its repository does not ship).  Each output ([b,pixels,3]) may be NULL. */
int shdr_synth_ldr_f32(const float* hdr, const float* t, const float* sigma_s, const float* sigma_c,
                       const float* noise_s, const float* noise_c, const float* crf,
                       float* out_hdr_t, float* out_clipped, float* out_ldr, float* out_quant,
                       int b, long long pixels_per_item, int k, void* stream);

/* ---- (C) reverse-mode gradients (training steps: train.py:186-194,
 *          joint_training.py:156-186, finetune_real_dataset.py:149-178) -------------
 * Each restates what TensorFlow's autodiff computes for the reference's op sequence. */

/* gradient of tf_utils.apply_rf (tf_utils.py:54-105): gy has the shape of x.
 * gx (nullable, shape of x) = gy (k-1) (rf[i1] - rf[i0]);
 * grf (nullable, [b,k], overwritten) = scatter-add of gy (y1-y) at i0 and gy (y-y0) at i1
 * (fp32 atomics: the summation order is not deterministic, like TF's gather_nd gradient). */
int shdr_apply_rf_bwd_f32(const float* x, const float* rf, const float* gy, float* gx,
                          float* grf, int b, long long elems_per_item, int k, void* stream);
/* gradient of model._increase (linearization_net.py:368-392) w.r.t. rf[b,k], 2 <= k <= 24576. */
int shdr_increase_bwd_f32(const float* rf, const float* gout, float* grf, int b, int k,
                          void* stream);
/* gradient of shdr_invcrf_build_f32 w.r.t. w[b,11]: gcurve[b,1024] is the gradient of the
 * curve (monotone = 0: of g0 + hinv.w; monotone != 0: of _increase(g0 + hinv.w)). */
int shdr_invcrf_build_bwd_f32(const float* w, const float* gcurve, float* gw, int b,
                              int monotone, void* stream);
/* gradient of shdr_frontend_f32(pool_k = 0) w.r.t. img: gfeat[n,h,w,93] -> gimg[n,h,w,3]
 * (identity slice + soft-histogram slopes + transposed REFLECT Sobel). */
int shdr_frontend_bwd_f32(const float* img, const float* gfeat, float* gimg, int n, int h,
                          int w, void* stream);
/* gradient of shdr_soft_hist_f32(pool_k = 0, dense output) w.r.t. img: ghist[n,h,w,c*bins]. */
int shdr_soft_hist_bwd_f32(const float* img, const float* ghist, float* gimg, int n, int h,
                           int w, int c, int bins, void* stream);

/* ---- DLPack entry points ------------------------------------------------------
 * Same operations on DLManagedTensor* (DLPack v0.x ABI, as produced by
 * tf.experimental.dlpack.to_dlpack / torch.utils.dlpack.to_dlpack).  Inputs are
 * validated (kDLCUDA, float32, lanes 1, compact row-major or NULL strides) and never
 * written or consumed (the caller still owns the capsule).  Outputs are allocated by
 * the library on the inputs' device and returned as a new DLManagedTensor whose
 * deleter frees the device memory -- wrap it in a "dltensor" capsule and hand it to
 * from_dlpack. */
struct DLManagedTensor;
int shdr_dl_frontend(const struct DLManagedTensor* img, int pool_k, void* stream,
                     struct DLManagedTensor** out);
int shdr_dl_sobel6(const struct DLManagedTensor* img, void* stream,
                   struct DLManagedTensor** out);
/* shdr_frontend_bf16 on DLPack tensors: the result is a kDLBfloat / 16-bit tensor [n,h,w,93] */
int shdr_dl_frontend_bf16(const struct DLManagedTensor* img, void* stream,
                          struct DLManagedTensor** out);
/* shdr_frontend_f16 on DLPack tensors: the result is a kDLFloat / 16-bit tensor [n,h,w,93] */
int shdr_dl_frontend_f16(const struct DLManagedTensor* img, void* stream,
                         struct DLManagedTensor** out);
int shdr_dl_soft_hist(const struct DLManagedTensor* img, int bins, int pool_k,
                      void* stream, struct DLManagedTensor** out);
int shdr_dl_invcrf_build(const struct DLManagedTensor* w, int monotone, void* stream,
                         struct DLManagedTensor** out);
int shdr_dl_increase(const struct DLManagedTensor* rf, void* stream,
                     struct DLManagedTensor** out);
int shdr_dl_apply_rf(const struct DLManagedTensor* x, const struct DLManagedTensor* rf,
                     void* stream, struct DLManagedTensor** out);
/* allocate an uninitialised float32 CUDA tensor (for callers that build their own) */
int shdr_dl_alloc_f32(const int64_t* shape, int ndim, int device,
                      struct DLManagedTensor** out);
/* run the deleter of a tensor returned above that was never handed to a consumer */
void shdr_dl_release(struct DLManagedTensor* t);
/* Producer/consumer ordering of a library-owned tensor without a device-wide sync: every
 * shdr_dl_* op records a "ready" event on its stream; shdr_dl_mark_ready does the same for
 * work the caller enqueued itself.  shdr_dl_wait_ready makes `consumer_stream` wait for it
 * (on_host = 0) or blocks the host until the tensor's producer has finished (on_host != 0). */
int shdr_dl_mark_ready(struct DLManagedTensor* t, void* stream);
int shdr_dl_wait_ready(struct DLManagedTensor* t, void* consumer_stream, int on_host);
/* address of a `void (*)(PyObject*)` usable as PyCapsule destructor for a "dltensor" capsule:
 * it runs the tensor's deleter unless a consumer renamed the capsule ("used_dltensor"). */
void* shdr_dl_capsule_destructor(void);

/* ---- TF-free / torch-free device helpers (tests, bench, host-buffer API) ------- */
int shdr_malloc(void** p, size_t bytes, int device);
int shdr_free(void* p, int device);
int shdr_malloc_host(void** p, size_t bytes);              /* pinned */
int shdr_free_host(void* p);
int shdr_memset(void* p, int value, size_t bytes, int device, void* stream);
int shdr_h2d(void* dst_dev, const void* src_host, size_t bytes, int device, void* stream);
int shdr_d2h(void* dst_host, const void* src_dev, size_t bytes, int device, void* stream);
int shdr_sync(int device);                                 /* cudaDeviceSynchronize */
int shdr_stream_create(void** stream, int device);         /* non-blocking stream */
int shdr_stream_destroy(void* stream, int device);
int shdr_stream_sync(void* stream, int device);
int shdr_event_create(void** event, int device);
int shdr_event_destroy(void* event, int device);
int shdr_event_record(void* event, void* stream, int device);
int shdr_event_elapsed_ms(void* start, void* stop, float* ms);   /* syncs on stop */
int shdr_stream_wait_event(void* stream, void* event, int device);
/* number of kernels this library has launched in this process (all threads) */
long long shdr_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SHDR_H_ */
