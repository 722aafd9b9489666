"""Host-side mirror of the reference's call surface for the Linearization-Net
per-pixel path -- same names, argument meaning and layouts as

    linearization_net.model.histogram_layer(img, max_bin)        linearization_net.py:336
    tf.image.sobel_edges(img) + reshape to 6 channels            linearization_net.py:312-314
    concat([img, edge, hist4, hist8, hist16], -1)                linearization_net.py:322
    AEInvcrfDecodeNet.parse_invemor()                            linearization_net.py:217
    AEInvcrfDecodeNet.invcrf_pca_w_2_invcrf(w)                   linearization_net.py:231
    model._increase(rf)                                          linearization_net.py:369
    tf_utils.apply_rf(x, rf)                                     tf_utils.py:95

Inputs are CUDA float32 tensors handed over zero-copy by DLPack (anything with
``__dlpack__``: TF eager tensors via ``tf.experimental.dlpack``, torch tensors,
:class:`~.device.DeviceArray`); outputs are :class:`~.device.DeviceArray` objects that any
DLPack consumer (``tf.experimental.dlpack.from_dlpack``, ``torch.from_dlpack``) takes
without a copy.  All arithmetic happens in libshdr's sm_100a kernels; nothing here
computes on the host and nothing falls back to it.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _native as N
from .device import Borrowed, DeviceArray

BINS = (4, 8, 16)                 # linearization_net.py:322
POOL_K = 16                       # linearization_net.py:351


def _stream(stream):
    if stream is None:
        return None
    return getattr(stream, "handle", stream)


def _run_dl(fn, ins, *args, stream=None):
    """Call a shdr_dl_* entry: borrowed inputs first, then scalar args, stream, &out.

    Stream contract: the kernel is enqueued on ``stream`` (``None`` = the legacy default stream), so the INPUTS must
    already be ordered before that stream -- true for producers on the legacy stream (torch by default) and for
    :class:`DeviceArray` inputs, whose ready event is waited for here; a framework with private non-blocking streams
    (TensorFlow) must synchronise first, which ``tf_adapter`` does.  The borrowed capsules are released when this
    returns, i.e. possibly before the kernel has run: callers that may free the inputs right away keep them alive
    until the output is ready (``tf_adapter`` holds them until its host-side wait)."""
    borrowed = [Borrowed(x) for x in ins]
    sh = _stream(stream)
    for b, x in zip(borrowed, ins):
        if isinstance(x, DeviceArray):
            N.check(N.lib.shdr_dl_wait_ready(b.managed, sh, 0))
    out = C.c_void_p()
    N.check(fn(*[b.managed for b in borrowed], *args, sh, C.byref(out)))
    return DeviceArray(out.value)


def _dev_inputs(stream, *xs):
    """Borrow inputs for a raw-pointer entry point; DeviceArray inputs make ``stream`` wait for their producer."""
    bs = [Borrowed(x) for x in xs]
    sh = _stream(stream)
    for b, x in zip(bs, xs):
        if isinstance(x, DeviceArray):
            N.check(N.lib.shdr_dl_wait_ready(b.managed, sh, 0))
    return bs


# --------------------------------------------------------------------------- front end
def sobel_edges6(img, stream=None):
    """``tf.reshape(tf.image.sobel_edges(img), [b,h,w,2c])``: channel ``c*2+k``, k=0 dy, k=1 dx."""
    return _run_dl(N.lib.shdr_dl_sobel6, [img], stream=stream)


def histogram_layer(img, max_bin, pool=False, stream=None):
    """Spatial-aware soft histogram, ``[b,h,w,c] -> [b,h,w,c*max_bin]`` (bin-major channels).

    ``pool=True`` fuses the optional 16x16 / stride 1 / 'same' average pool of
    linearization_net.py:351 (TF border-count semantics); the reference ships with it off.
    """
    return _run_dl(N.lib.shdr_dl_soft_hist, [img], int(max_bin), POOL_K if pool else 0, stream=stream)


def frontend(img, pool=False, stream=None):
    """The 93-channel tensor the reference feeds to ``crfFeatureNet``:
    ``concat([img, edge6, hist4, hist8, hist16], -1)`` in one pass over ``img``."""
    return _run_dl(N.lib.shdr_dl_frontend, [img], POOL_K if pool else 0, stream=stream)


def frontend_bf16(img, stream=None):
    """:func:`frontend` (un-pooled) with the 93-channel tensor rounded to bfloat16 on the way out -- the
    reduced-precision output flag for a consumer that runs ``crfFeatureNet.conv1`` in bf16 (half the bytes written and
    handed over).  Changes numerics (8 mantissa bits): never the default, never what the parity tests gate."""
    return _run_dl(N.lib.shdr_dl_frontend_bf16, [img], stream=stream)


def conv1_pack_weights(kernel, stream=None):
    """Pack ``crfFeatureNet.conv1``'s kernel ``[7,7,93,64]`` (fp32 CUDA tensor, HWIO as Keras stores it,
    linearization_net.py:91) into the fp16 operand image :func:`frontend_conv1` streams through the tensor cores.
    Re-pack whenever the variable changes."""
    (bk,) = _dev_inputs(stream, kernel)
    if tuple(bk.shape) != (7, 7, N.FRONTEND_CH, 64):
        raise ValueError(f"conv1_pack_weights: kernel must be [7,7,93,64], got {bk.shape}")
    out = DeviceArray.empty((int(N.lib.shdr_conv1_packed_bytes()) // 4,), bk.device)
    N.check(N.lib.shdr_conv1_pack_weights_f32(bk.ptr, out.ptr, _stream(stream)))
    return out.mark_ready(stream)


def frontend_conv1(img, packed, bias=None, scale=None, relu=False, stream=None):
    """``crfFeatureNet.conv1(tf.concat([img, edge6, hist4, hist8, hist16], -1))`` in ONE kernel
    (linearization_net.py:312-322 -> :107): ``[n,h,w,3] -> [n, ceil(h/2), ceil(w/2), 64]`` fp32, the 93-channel tensor
    never touching HBM (tensor cores, fp16 operands, fp32 accumulation -- reduced precision, opt-in).
    ``packed`` comes from :func:`conv1_pack_weights`; ``bias [64]`` is conv1's bias; ``scale [64]`` / ``relu`` fold an
    inference-mode ``norm1`` / ``act1`` (:108-109): ``out = act(conv * scale + bias)``."""
    ins = [img, packed] + [a for a in (scale, bias) if a is not None]
    bs = _dev_inputs(stream, *ins)
    bi, bp = bs[0], bs[1]
    rest = bs[2:]
    bsc = rest.pop(0) if scale is not None else None
    bb = rest.pop(0) if bias is not None else None
    if len(bi.shape) != 4 or bi.shape[3] != 3:
        raise ValueError(f"frontend_conv1: img must be [n,h,w,3], got {bi.shape}")
    if int(np.prod(bp.shape, dtype=np.int64)) * 4 != int(N.lib.shdr_conv1_packed_bytes()):
        raise ValueError("frontend_conv1: `packed` is not the output of conv1_pack_weights")
    for nm, b_ in (("scale", bsc), ("bias", bb)):
        if b_ is not None and tuple(b_.shape) != (64,):
            raise ValueError(f"frontend_conv1: {nm} must be [64], got {b_.shape}")
    n, h, w, _ = bi.shape
    out = DeviceArray.empty((n, (h + 1) // 2, (w + 1) // 2, 64), bi.device)
    N.check(N.lib.shdr_frontend_conv1_f32(bi.ptr, bp.ptr, bsc.ptr if bsc else None, bb.ptr if bb else None,
                                          1 if relu else 0, out.ptr, n, h, w, _stream(stream)))
    return out.mark_ready(stream)


def frontend_f16(img, stream=None):
    """:func:`frontend_bf16` with IEEE half precision (fp16) instead of bfloat16: exactly ``float16(fp32 result)``, for a
    consumer running in TensorFlow's ``mixed_float16`` policy (``numpy()`` returns ``float16``)."""
    return _run_dl(N.lib.shdr_dl_frontend_f16, [img], stream=stream)


def hist_multi(img, pool=False, stream=None):
    """``concat([hist4, hist8, hist16], -1)`` -> ``[b,h,w,84]`` in one launch."""
    (b,) = _dev_inputs(stream, img)
    if len(b.shape) != 4 or b.shape[3] != 3:
        raise ValueError(f"hist_multi: img must be [n,h,w,3], got {b.shape}")
    n, h, w, _ = b.shape
    out = DeviceArray.empty((n, h, w, N.HIST_CH), b.device)
    N.check(N.lib.shdr_hist_multi_f32(b.ptr, out.ptr, n, h, w, POOL_K if pool else 0, _stream(stream)))
    return out.mark_ready(stream)


# --------------------------------------------------------------------------- EMoR table
_table_cache = {}
_registered_key = None            # the table file currently installed in libshdr (by parse_invemor)


def _read_block(lines, tag):
    try:
        start = lines.index(tag) + 1
    except ValueError:
        raise ValueError(f"EMoR table: tag line {tag!r} not found") from None
    tokens = " ".join(lines[start:start + N.EMOR_SAMPLES // 4]).split()
    if len(tokens) != N.EMOR_SAMPLES:
        raise ValueError(f"EMoR table: expected {N.EMOR_SAMPLES} values after {tag!r}, got {len(tokens)}")
    return np.array(tokens, dtype=np.float32)


def parse_invemor(path="invemor.txt", register=True):
    """Read ``invemor.txt`` (CWD-relative by default, as the reference does) and return
    ``(B[1024], g0[1024], hinv[1024,11])`` float32.  Unlike the reference, which re-parses
    the 440 KB text on every ``invcrf_pca_w_2_invcrf`` call, the result is cached per file
    and (``register=True``) installed as libshdr's device-resident table."""
    st = os.stat(path)
    key = (os.path.abspath(path), st.st_mtime_ns, st.st_size)
    hit = _table_cache.get(key)
    if hit is None:
        with open(path, "r") as f:
            lines = [ln.strip() for ln in f]
        b = _read_block(lines, "B =")
        g0 = _read_block(lines, "g0 =")
        hinv = np.stack([_read_block(lines, f"hinv({i + 1})=") for i in range(N.EMOR_NCOMP)], axis=-1)
        hit = _table_cache[key] = (b, g0, np.ascontiguousarray(hinv))
    global _registered_key
    if register and _registered_key != key:      # the reference re-parses on every call; re-register only on a change
        set_emor_table(hit[1], hit[2])
        _registered_key = key
    return hit


def set_emor_table(g0, hinv):
    """Install ``g0[1024]`` / ``hinv[1024,11]`` (host arrays) as the table the kernels use."""
    g0 = np.ascontiguousarray(g0, dtype=np.float32)
    hinv = np.ascontiguousarray(hinv, dtype=np.float32)
    if g0.shape != (N.EMOR_SAMPLES,) or hinv.shape != (N.EMOR_SAMPLES, N.EMOR_NCOMP):
        raise ValueError(f"EMoR table must be g0[1024], hinv[1024,11]; got {g0.shape}, {hinv.shape}")
    global _registered_key
    _registered_key = None                       # an explicit table overrides whatever parse_invemor installed
    N.check(N.lib.shdr_set_emor_table(g0.ctypes.data, hinv.ctypes.data, N.EMOR_SAMPLES, N.EMOR_NCOMP))


# --------------------------------------------------------------------------- inverse CRF
def invcrf_pca_w_2_invcrf(invcrf_pca_w, stream=None):
    """``[b,11] -> [b,1024]``: ``g0 + hinv . w`` (table from :func:`parse_invemor`)."""
    return _run_dl(N.lib.shdr_dl_invcrf_build, [invcrf_pca_w], 0, stream=stream)


def invcrf_build(invcrf_pca_w, monotone=True, stream=None):
    """PCA reconstruction and (``monotone``) ``_increase`` in one launch."""
    return _run_dl(N.lib.shdr_dl_invcrf_build, [invcrf_pca_w], 1 if monotone else 0, stream=stream)


def _increase(rf, stream=None):
    """Monotonic enforcement: diff, shift by relu(-min), normalise, cumsum, left-pad 0."""
    return _run_dl(N.lib.shdr_dl_increase, [rf], stream=stream)


def apply_rf(x, rf, stream=None):
    """``x [b, s...]``, ``rf [b,k]`` -> per-element linear-interpolated lookup, shape of ``x``."""
    return _run_dl(N.lib.shdr_dl_apply_rf, [x, rf], stream=stream)


def linearize(x, invcrf_pca_w, stream=None):
    """``apply_rf(x, _increase(invcrf_pca_w_2_invcrf(w)))`` back to back on one stream
    (linearization_net.py:325-328 + test_real_refinement.py:95).  Returns ``(y, curve)``."""
    bx, bw = _dev_inputs(stream, x, invcrf_pca_w)
    if len(bw.shape) != 2 or bw.shape[1] != N.EMOR_NCOMP or not bx.shape or bx.shape[0] != bw.shape[0]:
        raise ValueError(f"linearize: x {bx.shape} / w {bw.shape} mismatch (w must be [b,11])")
    b = bx.shape[0]
    per = int(np.prod(bx.shape[1:], dtype=np.int64))
    y = DeviceArray.empty(bx.shape, bx.device)
    curve = DeviceArray.empty((b, N.EMOR_SAMPLES), bx.device)
    N.check(N.lib.shdr_linearize_f32(bx.ptr, bw.ptr, y.ptr, curve.ptr, b, per, _stream(stream)))
    return y.mark_ready(stream), curve.mark_ready(stream)


def linearize_ex(x, invcrf_pca_w=None, rf=None, clip=True, alpha_threshold=None, want_clipped=False, stream=None):
    """The inference graph around ``apply_rf`` in ONE pass over the image (test_real_refinement.py:91-101)::

        C_pred = tf.clip_by_value(pred_deq, 0, 1)                       # clip=True
        B_pred = tf_utils.apply_rf(C_pred, pred_invcrf)
        alpha  = min(1, max(0, reduce_max(B_pred, 3) - 1 + thr) / thr)  # alpha_threshold=thr, tiled to 3 channels

    ``x`` is ``[b,h,w,3]``.  The curve comes either from the PCA weights ``invcrf_pca_w [b,11]`` (build + ``_increase``
    fused in front, as :func:`linearize`) or from ``rf [b,k]``.  Returns a dict with ``y`` (B_pred) and, when
    requested, ``clipped`` (C_pred), ``alpha`` ``[b,h,w,3]`` and ``curve``."""
    if (invcrf_pca_w is None) == (rf is None):
        raise ValueError("linearize_ex: give exactly one of invcrf_pca_w and rf")
    bx, bc = _dev_inputs(stream, x, invcrf_pca_w if rf is None else rf)
    if len(bx.shape) < 2 or bx.shape[-1] != 3 or len(bc.shape) != 2 or bc.shape[0] != bx.shape[0]:
        raise ValueError(f"linearize_ex: x {bx.shape} must be [b,...,3] and the curve input [b,*], got {bc.shape}")
    if alpha_threshold is not None and not alpha_threshold > 0:
        raise ValueError("linearize_ex: alpha_threshold must be > 0")
    b = bx.shape[0]
    npx = int(np.prod(bx.shape[1:-1], dtype=np.int64))
    sh = _stream(stream)
    out = {"y": DeviceArray.empty(bx.shape, bx.device)}
    cl = DeviceArray.empty(bx.shape, bx.device) if want_clipped else None
    al = DeviceArray.empty(bx.shape, bx.device) if alpha_threshold is not None else None
    thr = float(alpha_threshold) if alpha_threshold is not None else 1.0
    if rf is None:
        if bc.shape[1] != N.EMOR_NCOMP:
            raise ValueError(f"linearize_ex: invcrf_pca_w must be [b,11], got {bc.shape}")
        out["curve"] = DeviceArray.empty((b, N.EMOR_SAMPLES), bx.device)
        N.check(N.lib.shdr_linearize_ex_f32(bx.ptr, bc.ptr, out["y"].ptr, out["curve"].ptr, cl.ptr if cl else None,
                                            al.ptr if al else None, b, npx, 1 if clip else 0, thr, sh))
    else:
        N.check(N.lib.shdr_apply_rf_ex_f32(bx.ptr, bc.ptr, out["y"].ptr, cl.ptr if cl else None,
                                           al.ptr if al else None, b, npx, bc.shape[1], 1 if clip else 0, thr, sh))
    if cl is not None:
        out["clipped"] = cl
    if al is not None:
        out["alpha"] = al
    for v in out.values():
        v.mark_ready(stream)
    return out


def synth_ldr(hdr, t, sigma_s, sigma_c, noise_s, noise_c, crf, outputs=("hdr_t", "clipped", "ldr", "quant"), stream=None):
    """The per-pixel part of the reference's synthetic-LDR generator ``_preprocessing`` (train.py:28-51,
    joint_training.py:26-47) in one pass: exposure ``hdr*t``, Poisson + Gaussian noise, relu, clip to [0,1], forward
    CRF (``apply_rf``), 8-bit quantisation.  ``hdr``/``noise_s``/``noise_c``: ``[b,h,w,3]`` (unit-normal samples are
    inputs: the random generator stays with the caller); ``t [b]``; ``sigma_s``/``sigma_c [b,3]``; ``crf [b,k]``.
    Returns a dict with the requested outputs (``hdr_t``, ``clipped``, ``ldr``, ``quant`` = ``round(ldr*255)``)."""
    bh, bt, bss, bsc, bns, bnc, bcrf = _dev_inputs(stream, hdr, t, sigma_s, sigma_c, noise_s, noise_c, crf)
    b = bh.shape[0] if bh.shape else 0
    if (len(bh.shape) < 2 or bh.shape[-1] != 3 or tuple(bns.shape) != tuple(bh.shape) or tuple(bnc.shape) != tuple(bh.shape)
            or int(np.prod(bt.shape)) != b or tuple(bss.shape) != (b, 3) or tuple(bsc.shape) != (b, 3)
            or len(bcrf.shape) != 2 or bcrf.shape[0] != b):
        raise ValueError(f"synth_ldr: shape mismatch: hdr {bh.shape}, t {bt.shape}, sigma {bss.shape}/{bsc.shape}, "
                         f"noise {bns.shape}/{bnc.shape}, crf {bcrf.shape}")
    names = ("hdr_t", "clipped", "ldr", "quant")
    bad = [o for o in outputs if o not in names]
    if bad or not outputs:
        raise ValueError(f"synth_ldr: outputs must be a non-empty subset of {names}, got {outputs}")
    out = {o: DeviceArray.empty(bh.shape, bh.device) for o in outputs}
    ptr = [out[nm].ptr if nm in out else None for nm in names]
    npx = int(np.prod(bh.shape[1:-1], dtype=np.int64))
    N.check(N.lib.shdr_synth_ldr_f32(bh.ptr, bt.ptr, bss.ptr, bsc.ptr, bns.ptr, bnc.ptr, bcrf.ptr, *ptr, b, npx,
                                     bcrf.shape[1], _stream(stream)))
    for v in out.values():
        v.mark_ready(stream)
    return out


# --------------------------------------------------------------------------- gradients (training steps)
def apply_rf_bwd(x, rf, gy, need_gx=True, need_grf=True, stream=None):
    """Gradient of :func:`apply_rf`: returns ``(gx, grf)`` (``None`` for the one not asked for).
    ``gx = gy (k-1) (rf[i1] - rf[i0])``; ``grf[b]`` = scatter-add of ``gy (y1-y)`` at ``i0`` and ``gy (y-y0)`` at ``i1``
    -- what TF's autodiff gives for tf_utils.py:54-105 (train.py:186-194)."""
    bx, br, bg = _dev_inputs(stream, x, rf, gy)
    if len(br.shape) != 2 or not bx.shape or bx.shape[0] != br.shape[0] or tuple(bg.shape) != tuple(bx.shape):
        raise ValueError(f"apply_rf_bwd: x {bx.shape}, rf {br.shape}, gy {bg.shape} mismatch")
    b, k = br.shape
    per = int(np.prod(bx.shape[1:], dtype=np.int64))
    gx = DeviceArray.empty(bx.shape, bx.device) if need_gx else None
    grf = DeviceArray.empty((b, k), bx.device) if need_grf else None
    N.check(N.lib.shdr_apply_rf_bwd_f32(bx.ptr, br.ptr, bg.ptr, gx.ptr if gx else None, grf.ptr if grf else None,
                                        b, per, k, _stream(stream)))
    return (gx.mark_ready(stream) if gx else None), (grf.mark_ready(stream) if grf else None)


def _increase_bwd(rf, gout, stream=None):
    """Gradient of :func:`_increase` w.r.t. ``rf [b,k]`` (linearization_net.py:368-392)."""
    br, bg = _dev_inputs(stream, rf, gout)
    if len(br.shape) != 2 or tuple(bg.shape) != tuple(br.shape):
        raise ValueError(f"_increase_bwd: rf {br.shape} / gout {bg.shape} mismatch")
    out = DeviceArray.empty(br.shape, br.device)
    N.check(N.lib.shdr_increase_bwd_f32(br.ptr, bg.ptr, out.ptr, br.shape[0], br.shape[1], _stream(stream)))
    return out.mark_ready(stream)


def invcrf_build_bwd(invcrf_pca_w, gcurve, monotone=False, stream=None):
    """Gradient of :func:`invcrf_pca_w_2_invcrf` (``monotone=False``) or :func:`invcrf_build` (PCA + ``_increase``)
    w.r.t. ``w [b,11]``; ``gcurve`` is ``[b,1024]``."""
    bw, bg = _dev_inputs(stream, invcrf_pca_w, gcurve)
    if len(bw.shape) != 2 or bw.shape[1] != N.EMOR_NCOMP or tuple(bg.shape) != (bw.shape[0], N.EMOR_SAMPLES):
        raise ValueError(f"invcrf_build_bwd: w {bw.shape} / gcurve {bg.shape} mismatch")
    out = DeviceArray.empty(bw.shape, bw.device)
    N.check(N.lib.shdr_invcrf_build_bwd_f32(bw.ptr, bg.ptr, out.ptr, bw.shape[0], 1 if monotone else 0, _stream(stream)))
    return out.mark_ready(stream)


def frontend_bwd(img, gfeat, stream=None):
    """Gradient of :func:`frontend` (un-pooled) w.r.t. ``img``: ``gfeat [n,h,w,93] -> [n,h,w,3]``."""
    bi, bg = _dev_inputs(stream, img, gfeat)
    if len(bi.shape) != 4 or bi.shape[3] != 3 or tuple(bg.shape) != tuple(bi.shape[:3]) + (N.FRONTEND_CH,):
        raise ValueError(f"frontend_bwd: img {bi.shape} / gfeat {bg.shape} mismatch")
    n, h, w, _ = bi.shape
    out = DeviceArray.empty(bi.shape, bi.device)
    N.check(N.lib.shdr_frontend_bwd_f32(bi.ptr, bg.ptr, out.ptr, n, h, w, _stream(stream)))
    return out.mark_ready(stream)


def histogram_layer_bwd(img, ghist, max_bin, stream=None):
    """Gradient of :func:`histogram_layer` (un-pooled) w.r.t. ``img``."""
    bi, bg = _dev_inputs(stream, img, ghist)
    if len(bi.shape) != 4 or tuple(bg.shape) != tuple(bi.shape[:3]) + (bi.shape[3] * int(max_bin),):
        raise ValueError(f"histogram_layer_bwd: img {bi.shape} / ghist {bg.shape} mismatch")
    n, h, w, c = bi.shape
    out = DeviceArray.empty(bi.shape, bi.device)
    N.check(N.lib.shdr_soft_hist_bwd_f32(bi.ptr, bg.ptr, out.ptr, n, h, w, c, int(max_bin), _stream(stream)))
    return out.mark_ready(stream)


# --------------------------------------------------------------------------- reference-shaped classes
class AEInvcrfDecodeNet:
    """The two non-Keras methods of the reference class, same signatures."""

    def __init__(self, table_path="invemor.txt"):
        self.s = N.EMOR_SAMPLES
        self.n_p = N.EMOR_NCOMP + 1
        self.table_path = table_path

    def parse_invemor(self):
        return parse_invemor(self.table_path)

    def invcrf_pca_w_2_invcrf(self, invcrf_pca_w):
        self.parse_invemor()                 # cached; keeps the reference's "table from CWD" contract
        return invcrf_pca_w_2_invcrf(invcrf_pca_w)


class model:
    """The per-pixel methods of ``linearization_net.model`` (backbone excluded)."""

    def histogram_layer(self, img, max_bin):
        return histogram_layer(img, max_bin)

    def frontend(self, img):
        return frontend(img)

    @staticmethod
    def _increase(rf):
        return _increase(rf)
