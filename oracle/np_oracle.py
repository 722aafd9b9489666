"""NumPy restatement of the reference's Linearization-Net per-pixel path.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``) -- PARITY UNPINNED.

Every function mirrors the reference op-for-op: one NumPy full-tensor op per
TensorFlow op, in the same order, rounding to ``dtype`` (float32 by default) at
the same points.  Passing ``dtype=np.float64`` evaluates the same formulas in
double precision with double-precision constants (the "truth" the fp32 paths
are compared with).  [TF-sem] marks TensorFlow semantics that are not visible
in ``/root/reference`` (TensorFlow itself is not vendored there).

All ``file:line`` citations are into the reference repository
(ShinYwings/SingleHDR-tf2).
"""
from __future__ import annotations

import os

import numpy as np

BINS = (4, 8, 16)   # linearization_net.py:322
S = 1024            # linearization_net.py:181   samples per curve
NCOMP = 11          # linearization_net.py:182,225  PCA components used


# --------------------------------------------------------------------------
# (A) feature front end
# --------------------------------------------------------------------------
def _reflect_pad1(img):
    """tf.pad(img, [[0,0],[1,1],[1,1],[0,0]], 'REFLECT') [TF-sem]: index -1 -> 1,
    h -> h-2 (no edge repeat).  Requires h >= 2 and w >= 2."""
    if img.shape[1] < 2 or img.shape[2] < 2:
        raise ValueError("REFLECT pad needs h >= 2 and w >= 2")
    return np.pad(img, ((0, 0), (1, 1), (1, 1), (0, 0)), mode="reflect")


# tap order = row-major over the 3x3 filter, as a depthwise VALID correlation
# accumulates it [TF-sem]; products with +-1, +-2, 0 are exact in fp32.
_KY = ((-1.0, -2.0, -1.0), (0.0, 0.0, 0.0), (1.0, 2.0, 1.0))   # d/dy
_KX = ((-1.0, 0.0, 1.0), (-2.0, 0.0, 2.0), (-1.0, 0.0, 1.0))   # d/dx


def sobel_edges6(img, dtype=np.float32):
    """``tf.image.sobel_edges(img)`` reshaped to 6 channels.

    linearization_net.py:312-314.  Output channel ``c*2 + k`` with k=0 -> dy,
    k=1 -> dx (the ``[b,h,w,3,2]`` tensor flattened row-major).
    """
    img = np.asarray(img, dtype=dtype)
    n, h, w, c = img.shape
    p = _reflect_pad1(img)
    out = np.empty((n, h, w, c, 2), dtype=dtype)
    for k, ker in enumerate((_KY, _KX)):
        acc = np.zeros((n, h, w, c), dtype=dtype)
        for r in range(3):
            for s in range(3):
                wt = ker[r][s]
                if wt == 0.0:
                    continue  # + 0*x is exact; skipping keeps the order of the rest
                acc = acc + dtype(wt) * p[:, r:r + h, s:s + w, :]
        out[..., k] = acc
    return out.reshape(n, h, w, c * 2)


def hist_centers(max_bin, dtype=np.float32):
    """Bin centres ``tf.divide(2.*i - 1., 2.*max_bin)`` for i = 1..max_bin.

    linearization_net.py:342,345.  [TF-sem] both Python floats become float32
    tensors, so the quotient is an fp32 division.
    """
    i = np.arange(1, max_bin + 1)
    return (2.0 * i - 1.0).astype(dtype) / dtype(2.0 * max_bin)


def histogram_layer(img, max_bin, dtype=np.float32):
    """``model.histogram_layer(img, max_bin)``, linearization_net.py:336-350.

    Per bin i=1..B: d=|img - (2i-1)/(2B)|; h = (d < 1/B) ? 1 - d*B : 0; bins
    concatenated on the channel axis (channel = (i-1)*C + c).
    """
    img = np.asarray(img, dtype=dtype)
    threshold = dtype(1.0 / max_bin)            # :339  python double -> tensor dtype
    centers = hist_centers(max_bin, dtype)      # :342,345
    bins = []
    for i in range(max_bin):                    # :344
        distance = np.abs(img - centers[i])     # :345
        histo = np.where(distance < threshold,  # :346  tf.less is strict
                         dtype(1.0) - distance * dtype(max_bin),
                         dtype(0.0))
        bins.append(histo.astype(dtype))
    return np.concatenate(bins, axis=-1)        # :349


def avg_pool_same(x, k=16, dtype=np.float32):
    """``average_pooling2d(x, k, 1, 'same')`` -- the optional pool of
    linearization_net.py:351 (dead code there; README.md:51).

    [TF-sem] SAME pads (k-1)//2 before and the rest after; the average divides
    by the number of in-bounds elements; the CPU kernel accumulates a window
    in input raster order.  Zero padding is used here only as an exact no-op
    (x + 0.0 == x), so the accumulation order over valid elements is raster.
    """
    x = np.asarray(x, dtype=dtype)
    n, h, w, c = x.shape
    pb = (k - 1) // 2
    pa = k - 1 - pb
    p = np.pad(x, ((0, 0), (pb, pa), (pb, pa), (0, 0)))
    acc = np.zeros_like(x)
    for dy in range(k):
        for dx in range(k):
            acc += p[:, dy:dy + h, dx:dx + w, :]
    ys = np.arange(h)
    xs = np.arange(w)
    cy = np.minimum(ys + pa, h - 1) - np.maximum(ys - pb, 0) + 1
    cx = np.minimum(xs + pa, w - 1) - np.maximum(xs - pb, 0) + 1
    cnt = (cy[:, None] * cx[None, :]).astype(dtype)
    return (acc / cnt[None, :, :, None]).astype(dtype)


def hist_multi(img, bins=BINS, pool_k=0, dtype=np.float32):
    """concat of ``histogram_layer(img, B)`` for B in ``bins`` (each optionally
    pooled), linearization_net.py:322 minus the img/edge pieces."""
    parts = []
    for b in bins:
        hst = histogram_layer(img, b, dtype)
        if pool_k:
            hst = avg_pool_same(hst, pool_k, dtype)
        parts.append(hst)
    return np.concatenate(parts, axis=-1)


def frontend(img, bins=BINS, pool_k=0, dtype=np.float32):
    """The 93-channel tensor fed to ``crfFeatureNet``:
    ``concat([img, edge6, hist4, hist8, hist16], -1)``, linearization_net.py:312-322."""
    img = np.asarray(img, dtype=dtype)
    return np.concatenate(
        [img, sobel_edges6(img, dtype), hist_multi(img, bins, pool_k, dtype)], axis=-1)


def bf16_round(x):
    """float32 -> bfloat16 (round to nearest even) -> float32: what ``cvt.rn.bf16.f32`` does to finite values."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u >> 16) & 1) + 0x7FFF
    return ((u + r) & 0xFFFF0000).astype(np.uint32).view(np.float32)


def half_round(x):
    """float32 -> IEEE half (round to nearest even, subnormals kept) -> float32, saturating at +-65504 instead of
    overflowing to inf (what the fused conv1 kernel does to its operands); NaN stays NaN."""
    x = np.asarray(x, dtype=np.float32)
    return np.clip(x, np.float32(-65504.0), np.float32(65504.0)).astype(np.float16).astype(np.float32)


def conv2d_same_s2(x, kernel, bias=None, dtype=np.float64):
    """``tf.keras.layers.Conv2D(filters, (kh, kw), strides=(2, 2), padding='SAME')`` on NHWC ``x`` with an HWIO
    ``kernel`` -- ``crfFeatureNet.conv1``, linearization_net.py:91,107 (7x7, 93 -> 64, bias).  TF 'SAME':
    ``out = ceil(in / 2)``, ``pad = max((out - 1) * 2 + k - in, 0)`` with ``pad // 2`` in front and the rest behind,
    zeros; cross-correlation (no kernel flip).  One matmul per tap, accumulated in ``dtype``."""
    x = np.asarray(x, dtype=dtype)
    kernel = np.asarray(kernel, dtype=dtype)
    n, h, w, c = x.shape
    kh, kw, ci, co = kernel.shape
    assert ci == c
    oh, ow = (h + 1) // 2, (w + 1) // 2
    ph, pw = max((oh - 1) * 2 + kh - h, 0), max((ow - 1) * 2 + kw - w, 0)
    p = np.pad(x, ((0, 0), (ph // 2, ph - ph // 2), (pw // 2, pw - pw // 2), (0, 0)))
    out = np.zeros((n, oh, ow, co), dtype)
    for ky in range(kh):
        for kx in range(kw):
            out += p[:, ky:ky + 2 * oh - 1:2, kx:kx + 2 * ow - 1:2, :] @ kernel[ky, kx]
    if bias is not None:
        out += np.asarray(bias, dtype)
    return out


def frontend_conv1(img, kernel, bias=None, half_operands=False, dtype=np.float64):
    """``crfFeatureNet.conv1(concat([img, edge6, hist4, hist8, hist16]))``: linearization_net.py:312-322 -> :107.
    The features are the fp32 ones of :func:`frontend`; ``half_operands=True`` rounds features and kernel to fp16
    first (what the fused tensor-core kernel multiplies), the sum is carried in ``dtype``."""
    feat = frontend(img)
    kernel = np.asarray(kernel, np.float32)
    if half_operands:
        feat, kernel = half_round(feat), half_round(kernel)
    return conv2d_same_s2(feat, kernel, bias, dtype)


# --------------------------------------------------------------------------
# (B) inverse-CRF stage
# --------------------------------------------------------------------------
def _parse(lines, tag):
    """``AEInvcrfDecodeNet._parse``, linearization_net.py:255-268: 256 lines of
    four tokens after the line equal to ``tag``."""
    for line_idx, line in enumerate(lines):
        if line == tag:
            break
    else:
        raise ValueError(f"tag {tag!r} not found")
    r = []
    for idx in range(line_idx + 1, line_idx + 1 + 256):
        r += lines[idx].split()
    return np.float32(r)


def parse_invemor(path="invemor.txt"):
    """``AEInvcrfDecodeNet.parse_invemor``, linearization_net.py:217-227.
    Returns (B[1024], g0[1024], hinv[1024,11]) float32."""
    with open(os.path.join(path), "r") as f:
        lines = [line.strip() for line in f.readlines()]
    b = _parse(lines, "B =")
    g0 = _parse(lines, "g0 =")
    hinv = np.stack([_parse(lines, f"hinv({i + 1})=") for i in range(NCOMP)], axis=-1)
    return b, g0, hinv


parse_table = parse_invemor


def invcrf_pca_w_2_invcrf(w, g0, hinv, dtype=np.float32):
    """``AEInvcrfDecodeNet.invcrf_pca_w_2_invcrf``, linearization_net.py:231-253:
    ``g0 + matmul(tile(hinv)[b,s,11], w[b,11,1])`` squeezed to [b,s]."""
    w = np.asarray(w, dtype=dtype)
    b = w.shape[0]
    g0_ = np.asarray(g0, dtype=dtype).reshape(1, -1, 1)                 # :239
    h_ = np.tile(np.asarray(hinv, dtype=dtype)[None], (b, 1, 1))        # :241-243
    inv = g0_ + np.matmul(h_, w[:, :, None])                            # :246-249
    return inv[..., 0].astype(dtype)                                    # :251


def increase(rf, dtype=np.float32):
    """``model._increase``, linearization_net.py:368-392."""
    rf = np.asarray(rf, dtype=dtype)
    g = rf[:, 1:] - rf[:, :-1]                                  # :370
    min_g = np.min(g, axis=-1, keepdims=True)                   # :373
    r = np.maximum(-min_g, dtype(0.0))                          # :377 relu(-min)
    new_g = g + r                                               # :380
    with np.errstate(invalid="ignore", divide="ignore"):
        new_g = new_g / np.sum(new_g, axis=-1, keepdims=True, dtype=dtype)   # :383
    new_rf = np.cumsum(new_g, axis=-1, dtype=dtype)             # :386 inclusive, sequential
    return np.pad(new_rf, ((0, 0), (1, 0))).astype(dtype)       # :389


def apply_rf(x, rf, dtype=np.float32):
    """``tf_utils.apply_rf`` -> ``interp_1d`` -> ``sample_1d``,
    tf_utils.py:95-105, 70-93, 54-68."""
    x = np.asarray(x, dtype=dtype)
    rf = np.asarray(rf, dtype=dtype)
    b = x.shape[0]
    k = rf.shape[1]
    y = dtype(k - 1) * x.reshape(b, -1)                         # :103
    y_0 = np.floor(y)                                           # :77
    y_1 = y_0 + dtype(1.0)                                      # :78
    with np.errstate(invalid="ignore"):
        i0 = np.clip(y_0.astype(np.int32), 0, k - 1)            # :82,66
        i1 = np.clip(y_1.astype(np.int32), 0, k - 1)
    rows = np.arange(b)[:, None]                                # :61-63
    v0 = rf[rows, i0]                                           # :68 gather_nd
    v1 = rf[rows, i1]
    w_0 = y_1 - y                                               # :87
    w_1 = y - y_0                                               # :88
    out = w_0 * v0 + w_1 * v1                                   # :93  mul, mul, add
    return out.reshape(x.shape).astype(dtype)                   # :105


def linearize(x, w, g0, hinv, dtype=np.float32):
    """PCA reconstruct -> ``_increase`` -> ``apply_rf``: what the inference
    graph does between ``Dense(11)`` and ``B_pred``
    (linearization_net.py:325-328, test_real_refinement.py:94-95)."""
    curve = increase(invcrf_pca_w_2_invcrf(w, g0, hinv, dtype), dtype)
    return apply_rf(x, curve, dtype), curve


# --------------------------------------------------------------------------
# (C) the steps either side of apply_rf in the inference graph
# --------------------------------------------------------------------------
def clip01(x, dtype=np.float32):
    """``tf.clip_by_value(pred_deq, 0, 1)``, test_real_refinement.py:91."""
    return np.minimum(np.maximum(np.asarray(x, dtype=dtype), dtype(0.0)), dtype(1.0))


def alpha_mask(b_pred, thr, dtype=np.float32):
    """The hallucination blend mask of test_real_refinement.py:98-101 (train.py:207-211):
    ``alpha = reduce_max(B_pred, axis=3)``; ``min(1, max(0, alpha - 1 + thr) / thr)``; tiled over 3 channels."""
    b_pred = np.asarray(b_pred, dtype=dtype)
    a = np.max(b_pred, axis=3)
    a = np.minimum(dtype(1.0), np.maximum(dtype(0.0), (a - dtype(1.0)) + dtype(thr)) / dtype(thr))
    return np.tile(a[..., None], (1, 1, 1, 3)).astype(dtype)


def linearize_ex(x, rf, thr, clip=True, dtype=np.float32):
    """clip -> apply_rf -> alpha mask, as the inference graph chains them.  Returns (C_pred, B_pred, alpha)."""
    c = clip01(x, dtype) if clip else np.asarray(x, dtype=dtype)
    y = apply_rf(c, rf, dtype)
    return c, y, alpha_mask(y, thr, dtype)


def synth_ldr(hdr, t, sigma_s, sigma_c, noise_s, noise_c, crf, dtype=np.float32):
    """The per-pixel part of ``_preprocessing`` (train.py:28-51): returns (_hdr_t, clipped_hdr_t, ldr, quantized_hdr).
    ``noise_s`` / ``noise_c`` stand for the two ``tf.random.normal`` tensors, ``sigma_s`` / ``sigma_c`` ([b,3]) for the
    ``0.08/6 * uniform`` and ``0.005 * uniform`` factors of shape [b,1,1,3]."""
    hdr = np.asarray(hdr, dtype=dtype)
    b = hdr.shape[0]
    _hdr_t = hdr * np.asarray(t, dtype=dtype).reshape(b, 1, 1, 1)                      # :31
    ss = np.asarray(sigma_s, dtype=dtype).reshape(b, 1, 1, 3)
    sc = np.asarray(sigma_c, dtype=dtype).reshape(b, 1, 1, 3)
    noise_s_map = ss * _hdr_t                                                        # :37
    noise_s_t = np.asarray(noise_s, dtype=dtype) * noise_s_map                       # :38
    temp_x = _hdr_t + noise_s_t                                                      # :39
    noise_c_t = sc * np.asarray(noise_c, dtype=dtype)                                # :40
    temp_x = temp_x + noise_c_t                                                      # :41
    _hdr_t = np.maximum(temp_x, dtype(0.0))                                          # :42 relu
    clipped = clip01(_hdr_t, dtype)                                                  # :45
    ldr = apply_rf(clipped, crf, dtype)                                              # :48
    quant = np.round(ldr * dtype(255.0))                                             # :51 tf.round: half to even
    return _hdr_t, clipped, ldr, quant.astype(dtype)


# --------------------------------------------------------------------------
# (D) reverse-mode gradients: what TensorFlow's autodiff computes for the reference's op sequences
#     (train.py:186-194, joint_training.py:156-186, finetune_real_dataset.py:149-178).  [TF-sem]: floor / cast /
#     less have no gradient; abs -> sign (0 at 0); where routes to the taken branch; reduce_min splits evenly among
#     ties; gather_nd -> scatter-add; relu'(0) = 0.
# --------------------------------------------------------------------------
def apply_rf_grad(x, rf, gy, dtype=np.float64, index_dtype=np.float32):
    """Gradient of :func:`apply_rf` w.r.t. ``x`` and ``rf``.  The positions ``y = (k-1) x``, their floor and the two
    weights are evaluated in ``index_dtype`` (float32: the bins and weights the fp32 forward pass used -- a value of
    ``y`` within rounding of an integer must land in the same bin as in the forward); products and sums in ``dtype``."""
    rf = np.asarray(rf, dtype=dtype); gy = np.asarray(gy, dtype=dtype)
    b, k = rf.shape
    xf = np.asarray(x, dtype=index_dtype).reshape(b, -1)
    gf = gy.reshape(b, -1)
    y = index_dtype(k - 1) * xf
    y0 = np.floor(y); y1 = y0 + index_dtype(1.0)
    with np.errstate(invalid="ignore"):
        i0 = np.clip(y0.astype(np.int64), 0, k - 1); i1 = np.clip(y1.astype(np.int64), 0, k - 1)
    w0 = (y1 - y).astype(dtype); w1 = (y - y0).astype(dtype)
    rows = np.arange(b)[:, None]
    gx = dtype(k - 1) * (gf * rf[rows, i1] - gf * rf[rows, i0])
    grf = np.zeros_like(rf)
    np.add.at(grf, (np.broadcast_to(rows, i0.shape), i0), gf * w0)
    np.add.at(grf, (np.broadcast_to(rows, i1.shape), i1), gf * w1)
    return gx.reshape(np.shape(x)), grf


def increase_grad(rf, gout, dtype=np.float64):
    """Gradient of :func:`increase` w.r.t. ``rf``."""
    rf = np.asarray(rf, dtype=dtype); gout = np.asarray(gout, dtype=dtype)
    g = rf[:, 1:] - rf[:, :-1]
    m = np.min(g, axis=-1, keepdims=True)
    r = np.maximum(-m, 0)
    u = g + r
    s = np.sum(u, axis=-1, keepdims=True)
    n = u / s
    dn = np.cumsum(gout[:, :0:-1], axis=-1)[:, ::-1]            # reverse inclusive cumsum of gout[:, 1:]
    du = (dn - np.sum(dn * n, axis=-1, keepdims=True)) / s
    dr = np.sum(du, axis=-1, keepdims=True)
    tie = (g == m)
    dg = du + np.where(m < 0, -dr, 0) * tie / np.sum(tie, axis=-1, keepdims=True)
    grf = np.zeros_like(rf)
    grf[:, 1:] += dg
    grf[:, :-1] -= dg
    return grf


def invcrf_pca_grad(gcurve, hinv, dtype=np.float64):
    """Gradient of :func:`invcrf_pca_w_2_invcrf` w.r.t. ``w``: ``gcurve [b,1024] @ hinv [1024,11]``."""
    return np.asarray(gcurve, dtype=dtype) @ np.asarray(hinv, dtype=dtype)


def histogram_layer_grad(img, ghist, max_bin, dtype=np.float64):
    """Gradient of :func:`histogram_layer` w.r.t. ``img``; the compare uses the fp32 forward's values so that the
    taken branch is the one the fp32 forward took."""
    img32 = np.asarray(img, dtype=np.float32)
    c = img32.shape[-1]
    thr = np.float32(1.0 / max_bin)
    centers = hist_centers(max_bin, np.float32)
    gimg = np.zeros(img32.shape, dtype=dtype)
    ghist = np.asarray(ghist, dtype=dtype)
    for i in range(max_bin):
        d = img32 - centers[i]
        taken = np.abs(d) < thr
        gimg += np.where(taken, -dtype(max_bin) * np.sign(d).astype(dtype) * ghist[..., i * c:(i + 1) * c], 0)
    return gimg


def sobel_edges6_grad(gedge, shape, dtype=np.float64):
    """Gradient of :func:`sobel_edges6` w.r.t. ``img``: scatter every tap of the REFLECT-padded correlation back."""
    n, h, w, c = shape
    ge = np.asarray(gedge, dtype=dtype).reshape(n, h, w, c, 2)
    gp = np.zeros((n, h + 2, w + 2, c), dtype=dtype)
    for k, ker in enumerate((_KY, _KX)):
        for r in range(3):
            for s in range(3):
                if ker[r][s] != 0.0:
                    gp[:, r:r + h, s:s + w, :] += dtype(ker[r][s]) * ge[..., k]
    g = gp[:, 1:-1, 1:-1, :].copy()
    # fold the reflected border back: padded row 0 mirrors row 1, padded row h+1 mirrors row h-2 (same for columns)
    gfull = gp
    g[:, 1, :, :] += gfull[:, 0, 1:-1, :]
    g[:, h - 2, :, :] += gfull[:, h + 1, 1:-1, :]
    g[:, :, 1, :] += gfull[:, 1:-1, 0, :]
    g[:, :, w - 2, :] += gfull[:, 1:-1, w + 1, :]
    g[:, 1, 1, :] += gfull[:, 0, 0, :]
    g[:, 1, w - 2, :] += gfull[:, 0, w + 1, :]
    g[:, h - 2, 1, :] += gfull[:, h + 1, 0, :]
    g[:, h - 2, w - 2, :] += gfull[:, h + 1, w + 1, :]
    return g


def frontend_grad(img, gfeat, bins=BINS, dtype=np.float64):
    """Gradient of :func:`frontend` (un-pooled) w.r.t. ``img``."""
    img32 = np.asarray(img, dtype=np.float32)
    gfeat = np.asarray(gfeat, dtype=dtype)
    g = gfeat[..., :3].copy() + sobel_edges6_grad(gfeat[..., 3:9], img32.shape, dtype)
    off = 9
    for b in bins:
        g += histogram_layer_grad(img32, gfeat[..., off:off + 3 * b], b, dtype)
        off += 3 * b
    return g
