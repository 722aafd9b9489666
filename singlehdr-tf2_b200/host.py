"""Host-buffer API: numpy in, numpy out, with the host<->device copies pipelined against the
kernels.  This is the end-to-end path a caller without device tensors uses (and the one
``bench.py`` reports as ``e2e``): a batch is cut into chunks of whole images; each chunk goes
H2D -> kernel -> D2H on its own stream, and ``slots`` chunks are in flight so that the
PCIe transfers of neighbouring chunks overlap each other and the compute.

Still no CPU arithmetic: the host only moves bytes.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N
from .device import Stream


class _Slot:
    def __init__(self, device, in_bytes, out_bytes):
        self.device = device
        self.stream = Stream(device)
        self.d_in, self.d_out = C.c_void_p(), C.c_void_p()
        N.check(N.lib.shdr_malloc(C.byref(self.d_in), in_bytes, device))
        N.check(N.lib.shdr_malloc(C.byref(self.d_out), out_bytes, device))

    def free(self):
        for p in (self.d_in, self.d_out):
            if p.value:
                N.lib.shdr_free(p.value, self.device)
                p.value = None


class HostPipeline:
    """Runs ``op(d_in, d_out, n_items, stream)`` over a host batch in pipelined chunks.

    in_item_bytes / out_item_bytes: bytes of one batch item on input / output.
    """

    def __init__(self, op, in_item_bytes, out_item_bytes, items_per_chunk, device=0, slots=3):
        N.require_gpu()
        self.op = op
        self.ib, self.ob = int(in_item_bytes), int(out_item_bytes)
        self.chunk = max(1, int(items_per_chunk))
        self.device = device
        self.slots = [_Slot(device, self.ib * self.chunk, self.ob * self.chunk) for _ in range(slots)]

    def run(self, src: np.ndarray, dst: np.ndarray, n_items: int):
        """src/dst: C-contiguous float32 host arrays (pinned memory makes the copies asynchronous)."""
        assert src.flags.c_contiguous and dst.flags.c_contiguous
        assert src.nbytes == n_items * self.ib and dst.nbytes == n_items * self.ob, "host buffer size mismatch"
        sp, dp = src.ctypes.data, dst.ctypes.data
        k = 0
        for i0 in range(0, n_items, self.chunk):
            m = min(self.chunk, n_items - i0)
            s = self.slots[k % len(self.slots)]
            k += 1
            st = s.stream.handle   # stream order protects the slot's buffers from the previous chunk
            N.check(N.lib.shdr_h2d(s.d_in.value, sp + i0 * self.ib, m * self.ib, self.device, st))
            self.op(s.d_in.value, s.d_out.value, m, st, i0)
            N.check(N.lib.shdr_d2h(dp + i0 * self.ob, s.d_out.value, m * self.ob, self.device, st))
        for s in self.slots:
            s.stream.sync()

    def close(self):
        for s in self.slots:
            s.free()
        self.slots = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _as_f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _chunk_for(item_out_bytes, n, target_bytes=256 << 20):
    return max(1, min(n, target_bytes // max(1, item_out_bytes)))


def frontend_host(img, out=None, pool=False, device=0):
    """numpy ``[n,h,w,3]`` -> numpy ``[n,h,w,93]``."""
    img = _as_f32(img)
    n, h, w, c = img.shape
    if c != 3:
        raise ValueError("frontend: img must have 3 channels")
    if out is None:
        out = np.empty((n, h, w, N.FRONTEND_CH), np.float32)
    pk = 16 if pool else 0

    def op(d_in, d_out, m, st, _i0):
        N.check(N.lib.shdr_frontend_f32(d_in, d_out, m, h, w, pk, st))

    ib, ob = h * w * 3 * 4, h * w * N.FRONTEND_CH * 4
    p = HostPipeline(op, ib, ob, _chunk_for(ob, n), device)
    try:
        p.run(img, out, n)
    finally:
        p.close()
    return out


def frontend_conv1_host(img, kernel, bias=None, scale=None, relu=False, out=None, device=0):
    """numpy ``[n,h,w,3]`` + conv1's kernel ``[7,7,93,64]`` (and bias / folded-norm scale ``[64]``) -> numpy
    ``[n, ceil(h/2), ceil(w/2), 64]``: the front end fused into ``crfFeatureNet.conv1`` (tensor cores, fp16 operands).
    Per input pixel 12 B go to the device and 64 B come back -- against 372 B for the 93-channel tensor."""
    img = _as_f32(img)
    n, h, w, c = img.shape
    if c != 3:
        raise ValueError("frontend_conv1: img must have 3 channels")
    kernel = _as_f32(kernel)
    if kernel.shape != (7, 7, N.FRONTEND_CH, 64):
        raise ValueError(f"frontend_conv1: kernel must be [7,7,93,64], got {kernel.shape}")
    oh, ow = (h + 1) // 2, (w + 1) // 2
    if out is None:
        out = np.empty((n, oh, ow, 64), np.float32)
    N.require_gpu()
    small = {"kernel": kernel}
    for name, v in (("bias", bias), ("scale", scale)):
        if v is not None:
            v = _as_f32(v)
            if v.shape != (64,):
                raise ValueError(f"frontend_conv1: {name} must be [64], got {v.shape}")
            small[name] = v
    dev = {k: C.c_void_p() for k in list(small) + ["packed"]}
    try:
        for k, v in small.items():
            N.check(N.lib.shdr_malloc(C.byref(dev[k]), v.nbytes, device))
            N.check(N.lib.shdr_h2d(dev[k].value, v.ctypes.data, v.nbytes, device, None))
        N.check(N.lib.shdr_malloc(C.byref(dev["packed"]), int(N.lib.shdr_conv1_packed_bytes()), device))
        N.check(N.lib.shdr_conv1_pack_weights_f32(dev["kernel"].value, dev["packed"].value, None))
        N.check(N.lib.shdr_sync(device))
        d_scale = dev["scale"].value if "scale" in dev else None
        d_bias = dev["bias"].value if "bias" in dev else None

        def op(d_in, d_out, m, st, _i0):
            N.check(N.lib.shdr_frontend_conv1_f32(d_in, dev["packed"].value, d_scale, d_bias, 1 if relu else 0, d_out,
                                                  m, h, w, st))

        ib, ob = h * w * 3 * 4, oh * ow * 64 * 4
        p = HostPipeline(op, ib, ob, _chunk_for(ob, n, 64 << 20), device)
        try:
            p.run(img, out, n)
        finally:
            p.close()
    finally:
        for v in dev.values():
            if v.value:
                N.lib.shdr_free(v.value, device)
    return out


def hist_multi_host(img, out=None, pool=False, device=0):
    """numpy ``[n,h,w,3]`` -> numpy ``[n,h,w,84]`` (hist4 | hist8 | hist16, optionally pooled)."""
    img = _as_f32(img)
    n, h, w, c = img.shape
    if c != 3:
        raise ValueError("hist_multi: img must have 3 channels")
    if out is None:
        out = np.empty((n, h, w, N.HIST_CH), np.float32)
    pk = 16 if pool else 0

    def op(d_in, d_out, m, st, _i0):
        N.check(N.lib.shdr_hist_multi_f32(d_in, d_out, m, h, w, pk, st))

    ib, ob = h * w * 3 * 4, h * w * N.HIST_CH * 4
    p = HostPipeline(op, ib, ob, _chunk_for(ob, n), device)
    try:
        p.run(img, out, n)
    finally:
        p.close()
    return out


def linearize_host(x, w, out=None, device=0):
    """numpy ``x [b, ...]``, ``w [b,11]`` -> ``(y, curve[b,1024])`` (PCA + _increase + apply_rf)."""
    x, w = _as_f32(x), _as_f32(w)
    b = x.shape[0]
    per = int(np.prod(x.shape[1:], dtype=np.int64))
    if w.shape != (b, N.EMOR_NCOMP):
        raise ValueError(f"linearize: w must be [{b},11], got {w.shape}")
    if out is None:
        out = np.empty_like(x)
    N.require_gpu()
    d_w, d_curve = C.c_void_p(), C.c_void_p()
    N.check(N.lib.shdr_malloc(C.byref(d_w), w.nbytes, device))
    N.check(N.lib.shdr_malloc(C.byref(d_curve), b * N.EMOR_SAMPLES * 4, device))
    curve = np.empty((b, N.EMOR_SAMPLES), np.float32)
    try:
        N.check(N.lib.shdr_h2d(d_w.value, w.ctypes.data, w.nbytes, device, None))
        N.check(N.lib.shdr_invcrf_build_f32(d_w.value, d_curve.value, b, 1, None))
        N.check(N.lib.shdr_d2h(curve.ctypes.data, d_curve.value, curve.nbytes, device, None))
        N.check(N.lib.shdr_sync(device))

        def op(d_in, d_out, m, st, i0):
            rf = d_curve.value + i0 * N.EMOR_SAMPLES * 4
            N.check(N.lib.shdr_apply_rf_f32(d_in, rf, d_out, m, per, N.EMOR_SAMPLES, st))

        p = HostPipeline(op, per * 4, per * 4, _chunk_for(per * 4, b, 64 << 20), device)
        try:
            p.run(x, out, b)
        finally:
            p.close()
    finally:
        N.lib.shdr_free(d_w.value, device)
        N.lib.shdr_free(d_curve.value, device)
    return out, curve


def apply_rf_host(x, rf, out=None, device=0):
    """numpy ``x [b, ...]``, ``rf [b,k]`` -> numpy, shape of ``x``."""
    x, rf = _as_f32(x), _as_f32(rf)
    b = x.shape[0]
    k = rf.shape[1]
    per = int(np.prod(x.shape[1:], dtype=np.int64))
    if rf.shape[0] != b:
        raise ValueError("apply_rf: batch of x and rf differ")
    if out is None:
        out = np.empty_like(x)
    N.require_gpu()
    d_rf = C.c_void_p()
    N.check(N.lib.shdr_malloc(C.byref(d_rf), max(rf.nbytes, 4), device))
    try:
        N.check(N.lib.shdr_h2d(d_rf.value, rf.ctypes.data, rf.nbytes, device, None))
        N.check(N.lib.shdr_sync(device))

        def op(d_in, d_out, m, st, i0):
            N.check(N.lib.shdr_apply_rf_f32(d_in, d_rf.value + i0 * k * 4, d_out, m, per, k, st))

        p = HostPipeline(op, per * 4, per * 4, _chunk_for(per * 4, b, 64 << 20), device)
        try:
            p.run(x, out, b)
        finally:
            p.close()
    finally:
        N.lib.shdr_free(d_rf.value, device)
    return out
