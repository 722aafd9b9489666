"""TensorFlow-2 adapter: drops the native path in under the reference's model code.

TensorFlow is NOT installed in the build image or on the GPU boxes of this project, so this module is import-safe
without it, everything that touches ``tf`` is resolved lazily, and it is exercised against a stand-in ``tensorflow``
module (``tests/test_gpu_tf_adapter.py``); real-TF execution remains unverified (``tools/make_tf_golden.py`` is the
harness that pins it the day TF is available).

What ``patch(linearization_net, tf_utils)`` does, in place (no new variables, checkpoint layout unchanged):

* ``linearization_net.model.call`` (:310-334) is REPLACED by a body that calls the fused 93-channel front end (one
  kernel instead of ``sobel_edges`` + 3 x ``histogram_layer`` + ``concat``), then the stock Keras sub-networks
  ``self.crf_feature_net`` / ``self.ae_invcrf_decode_net``, then the native ``_increase``; the two
  ``tf.summary.image`` lines keep working on slices of the fused tensor;
* ``model.histogram_layer``, ``model._increase``, ``AEInvcrfDecodeNet.invcrf_pca_w_2_invcrf`` and
  ``tf_utils.apply_rf`` are replaced one for one.

* opt-in ``fuse_conv1=True`` (inference only): ``crfFeatureNet.call`` (:101-116) additionally gets the front end
  fused INTO its first layer -- ``conv1`` (7x7/2 'SAME', 93 -> 64, bias), the inference-mode ``norm1`` folded into a
  per-channel scale / shift, and ``act1`` -- as one tensor-core kernel that takes the 3-channel image; the 93-channel
  tensor never exists.  fp16 operands / fp32 accumulation (reduced precision, like TF's own mixed_float16 conv);
  whenever ``training`` is truthy the stock path (fp32 front end + Keras ``conv1``, differentiable) runs instead.

Every replacement carries its gradient through ``tf.custom_gradient`` (kernels in csrc/backward.cu), so the patched
modules work inside the reference's training steps (train.py:182-194, joint_training.py:137-186,
finetune_real_dataset.py:144-178) as well as in inference (test_real_refinement.py:86-106).  Tensors cross via
``tf.experimental.dlpack`` (zero copy); inside a ``@tf.function`` each call (forward and gradient) is a
``tf.py_function``.

Stream ordering (TensorFlow runs its kernels on private non-blocking streams that DLPack does not expose): before a
native launch the device is synchronised, so that TF has finished producing the inputs; the launch goes to the legacy
default stream; the host then waits for that kernel's completion event before the result is handed to TF and before
the input capsules are released.  Correct for any producer, at the price of running the op synchronously.

Usage:

    import linearization_net, tf_utils           # the reference's modules
    import shdr.tf_adapter as shdr_tf
    shdr_tf.patch(linearization_net, tf_utils)   # SHDR_NATIVE=0 leaves the stock TF ops in place
"""
from __future__ import annotations

import os

import numpy as np

from . import _native as N
from . import layers
from .device import Borrowed, DeviceArray, synchronize


def _tf():
    import tensorflow as tf
    return tf


def _run_native(native_fn, tensors):
    """Eager body: TF tensors in -> native kernel -> TF tensor(s) out (see the module docstring for the ordering)."""
    tf = _tf()
    held = [Borrowed(t) for t in tensors]            # keeps the DLPack capsules (and TF's buffers) alive until return
    for dev in {b.device for b in held}:
        synchronize(dev)                             # TF's streams have produced the inputs
    outs = native_fn(*held)
    single = not isinstance(outs, (tuple, list))
    res = []
    for o in ([outs] if single else outs):
        # __dlpack__() (stream=None) waits on the HOST for the producing kernel's event
        res.append(None if o is None else tf.experimental.dlpack.from_dlpack(o.__dlpack__()))
    return res[0] if single else tuple(res)


def _lift(native_fn, n_out, shape_fns):
    """``native_fn(*tensors) -> DeviceArray | tuple`` as a TF op usable in eager and graph mode."""
    def call(*tensors):
        tf = _tf()
        if tf.executing_eagerly():
            return _run_native(native_fn, tensors)
        y = tf.py_function(lambda *ts: _run_native(native_fn, ts), list(tensors),
                           tf.float32 if n_out == 1 else [tf.float32] * n_out)
        ys = [y] if n_out == 1 else list(y)
        for t, fn in zip(ys, shape_fns):
            t.set_shape(fn(*[x.shape for x in tensors]))
        return ys[0] if n_out == 1 else tuple(ys)
    return call


def _with_grad(forward, backward):
    """forward(*xs) -> y and backward(*xs, dy) -> tuple of gradients (one per x), glued with tf.custom_gradient."""
    def call(*tensors):
        tf = _tf()

        @tf.custom_gradient
        def op(*xs):
            y = forward(*xs)

            def grad(dy):
                g = backward(*xs, dy)
                return g if len(xs) > 1 else g[0]
            return y, grad
        return op(*tensors)
    return call


# ------------------------------------------------------------------ the differentiable ops
def _same(s):
    return s


def _make_ops(tf_shape_last):
    """Build the lifted forward/backward pairs once; tf_shape_last(s, n) = s with its last dim replaced by n."""
    ops = {}
    ops["apply_rf"] = _with_grad(
        _lift(layers.apply_rf, 1, [lambda sx, sr: sx]),
        _lift(layers.apply_rf_bwd, 2, [lambda sx, sr, sg: sx, lambda sx, sr, sg: sr]))
    ops["increase"] = _with_grad(
        _lift(layers._increase, 1, [_same]),
        lambda rf, dy: (_lift(layers._increase_bwd, 1, [lambda sr, sg: sr])(rf, dy),))
    ops["pca"] = _with_grad(
        _lift(layers.invcrf_pca_w_2_invcrf, 1, [lambda s: tf_shape_last(s, 1024)]),
        lambda w, dy: (_lift(lambda a, b: layers.invcrf_build_bwd(a, b, monotone=False), 1,
                             [lambda sw, sg: sw])(w, dy),))
    ops["frontend"] = _with_grad(
        _lift(layers.frontend, 1, [lambda s: tf_shape_last(s, 93)]),
        lambda img, dy: (_lift(layers.frontend_bwd, 1, [lambda si, sg: si])(img, dy),))

    def hist(max_bin):
        return _with_grad(
            _lift(lambda t: layers.histogram_layer(t, max_bin), 1, [lambda s: tf_shape_last(s, s[-1] * max_bin)]),
            lambda img, dy: (_lift(lambda a, b: layers.histogram_layer_bwd(a, b, max_bin), 1,
                                   [lambda si, sg: si])(img, dy),))
    ops["hist"] = hist
    return ops


def _small_to_host(b):
    """A small borrowed device tensor (conv bias, batch-norm vectors) as a numpy array."""
    out = np.empty(tuple(b.shape), np.float32)
    if out.size:
        N.check(N.lib.shdr_d2h(out.ctypes.data, b.ptr, out.nbytes, b.device, None))
        N.check(N.lib.shdr_stream_sync(None, b.device))
    return out


def _fused_conv1_native(eps):
    """(img, kernel, bias, gamma, beta, moving_mean, moving_variance) -> relu(norm1(conv1(front end(img)))), inference."""
    def run(img, kernel, bias, gamma, beta, mean, var):
        packed = layers.conv1_pack_weights(kernel)                     # the variable may have changed: re-pack (5 us)
        cb, g, b, m, v = (_small_to_host(t) for t in (bias, gamma, beta, mean, var))
        scale = (g / np.sqrt(v + np.float32(eps))).astype(np.float32)  # inference-mode BatchNormalization, folded
        shift = ((cb - m) * scale + b).astype(np.float32)
        dev = img.device
        return layers.frontend_conv1(img, packed, bias=DeviceArray.from_numpy(shift, dev),
                                     scale=DeviceArray.from_numpy(scale, dev), relu=True)
    return run


def _conv1_out_shape(si, *_):
    try:
        n, h, w = si[0], si[1], si[2]
        half = lambda d: None if d is None else (int(d) + 1) // 2      # noqa: E731
        return type(si)([n, half(h), half(w), 64]) if not isinstance(si, tuple) else (n, half(h), half(w), 64)
    except Exception:
        return None


def _shape_last(s, n):
    """Shape ``s`` with its last dimension replaced by ``n`` (tf.TensorShape or a plain tuple)."""
    try:
        return s[:-1].concatenate([n])
    except AttributeError:
        return tuple(s[:-1]) + (n,)


def patch(linearization_net=None, tf_utils=None, table_path="invemor.txt", replace_call=True, fuse_conv1=False):
    """Monkey-patch the reference modules in place.  Returns False (and changes nothing) when ``SHDR_NATIVE=0``.

    ``replace_call=False`` keeps the reference's ``model.call`` body (stock ``sobel_edges`` / ``concat``) and only
    swaps the per-layer methods.  ``fuse_conv1=True`` (needs ``replace_call``) also runs the front end inside
    ``crfFeatureNet.conv1`` on the tensor cores for inference calls (``training`` falsy); see the module docstring."""
    if os.environ.get("SHDR_NATIVE", "1") == "0":
        return False
    ops = _make_ops(_shape_last)
    if linearization_net is not None:
        layers.parse_invemor(table_path)
        m = linearization_net.model
        hist_cache = {}

        def histogram_layer(self, img, max_bin):
            if max_bin not in hist_cache:
                hist_cache[max_bin] = ops["hist"](max_bin)
            return hist_cache[max_bin](img)

        m.histogram_layer = histogram_layer
        m._increase = staticmethod(ops["increase"])
        linearization_net.AEInvcrfDecodeNet.invcrf_pca_w_2_invcrf = lambda self, w: ops["pca"](w)
        linearization_net.shdr_frontend = ops["frontend"]

        if replace_call and fuse_conv1:
            cfn = linearization_net.crfFeatureNet
            stock_cfn_call = cfn.call
            fused_cache = {}

            def cfn_call(self, ldr, training="training"):
                # linearization_net.py:101-116; a 3-channel input is the IMAGE (inference): conv1 + norm1 + act1 fused
                # with the front end.  Anything else (93-channel features, training) takes the stock layers.
                if training or ldr.shape[-1] != 3:
                    return stock_cfn_call(self, ldr, training)
                tf = _tf()
                eps = float(self.norm1.epsilon)
                if eps not in fused_cache:
                    fused_cache[eps] = _lift(_fused_conv1_native(eps), 1, [_conv1_out_shape])
                bias = self.conv1.bias if getattr(self.conv1, "use_bias", True) else tf.zeros([64], tf.float32)
                act1 = fused_cache[eps](ldr, *[tf.convert_to_tensor(v) for v in (
                    self.conv1.kernel, bias, self.norm1.gamma, self.norm1.beta, self.norm1.moving_mean,
                    self.norm1.moving_variance)])
                x = self.pool1(act1)
                for blk in (self.res1, self.res2, self.res3, self.res4, self.res5):
                    x = blk(x, training)
                return tf.reduce_mean(x, [1, 2], keepdims=False)

            cfn.call = cfn_call
            sobel_op = _lift(layers.sobel_edges6, 1, [lambda s: _shape_last(s, 6)])

        if replace_call:
            def call(self, img, training="training"):
                # linearization_net.py:310-334 with :312-322 (sobel_edges, 3 x histogram_layer, concat) fused
                tf = _tf()
                if fuse_conv1 and not training:
                    # the front end runs inside crfFeatureNet.conv1; the edge tensor is only built when the two
                    # tf.summary.image lines (:316-317) would actually record something
                    if getattr(tf.summary, "should_record_summaries", lambda: False)():
                        edge = sobel_op(img)
                        tf.summary.image('edge0', edge[:, :, :, 0:3])
                        tf.summary.image('edge1', edge[:, :, :, 3:6])
                    feature = tf.cast(self.crf_feature_net(img, training), tf.float32)
                    invcrf = self._increase(self.ae_invcrf_decode_net(feature))
                    return tf.cast(invcrf, tf.float32)
                feat93 = ops["frontend"](img)
                tf.summary.image('edge0', feat93[:, :, :, 3:6])      # == edge_1[:, :, :, 0:3]   (:316)
                tf.summary.image('edge1', feat93[:, :, :, 6:9])      # == edge_1[:, :, :, 3:6]   (:317)
                feature = self.crf_feature_net(feat93, training)
                feature = tf.cast(feature, tf.float32)
                invcrf = self.ae_invcrf_decode_net(feature)
                invcrf = self._increase(invcrf)
                return tf.cast(invcrf, tf.float32)

            m.call = call
    if tf_utils is not None:
        tf_utils.apply_rf = ops["apply_rf"]
    return True
