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

from . import layers
from .device import Borrowed, synchronize


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


def _shape_last(s, n):
    """Shape ``s`` with its last dimension replaced by ``n`` (tf.TensorShape or a plain tuple)."""
    try:
        return s[:-1].concatenate([n])
    except AttributeError:
        return tuple(s[:-1]) + (n,)


def patch(linearization_net=None, tf_utils=None, table_path="invemor.txt", replace_call=True):
    """Monkey-patch the reference modules in place.  Returns False (and changes nothing) when ``SHDR_NATIVE=0``.

    ``replace_call=False`` keeps the reference's ``model.call`` body (stock ``sobel_edges`` / ``concat``) and only
    swaps the per-layer methods."""
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

        if replace_call:
            def call(self, img, training="training"):
                # linearization_net.py:310-334 with :312-322 (sobel_edges, 3 x histogram_layer, concat) fused
                tf = _tf()
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
