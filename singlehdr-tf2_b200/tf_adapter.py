"""TensorFlow-2 adapter: drops the native path in under the reference's model code.

TensorFlow is NOT installed in the build image or on the GPU boxes of this project, so this
module is import-safe without it and everything that touches ``tf`` is resolved lazily.  It is
~60 lines of glue on purpose: tensors cross via ``tf.experimental.dlpack`` (zero copy), results
come back through ``from_dlpack``; inside a ``@tf.function`` the call is wrapped in
``tf.py_function`` (the reference's step functions are graph-traced:
train.py:182, joint_training.py:137, finetune_real_dataset.py:144, test_real_refinement.py:86).

Usage:

    import linearization_net, tf_utils           # the reference's modules
    import shdr.tf_adapter as shdr_tf
    shdr_tf.patch(linearization_net, tf_utils)   # in place; SHDR_NATIVE=0 leaves stock TF ops
"""
from __future__ import annotations

import os

from . import layers


def _tf():
    import tensorflow as tf
    return tf


def _wrap(native_fn, out_shape_fn):
    """Lift ``native_fn(*DeviceArray-compatible) -> DeviceArray`` to TF tensors."""
    def eager(*tensors):
        tf = _tf()
        outs = native_fn(*tensors)     # layers.* borrow the tensors through DLPack
        return tf.experimental.dlpack.from_dlpack(outs.__dlpack__())

    def call(*tensors):
        tf = _tf()
        if tf.executing_eagerly():
            return eager(*tensors)
        y = tf.py_function(eager, list(tensors), tf.float32)
        y.set_shape(out_shape_fn(*[t.shape for t in tensors]))
        return y
    return call


def patch(linearization_net=None, tf_utils=None, table_path="invemor.txt"):
    """Monkey-patch the reference modules in place (no new variables, checkpoint layout unchanged)."""
    if os.environ.get("SHDR_NATIVE", "1") == "0":
        return False
    if linearization_net is not None:
        layers.parse_invemor(table_path)
        m = linearization_net.model

        def histogram_layer(self, img, max_bin):
            f = _wrap(lambda t: layers.histogram_layer(t, max_bin),
                      lambda s: s[:-1].concatenate([s[-1] * max_bin]))
            return f(img)

        m.histogram_layer = histogram_layer
        m._increase = staticmethod(_wrap(layers._increase, lambda s: s))
        linearization_net.AEInvcrfDecodeNet.invcrf_pca_w_2_invcrf = (
            lambda self, w: _wrap(layers.invcrf_pca_w_2_invcrf, lambda s: s[:-1].concatenate([1024]))(w))
        # the fused 93-channel front end, for callers that replace model.call's concat (:312-322)
        linearization_net.shdr_frontend = _wrap(layers.frontend, lambda s: s[:-1].concatenate([93]))
    if tf_utils is not None:
        tf_utils.apply_rf = _wrap(layers.apply_rf, lambda sx, sr: sx)
    return True
