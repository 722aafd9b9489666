"""Runs the UNMODIFIED reference source (``/root/reference`` or ``$SHDR_REFERENCE_DIR``) for the per-pixel path --
TEST INFRASTRUCTURE: the parity anchor of the oracle and, when real TensorFlow is importable, the CPU baseline of
``bench.py`` (``kind: "reference"``).

Two backends:

* ``"tf"``      real TensorFlow (GPUs hidden, eager).  Not installable in this project's image; auto-detected.
* ``"standin"`` ``oracle/standin/tensorflow`` -- a NumPy implementation of the published semantics of the few TF ops
                these functions call.  Pins the reference's Python (op order, constants, indexing); cannot pin the
                arithmetic inside TF's primitives.

The reference functions are called exactly as the reference calls them: ``model.histogram_layer`` (an instance method
that does not use ``self``) and ``model._increase`` / ``tf_utils.apply_rf`` unbound or static;
``AEInvcrfDecodeNet.invcrf_pca_w_2_invcrf`` on a minimal object that borrows the class's own ``parse_invemor`` /
``_parse`` (the real constructor would build Keras layers); the front-end glue of ``model.call`` (:312-322) is the
only part restated here (three lines: ``sobel_edges``, ``reshape``, ``concat``), because ``call`` cannot be stopped
half way.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
STANDIN_DIR = os.path.join(HERE, "standin")
_state = {}


def reference_dir():
    return os.environ.get("SHDR_REFERENCE_DIR", "/root/reference")


def have_reference_source():
    d = reference_dir()
    return all(os.path.exists(os.path.join(d, f)) for f in ("linearization_net.py", "tf_utils.py", "invemor.txt"))


def have_real_tf():
    if "tf_real" not in _state:
        ok = False
        if not any(os.path.abspath(p) == STANDIN_DIR for p in sys.path):
            try:
                spec = importlib.util.find_spec("tensorflow")
                ok = spec is not None
            except Exception:
                ok = False
        _state["tf_real"] = ok
    return _state["tf_real"]


def available(backend=None):
    """Can the reference's own functions run here with this backend (None: real TF)?"""
    if not have_reference_source():
        return False
    return have_real_tf() if backend in (None, "tf") else backend == "standin"


def load(backend=None):
    """Import the reference's ``linearization_net`` and ``tf_utils`` on top of the chosen TensorFlow."""
    backend = backend or ("tf" if have_real_tf() else "standin")
    if _state.get("loaded") == backend:
        return _state["mods"]
    if "loaded" in _state:
        raise RuntimeError(f"reference already loaded with backend {_state['loaded']!r} in this process")
    if not have_reference_source():
        raise RuntimeError(f"reference source not found under {reference_dir()}")
    if backend == "standin":
        for m in [m for m in sys.modules if m == "tensorflow" or m.startswith("tensorflow.")]:
            del sys.modules[m]
        sys.path.insert(0, STANDIN_DIR)
    import tensorflow as tf
    if backend == "tf":
        try:
            tf.config.set_visible_devices([], "GPU")      # the baseline is the reference's CPU path
        except Exception:
            pass
    sys.path.insert(0, reference_dir())
    try:
        lin = importlib.import_module("linearization_net")
        tfu = importlib.import_module("tf_utils")
    finally:
        sys.path.remove(reference_dir())
    _state.update(loaded=backend, mods=(tf, lin, tfu))
    return _state["mods"]


def _np(t):
    return np.asarray(t.numpy() if hasattr(t, "numpy") else t)


def histogram_layer(img, max_bin, backend=None):
    tf, lin, _ = load(backend)
    return _np(lin.model.histogram_layer(None, tf.constant(np.asarray(img, np.float32)), max_bin))


def frontend(img, backend=None):
    """linearization_net.py:312-322: the tensor handed to crf_feature_net."""
    tf, lin, _ = load(backend)
    t = tf.constant(np.asarray(img, np.float32))
    edge_1 = tf.image.sobel_edges(t)                                                        # :312
    edge_1 = tf.reshape(edge_1, [tf.shape(t)[0], tf.shape(t)[1], tf.shape(t)[2], 6])          # :314
    h = lin.model.histogram_layer
    return _np(tf.concat([t, edge_1, h(None, t, 4), h(None, t, 8), h(None, t, 16)], -1))      # :322


def hist_multi(img, backend=None):
    tf, lin, _ = load(backend)
    t = tf.constant(np.asarray(img, np.float32))
    h = lin.model.histogram_layer
    return _np(tf.concat([h(None, t, 4), h(None, t, 8), h(None, t, 16)], -1))


def increase(rf, backend=None):
    tf, lin, _ = load(backend)
    return _np(lin.model._increase(tf.constant(np.asarray(rf, np.float32))))


def apply_rf(x, rf, backend=None):
    tf, _, tfu = load(backend)
    return _np(tfu.apply_rf(tf.constant(np.asarray(x, np.float32)), tf.constant(np.asarray(rf, np.float32))))


def _decoder(lin):
    cls = lin.AEInvcrfDecodeNet
    obj = types.SimpleNamespace()
    obj._parse = cls._parse
    obj.parse_invemor = types.MethodType(cls.parse_invemor, obj)
    obj.invcrf_pca_w_2_invcrf = types.MethodType(cls.invcrf_pca_w_2_invcrf, obj)
    return obj


def parse_invemor(backend=None):
    _, lin, _ = load(backend)
    cwd = os.getcwd()
    os.chdir(reference_dir())                       # the reference opens 'invemor.txt' relative to the CWD (:219)
    try:
        return _decoder(lin).parse_invemor()
    finally:
        os.chdir(cwd)


def invcrf_pca_w_2_invcrf(w, backend=None):
    tf, lin, _ = load(backend)
    cwd = os.getcwd()
    os.chdir(reference_dir())
    try:
        return _np(_decoder(lin).invcrf_pca_w_2_invcrf(tf.constant(np.asarray(w, np.float32))))
    finally:
        os.chdir(cwd)


def linearize(x, w, backend=None):
    """the inference graph between Dense(11) and B_pred (linearization_net.py:325-328, test_real_refinement.py:95)"""
    curve = increase(invcrf_pca_w_2_invcrf(w, backend), backend)
    return apply_rf(x, curve, backend), curve
