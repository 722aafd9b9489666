"""``tensorflow.keras`` names the reference's modules need at IMPORT time only (class bases).  The stand-in never
builds a Keras network: the per-pixel methods are called unbound."""


class Model:
    def __init__(self, *args, **kwargs):
        raise RuntimeError("the NumPy stand-in for TensorFlow cannot build Keras models; call the per-pixel methods unbound")


class _Anything:
    def __getattr__(self, name):
        raise RuntimeError(f"tensorflow.keras.{name} is not available in the NumPy stand-in")


layers = _Anything()
regularizers = _Anything()
