"""ctypes wrapper of oracle/shdr_oracle.c -- TEST INFRASTRUCTURE ONLY, PARITY UNPINNED (see oracle/__init__.py).

A second, independent restatement of the reference path (plain C, OpenMP over rows).  `available()` is False until
`make -C oracle` (or `__graft_entry__.build()`) has produced oracle/_build/libshdr_oracle.so.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libshdr_oracle.so")
_lib = None


def available() -> bool:
    return os.path.exists(_PATH)


def _l():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_PATH)
        _lib.shdr_oracle_threads.restype = C.c_int
    return _lib


def threads() -> int:
    return int(_l().shdr_oracle_threads())


def set_threads(n: int) -> None:
    """Override OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1 to every rank)."""
    _l().shdr_oracle_set_threads(int(n))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def sobel_edges6(img):
    img = _f32(img)
    n, h, w, c = img.shape
    out = np.empty((n, h, w, 2 * c), np.float32)
    _l().shdr_oracle_sobel6(_p(img), _p(out), n, h, w, c)
    return out


def histogram_layer(img, max_bin):
    img = _f32(img)
    c = img.shape[-1]
    out = np.empty(img.shape[:-1] + (c * max_bin,), np.float32)
    _l().shdr_oracle_hist(_p(img), _p(out), C.c_longlong(img.size // c), c, int(max_bin))
    return out


def avg_pool_same(x, k=16):
    x = _f32(x)
    n, h, w, c = x.shape
    out = np.empty_like(x)
    _l().shdr_oracle_avg_pool_same(_p(x), _p(out), n, h, w, c, int(k))
    return out


def frontend(img, bins=(4, 8, 16), pool_k=0, with_img_edge=True):
    img = _f32(img)
    n, h, w, c = img.shape
    assert c == 3
    ch = (9 if with_img_edge else 0) + 3 * sum(bins)
    out = np.empty((n, h, w, ch), np.float32)
    b = (C.c_int * len(bins))(*bins)
    _l().shdr_oracle_frontend(_p(img), _p(out), n, h, w, b, len(bins), int(pool_k), int(bool(with_img_edge)))
    return out


def hist_multi(img, bins=(4, 8, 16), pool_k=0):
    return frontend(img, bins, pool_k, with_img_edge=False)


def invcrf_pca_w_2_invcrf(w, g0, hinv):
    w, g0, hinv = _f32(w), _f32(g0), _f32(hinv)
    out = np.empty((w.shape[0], g0.shape[0]), np.float32)
    _l().shdr_oracle_pca(_p(w), _p(g0), _p(hinv), _p(out), w.shape[0], g0.shape[0], hinv.shape[1])
    return out


def increase(rf):
    rf = _f32(rf)
    out = np.empty_like(rf)
    _l().shdr_oracle_increase(_p(rf), _p(out), rf.shape[0], rf.shape[1])
    return out


def apply_rf(x, rf):
    x, rf = _f32(x), _f32(rf)
    out = np.empty_like(x)
    b = x.shape[0]
    _l().shdr_oracle_apply_rf(_p(x), _p(rf), _p(out), b, C.c_longlong(x.size // max(b, 1)), rf.shape[1])
    return out


def linearize(x, w, g0, hinv):
    curve = increase(invcrf_pca_w_2_invcrf(w, g0, hinv))
    return apply_rf(x, curve), curve
