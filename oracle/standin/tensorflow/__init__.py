"""A NumPy stand-in for the sliver of TensorFlow 2 that the reference's per-pixel functions touch
(``linearization_net.model.histogram_layer`` / ``_increase`` / ``AEInvcrfDecodeNet.invcrf_pca_w_2_invcrf`` and
``tf_utils.apply_rf``) -- TEST INFRASTRUCTURE.

TensorFlow cannot be installed in this project's image.  Putting this directory on ``sys.path`` lets the UNMODIFIED
reference source files be imported and executed eagerly, so that the oracle can be pinned against what the
reference's own Python does (op order, constants, index arithmetic, channel order).  What it cannot pin is the
arithmetic inside each TensorFlow primitive: every function here implements the PUBLISHED semantics of the TF op of
the same name in float32 NumPy ([TF-sem]: python scalars become tensors of the other operand's dtype; ``tf.cumsum`` is
an inclusive sequential sum; ``tf.cast(float -> int32)`` truncates; ``tf.gather_nd`` after ``clip_by_value`` never
goes out of range; ``tf.image.sobel_edges`` = REFLECT pad + depthwise correlation with [[-1,-2,-1],[0,0,0],[1,2,1]]
and its transpose, output last dims [c, 2] = [dy, dx]).  ``tools/make_tf_golden.py --backend tf`` produces the same
vectors from real TensorFlow the day one is available.
"""
import numpy as _np

float32 = _np.float32
int32 = _np.int32
__version__ = "0.0-numpy-standin"
IS_STANDIN = True


class _Shape(tuple):
    def as_list(self):
        return list(self)


class Tensor:
    __array_priority__ = 1000

    def __init__(self, a):
        self._a = _np.asarray(a)

    # -- introspection
    @property
    def shape(self):
        return _Shape(self._a.shape)

    @property
    def dtype(self):
        return self._a.dtype

    def get_shape(self):
        return _Shape(self._a.shape)

    def numpy(self):
        return self._a

    def __iter__(self):                      # `b, _, = w.get_shape()`-style unpacking is on the shape, not here
        return iter(Tensor(x) for x in self._a)

    def __getitem__(self, idx):
        return Tensor(self._a[idx])

    def __int__(self):
        return int(self._a)

    def __index__(self):
        return int(self._a)

    # -- arithmetic: a python scalar takes the tensor's dtype, as tf.convert_to_tensor does in a binary op
    def _other(self, o):
        if isinstance(o, Tensor):
            return o._a
        return _np.asarray(o, dtype=self._a.dtype)

    def __add__(self, o): return Tensor(self._a + self._other(o))
    def __radd__(self, o): return Tensor(self._other(o) + self._a)
    def __sub__(self, o): return Tensor(self._a - self._other(o))
    def __rsub__(self, o): return Tensor(self._other(o) - self._a)
    def __mul__(self, o): return Tensor(self._a * self._other(o))
    def __rmul__(self, o): return Tensor(self._other(o) * self._a)
    def __truediv__(self, o): return Tensor(self._a / self._other(o))
    def __rtruediv__(self, o): return Tensor(self._other(o) / self._a)
    def __floordiv__(self, o): return Tensor(self._a // self._other(o))
    def __mod__(self, o): return Tensor(self._a % self._other(o))
    def __neg__(self): return Tensor(-self._a)


def _a(x, like=None):
    if isinstance(x, Tensor):
        return x._a
    if like is not None:
        return _np.asarray(x, dtype=like.dtype)
    if isinstance(x, float):
        return _np.asarray(x, dtype=_np.float32)        # python floats default to float32 tensors
    if isinstance(x, int):
        return _np.asarray(x, dtype=_np.int32)
    return _np.asarray(x)


def _pair(x, y):
    """operands of a binary op: a python scalar takes the dtype of the tensor operand; two python floats -> float32"""
    if isinstance(x, Tensor) and not isinstance(y, Tensor):
        return x._a, _np.asarray(y, dtype=x._a.dtype)
    if isinstance(y, Tensor) and not isinstance(x, Tensor):
        return _np.asarray(x, dtype=y._a.dtype), y._a
    return _a(x), _a(y)


def constant(v, dtype=None):
    return Tensor(_np.asarray(v, dtype=dtype) if dtype is not None else _a(v))


convert_to_tensor = constant


def abs(x): return Tensor(_np.abs(_a(x)))                                    # noqa: A001
def floor(x): return Tensor(_np.floor(_a(x)))
def divide(x, y): p, q = _pair(x, y); return Tensor(p / q)
def multiply(x, y): p, q = _pair(x, y); return Tensor(p * q)
def subtract(x, y): p, q = _pair(x, y); return Tensor(p - q)
def add(x, y): p, q = _pair(x, y); return Tensor(p + q)
def less(x, y): p, q = _pair(x, y); return Tensor(p < q)
def minimum(x, y): p, q = _pair(x, y); return Tensor(_np.minimum(p, q))
def maximum(x, y): p, q = _pair(x, y); return Tensor(_np.maximum(p, q))


def where(c, x, y):
    c = _a(c)
    xa = _a(x) if isinstance(x, Tensor) else None
    ya = _a(y) if isinstance(y, Tensor) else None
    ref = xa if xa is not None else ya
    xa = xa if xa is not None else _np.asarray(x, dtype=ref.dtype)
    ya = ya if ya is not None else _np.asarray(y, dtype=ref.dtype)
    return Tensor(_np.where(c, xa, ya))


def cast(x, dtype):
    a = _a(x)
    if _np.issubdtype(dtype, _np.integer) and _np.issubdtype(a.dtype, _np.floating):
        with _np.errstate(invalid="ignore"):
            return Tensor(_np.trunc(a).astype(dtype))                        # truncation toward zero
    return Tensor(a.astype(dtype))


def clip_by_value(x, lo, hi):
    a = _a(x)
    return Tensor(_np.minimum(_np.maximum(a, _np.asarray(lo, a.dtype)), _np.asarray(hi, a.dtype)))


def concat(values, axis): return Tensor(_np.concatenate([_a(v) for v in values], axis=axis))
def stack(values, axis=0): return Tensor(_np.stack([_a(v) for v in values], axis=axis))
def reshape(x, shape): return Tensor(_np.reshape(_a(x), [int(_a(s)) for s in shape]))
def expand_dims(x, axis): return Tensor(_np.expand_dims(_a(x), axis))
def squeeze(x, axis=None): return Tensor(_np.squeeze(_a(x), axis=axis))
def tile(x, multiples): return Tensor(_np.tile(_a(x), [int(_a(m)) for m in multiples]))
def shape(x): return Tensor(_np.asarray(_a(x).shape, dtype=_np.int32))
def range(n, dtype=_np.int32): return Tensor(_np.arange(int(_a(n)), dtype=dtype))        # noqa: A001


def reduce_min(x, axis=None, keepdims=False): return Tensor(_np.min(_a(x), axis=axis, keepdims=keepdims))
def reduce_max(x, axis=None, keepdims=False):
    return Tensor(_np.max(_a(x), axis=tuple(axis) if isinstance(axis, list) else axis, keepdims=keepdims))


def reduce_sum(x, axis=None, keepdims=False):
    a = _a(x)
    return Tensor(_np.sum(a, axis=axis, keepdims=keepdims, dtype=a.dtype))


def cumsum(x, axis=0):
    a = _a(x)
    return Tensor(_np.cumsum(a, axis=axis, dtype=a.dtype))                   # inclusive, sequential in the dtype


def pad(x, paddings, mode="CONSTANT"):
    a = _a(x)
    pw = [(int(_a(p[0])), int(_a(p[1]))) for p in paddings]
    m = {"CONSTANT": "constant", "REFLECT": "reflect", "SYMMETRIC": "symmetric"}[mode.upper()]
    return Tensor(_np.pad(a, pw, mode=m))


def matmul(x, y): return Tensor(_np.matmul(_a(x), _a(y)))


def gather_nd(params, indices):
    p, i = _a(params), _a(indices)
    return Tensor(p[tuple(i[..., d] for d in _np.arange(i.shape[-1]))])


class _NN:
    @staticmethod
    def relu(x):
        a = _a(x)
        return Tensor(_np.maximum(a, _np.asarray(0, a.dtype)))

    tanh = staticmethod(lambda x: Tensor(_np.tanh(_a(x))))


nn = _NN()


class _Image:
    @staticmethod
    def sobel_edges(image):
        """[TF-sem] tf.image.sobel_edges: REFLECT pad 1, depthwise correlation with the two Sobel kernels,
        output [b, h, w, c, 2] with the last axis [dy, dx]; taps accumulated in row-major order."""
        a = _a(image)
        n, h, w, c = a.shape
        p = _np.pad(a, ((0, 0), (1, 1), (1, 1), (0, 0)), mode="reflect")
        ky = ((-1.0, -2.0, -1.0), (0.0, 0.0, 0.0), (1.0, 2.0, 1.0))
        kx = ((-1.0, 0.0, 1.0), (-2.0, 0.0, 2.0), (-1.0, 0.0, 1.0))
        out = _np.empty((n, h, w, c, 2), dtype=a.dtype)
        for k, ker in enumerate((ky, kx)):
            acc = _np.zeros((n, h, w, c), dtype=a.dtype)
            for r in (0, 1, 2):
                for s in (0, 1, 2):
                    if ker[r][s] != 0.0:
                        acc = acc + a.dtype.type(ker[r][s]) * p[:, r:r + h, s:s + w, :]
            out[..., k] = acc
        return Tensor(out)


image = _Image()


class _Summary:
    @staticmethod
    def image(*args, **kwargs):
        return None


summary = _Summary()


class _Config:
    @staticmethod
    def set_visible_devices(devices, kind=None):
        return None

    @staticmethod
    def list_physical_devices(kind=None):
        return []


config = _Config()


def executing_eagerly():
    return True


def function(f=None, **kwargs):
    return f if f is not None else (lambda g: g)


from . import keras  # noqa: E402,F401
