"""Framework-agnostic device plumbing over libshdr: DLPack in/out, device arrays,
pinned host buffers, streams and events.  Pure ctypes + numpy -- no TensorFlow and
no PyTorch import anywhere in this module.

DLPack contract (v0.x capsules, what ``tf.experimental.dlpack.to_dlpack`` and
``torch.utils.dlpack.to_dlpack`` / ``Tensor.__dlpack__`` produce):

* inputs: any object with ``__dlpack__`` (or a raw ``"dltensor"`` capsule) is
  *borrowed* -- the capsule is kept alive for the duration of the call, never
  renamed, never written;
* outputs: :class:`DeviceArray` owns a library-allocated ``DLManagedTensor``;
  ``__dlpack__()`` hands it over (one shot) as a ``"dltensor"`` capsule whose
  destructor frees the memory if no consumer takes it.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N

_pyapi = C.pythonapi
_pyapi.PyCapsule_GetPointer.restype = C.c_void_p
_pyapi.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
_pyapi.PyCapsule_IsValid.restype = C.c_int
_pyapi.PyCapsule_IsValid.argtypes = [C.py_object, C.c_char_p]
_pyapi.PyCapsule_New.restype = C.py_object
_pyapi.PyCapsule_New.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]

_DLTENSOR = b"dltensor"          # must outlive every capsule created with it
_CAPSULE_DTOR = N.lib.shdr_dl_capsule_destructor()


class _DLDevice(C.Structure):
    _fields_ = [("device_type", C.c_int32), ("device_id", C.c_int32)]


class _DLDataType(C.Structure):
    _fields_ = [("code", C.c_uint8), ("bits", C.c_uint8), ("lanes", C.c_uint16)]


class _DLTensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("device", _DLDevice), ("ndim", C.c_int32),
                ("dtype", _DLDataType), ("shape", C.POINTER(C.c_int64)),
                ("strides", C.POINTER(C.c_int64)), ("byte_offset", C.c_uint64)]


class _DLManagedTensor(C.Structure):
    _fields_ = [("dl_tensor", _DLTensor), ("manager_ctx", C.c_void_p), ("deleter", C.c_void_p)]


def _describe(managed_ptr):
    t = C.cast(managed_ptr, C.POINTER(_DLManagedTensor)).contents.dl_tensor
    shape = tuple(int(t.shape[i]) for i in range(t.ndim))
    return (t.data or 0) + t.byte_offset, shape, int(t.device.device_id), int(t.device.device_type)


def _itemsize(managed_ptr):
    t = C.cast(managed_ptr, C.POINTER(_DLManagedTensor)).contents.dl_tensor
    return int(t.dtype.bits) // 8


def _is_float(managed_ptr):
    """True for kDLFloat tensors (fp32, fp16); False for kDLBfloat."""
    t = C.cast(managed_ptr, C.POINTER(_DLManagedTensor)).contents.dl_tensor
    return int(t.dtype.code) == 2


class DeviceArray:
    """A CUDA tensor owned by libshdr (compact row-major): float32, or a 16-bit type for the reduced-precision front
    end (``numpy()`` then returns ``float16``, or the bit patterns as ``uint16`` for bfloat16)."""

    def __init__(self, managed_ptr):
        self._m = managed_ptr
        self.ptr, self.shape, self.device, _ = _describe(managed_ptr)
        self.itemsize = _itemsize(managed_ptr)
        self._float = _is_float(managed_ptr)

    # -- construction
    @classmethod
    def empty(cls, shape, device=0):
        shape = tuple(int(s) for s in shape)
        arr = (C.c_int64 * len(shape))(*shape)
        out = C.c_void_p()
        N.check(N.lib.shdr_dl_alloc_f32(arr, len(shape), device, C.byref(out)))
        return cls(out.value)

    @classmethod
    def from_numpy(cls, a, device=0, stream=None):
        a = np.ascontiguousarray(a, dtype=np.float32)
        d = cls.empty(a.shape, device)
        if a.size:
            N.check(N.lib.shdr_h2d(d.ptr, a.ctypes.data, a.nbytes, device, stream))
            N.check(N.lib.shdr_stream_sync(stream, device))
        d.mark_ready(stream)
        return d

    # -- properties
    @property
    def size(self):
        return int(np.prod(self.shape, dtype=np.int64))

    @property
    def nbytes(self):
        return self.size * self.itemsize

    def _alive(self):
        if self._m is None:
            raise RuntimeError("DeviceArray was exported with __dlpack__(); the consumer owns it now")

    def numpy(self, stream=None):
        self._alive()
        out = np.empty(self.shape, np.float32 if self.itemsize == 4 else (np.float16 if self._float else np.uint16))
        if out.size:
            # the copy stream waits for the producing kernel (whatever stream it ran on)
            N.check(N.lib.shdr_dl_wait_ready(self._m, getattr(stream, "handle", stream), 0))
            N.check(N.lib.shdr_d2h(out.ctypes.data, self.ptr, out.nbytes, self.device, stream))
            N.check(N.lib.shdr_stream_sync(stream, self.device))
        return out

    # -- DLPack producer protocol
    def __dlpack_device__(self):
        return (2, self.device)        # kDLCUDA

    def mark_ready(self, stream=None):
        """Record that the work just enqueued on ``stream`` produces this tensor (the ``shdr_dl_*`` ops do it
        themselves); ``__dlpack__`` / ``numpy`` then wait for exactly that work instead of the whole device."""
        self._alive()
        N.check(N.lib.shdr_dl_mark_ready(self._m, getattr(stream, "handle", stream)))
        return self

    def __dlpack__(self, stream=None, **_unused):
        """DLPack producer protocol.  ``stream``: the consumer's CUDA stream (an integer handle; 1 and 2 are the
        legacy / per-thread default streams) -> that stream is made to wait for the producing kernel; ``None`` (what
        ``from_dlpack(capsule)`` callers such as TF effectively give) -> the HOST waits for the producing kernel only
        (an event, not a device-wide synchronise); ``-1`` -> no synchronisation."""
        self._alive()
        if stream is None:
            N.check(N.lib.shdr_dl_wait_ready(self._m, None, 1))
        elif stream != -1:
            h = None if stream in (0, 1) else stream          # 1 = legacy default stream; 2 = per-thread default
            N.check(N.lib.shdr_dl_wait_ready(self._m, h, 0))
        cap = _pyapi.PyCapsule_New(self._m, _DLTENSOR, _CAPSULE_DTOR)
        self._m = None
        return cap

    def __del__(self):
        m, self._m = getattr(self, "_m", None), None
        if m is not None:
            N.lib.shdr_dl_release(m)

    def __repr__(self):
        return f"DeviceArray(shape={self.shape}, device={self.device}, ptr=0x{self.ptr:x})"


class Borrowed:
    """Keeps an input's DLPack capsule alive and exposes its DLManagedTensor*."""

    def __init__(self, obj):
        self._keep = obj
        if isinstance(obj, Borrowed):                 # already borrowed (tf_adapter holds inputs across a call)
            self.managed = obj.managed
            self.ptr, self.shape, self.device, self.device_type = obj.ptr, obj.shape, obj.device, obj.device_type
            return
        if isinstance(obj, DeviceArray):
            obj._alive()
            self.managed = obj._m
        else:
            cap = obj
            if type(obj).__name__ != "PyCapsule":
                cap = to_capsule(obj)
            if not _pyapi.PyCapsule_IsValid(cap, _DLTENSOR):
                raise TypeError("expected a DLPack 'dltensor' capsule or an object with __dlpack__")
            self._cap = cap
            self.managed = _pyapi.PyCapsule_GetPointer(cap, _DLTENSOR)
        self.ptr, self.shape, self.device, self.device_type = _describe(self.managed)


def to_capsule(obj):
    """DLPack capsule of a framework tensor (TF eager tensor, torch tensor, cupy ...)."""
    if hasattr(obj, "__dlpack__"):
        return obj.__dlpack__()
    mod = type(obj).__module__ or ""
    if mod.startswith("tensorflow"):
        import tensorflow as tf   # only reached for TensorFlow tensors
        return tf.experimental.dlpack.to_dlpack(obj)
    raise TypeError(f"cannot get a DLPack capsule from {type(obj)!r}")


# ---------------------------------------------------------------------------
class PinnedArray:
    """float32 numpy view over page-locked host memory (cudaHostAlloc)."""

    def __init__(self, shape):
        self.shape = tuple(int(s) for s in shape)
        n = int(np.prod(self.shape, dtype=np.int64))
        p = C.c_void_p()
        N.check(N.lib.shdr_malloc_host(C.byref(p), max(n, 1) * 4))
        self._p = p.value
        buf = (C.c_float * max(n, 1)).from_address(self._p)
        self.array = np.frombuffer(buf, dtype=np.float32, count=n).reshape(self.shape)

    @property
    def ptr(self):
        return self._p

    def free(self):
        if self._p:
            self.array = None
            N.lib.shdr_free_host(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Stream:
    def __init__(self, device=0):
        self.device = device
        p = C.c_void_p()
        N.check(N.lib.shdr_stream_create(C.byref(p), device))
        self.handle = p.value

    def sync(self):
        N.check(N.lib.shdr_stream_sync(self.handle, self.device))

    def wait(self, event):
        N.check(N.lib.shdr_stream_wait_event(self.handle, event.handle, self.device))

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            N.lib.shdr_stream_destroy(h, self.device)


class Event:
    def __init__(self, device=0):
        self.device = device
        p = C.c_void_p()
        N.check(N.lib.shdr_event_create(C.byref(p), device))
        self.handle = p.value

    def record(self, stream=None):
        h = stream.handle if isinstance(stream, Stream) else stream
        N.check(N.lib.shdr_event_record(self.handle, h, self.device))

    def elapsed_ms(self, stop):
        ms = C.c_float()
        N.check(N.lib.shdr_event_elapsed_ms(self.handle, stop.handle, C.byref(ms)))
        return ms.value

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            N.lib.shdr_event_destroy(h, self.device)


def synchronize(device=0):
    N.check(N.lib.shdr_sync(device))
