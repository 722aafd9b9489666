"""ctypes binding of libshdr.so (declared in include/shdr.h).

There is deliberately no fallback of any kind here: if the shared library is
missing the import fails, and if no CUDA device is usable the compute entry
points raise :class:`ShdrError` carrying the library's message.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SHDR_LIB", os.path.join(_HERE, "libshdr.so"))

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_NOTABLE = 0, -1, -2, -3, -4
EMOR_SAMPLES, EMOR_NCOMP, FRONTEND_CH, HIST_CH = 1024, 11, 93, 84


class ShdrError(RuntimeError):
    """A libshdr entry point returned a negative status."""

    def __init__(self, code, message):
        super().__init__(f"libshdr error {code}: {message}")
        self.code = code


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        f"or singlehdr-tf2_b200/csrc/build.sh -- there is no CPU or PyTorch fallback for this path")

lib = C.CDLL(LIB_PATH)

_f = C.POINTER(C.c_float)
_vp = C.c_void_p
_i, _ll, _sz = C.c_int, C.c_longlong, C.c_size_t
_dl = C.c_void_p            # DLManagedTensor*
_dlp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); every symbol include/shdr.h declares
SIGNATURES = {
    "shdr_version": (_i, []),
    "shdr_last_error": (C.c_char_p, []),
    "shdr_device_count": (_i, [C.POINTER(_i)]),
    "shdr_set_emor_table": (_i, [_vp, _vp, _i, _i]),
    "shdr_sobel6_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "shdr_soft_hist_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "shdr_hist_multi_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "shdr_frontend_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "shdr_frontend_bf16": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "shdr_frontend_f16": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "shdr_conv1_packed_bytes": (_sz, []),
    "shdr_conv1_pack_weights_f32": (_i, [_vp, _vp, _vp]),
    "shdr_frontend_conv1_f32": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _vp]),
    "shdr_invcrf_build_f32": (_i, [_vp, _vp, _i, _i, _vp]),
    "shdr_increase_f32": (_i, [_vp, _vp, _i, _i, _vp]),
    "shdr_apply_rf_f32": (_i, [_vp, _vp, _vp, _i, _ll, _i, _vp]),
    "shdr_linearize_f32": (_i, [_vp, _vp, _vp, _vp, _i, _ll, _vp]),
    "shdr_apply_rf_ex_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _ll, _i, _i, C.c_float, _vp]),
    "shdr_linearize_ex_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _ll, _i, C.c_float, _vp]),
    "shdr_synth_ldr_f32": (_i, [_vp] * 11 + [_i, _ll, _i, _vp]),
    "shdr_apply_rf_bwd_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _ll, _i, _vp]),
    "shdr_increase_bwd_f32": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "shdr_invcrf_build_bwd_f32": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "shdr_frontend_bwd_f32": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "shdr_soft_hist_bwd_f32": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "shdr_dl_frontend": (_i, [_dl, _i, _vp, _dlp]),
    "shdr_dl_sobel6": (_i, [_dl, _vp, _dlp]),
    "shdr_dl_frontend_bf16": (_i, [_dl, _vp, _dlp]),
    "shdr_dl_frontend_f16": (_i, [_dl, _vp, _dlp]),
    "shdr_dl_soft_hist": (_i, [_dl, _i, _i, _vp, _dlp]),
    "shdr_dl_invcrf_build": (_i, [_dl, _i, _vp, _dlp]),
    "shdr_dl_increase": (_i, [_dl, _vp, _dlp]),
    "shdr_dl_apply_rf": (_i, [_dl, _dl, _vp, _dlp]),
    "shdr_dl_alloc_f32": (_i, [C.POINTER(C.c_int64), _i, _i, _dlp]),
    "shdr_dl_release": (None, [_dl]),
    "shdr_dl_mark_ready": (_i, [_dl, _vp]),
    "shdr_dl_wait_ready": (_i, [_dl, _vp, _i]),
    "shdr_dl_capsule_destructor": (_vp, []),
    "shdr_malloc": (_i, [_dlp, _sz, _i]),
    "shdr_free": (_i, [_vp, _i]),
    "shdr_malloc_host": (_i, [_dlp, _sz]),
    "shdr_free_host": (_i, [_vp]),
    "shdr_memset": (_i, [_vp, _i, _sz, _i, _vp]),
    "shdr_h2d": (_i, [_vp, _vp, _sz, _i, _vp]),
    "shdr_d2h": (_i, [_vp, _vp, _sz, _i, _vp]),
    "shdr_sync": (_i, [_i]),
    "shdr_stream_create": (_i, [_dlp, _i]),
    "shdr_stream_destroy": (_i, [_vp, _i]),
    "shdr_stream_sync": (_i, [_vp, _i]),
    "shdr_event_create": (_i, [_dlp, _i]),
    "shdr_event_destroy": (_i, [_vp, _i]),
    "shdr_event_record": (_i, [_vp, _vp, _i]),
    "shdr_event_elapsed_ms": (_i, [_vp, _vp, C.POINTER(C.c_float)]),
    "shdr_stream_wait_event": (_i, [_vp, _vp, _i]),
    "shdr_launch_count": (_ll, []),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError here = header and library disagree
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    return lib.shdr_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != OK:
        raise ShdrError(rc, last_error())


def device_count() -> int:
    n = C.c_int(0)
    check(lib.shdr_device_count(C.byref(n)))
    return n.value


def require_gpu() -> int:
    n = device_count()
    if n < 1:
        raise ShdrError(ERR_CUDA, "no CUDA device visible: libshdr has no CPU path")
    return n


def launch_count() -> int:
    return int(lib.shdr_launch_count())
