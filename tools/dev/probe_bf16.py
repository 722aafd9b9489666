"""dev probe: fp32 vs bf16 output of the un-pooled 93-channel front end"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ctypes as C
import numpy as np
import shdr
from shdr import _native as N
shdr.require_gpu()
def timeit(fn, reps=20):
    for _ in range(3): fn()
    shdr.synchronize()
    e0, e1 = shdr.Event(), shdr.Event()
    e0.record(None)
    for _ in range(reps): fn()
    e1.record(None)
    return e0.elapsed_ms(e1) / reps
for (n, h, w) in [(8, 512, 512), (2, 2160, 3840)]:
    img = shdr.DeviceArray.from_numpy(np.random.default_rng(0).random((n, h, w, 3), dtype=np.float32))
    o32 = shdr.DeviceArray.empty((n, h, w, 93))
    p16 = C.c_void_p(); N.check(N.lib.shdr_malloc(C.byref(p16), n * h * w * 93 * 2, 0))
    px = n * h * w
    t32 = timeit(lambda: N.check(N.lib.shdr_frontend_f32(img.ptr, o32.ptr, n, h, w, 0, None)))
    t16 = timeit(lambda: N.check(N.lib.shdr_frontend_bf16(img.ptr, p16.value, n, h, w, None)))
    print(f"{n}x{h}x{w}: fp32 {t32:.4f} ms ({px*384/t32/1e6:.0f} GB/s)  bf16 {t16:.4f} ms ({px*198/t16/1e6:.0f} GB/s of its own 198 B/px)  speed-up {t32/t16:.2f}x")
    N.lib.shdr_free(p16.value, 0)
