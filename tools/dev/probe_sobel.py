"""dev probe: stand-alone Sobel (strip kernel vs the generic kernel via a strided call)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
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
for (n, h, w) in [(8, 512, 512), (16, 1024, 1024)]:
    img = shdr.DeviceArray.from_numpy(np.random.default_rng(0).random((n, h, w, 3), dtype=np.float32))
    o6 = shdr.DeviceArray.empty((n, h, w, 6)); o7 = shdr.DeviceArray.empty((n, h, w, 7))
    px = n * h * w
    ts = timeit(lambda: N.check(N.lib.shdr_sobel6_f32(img.ptr, o6.ptr, n, h, w, 3, 6, 0, None)))
    tg = timeit(lambda: N.check(N.lib.shdr_sobel6_f32(img.ptr, o7.ptr, n, h, w, 3, 7, 0, None)))   # strided -> generic kernel
    print(f"{n}x{h}x{w}: strip {ts:.4f} ms ({px*36/ts/1e6:.0f} GB/s of 36 B/px)  generic {tg:.4f} ms ({px*36/tg/1e6:.0f} GB/s)")
