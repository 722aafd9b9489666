"""dev probe: time pooled hist_multi / frontend for a few shapes (CUDA events via libshdr), ms and roofline frac"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import shdr
from shdr import _native as N
shdr.require_gpu()
PEAK = 6458.4
def timeit(fn, reps=20):
    for _ in range(3): fn()
    shdr.synchronize()
    e0, e1 = shdr.Event(), shdr.Event()
    e0.record(None)
    for _ in range(reps): fn()
    e1.record(None)
    return e0.elapsed_ms(e1) / reps
for (n, h, w) in [(8, 512, 512), (32, 512, 512), (1, 2160, 3840), (4, 1080, 1920), (2, 256, 256)]:
    img = shdr.DeviceArray.from_numpy(np.random.default_rng(0).random((n, h, w, 3), dtype=np.float32))
    o84 = shdr.DeviceArray.empty((n, h, w, 84)); o93 = shdr.DeviceArray.empty((n, h, w, 93))
    px = n * h * w
    t84 = timeit(lambda: N.check(N.lib.shdr_hist_multi_f32(img.ptr, o84.ptr, n, h, w, 16, None)))
    t93 = timeit(lambda: N.check(N.lib.shdr_frontend_f32(img.ptr, o93.ptr, n, h, w, 16, None)))
    print(f"{n}x{h}x{w}: pooled84 {t84:.4f} ms frac {px*348/t84/1e6/PEAK:.3f} | pooled93 {t93:.4f} ms frac {px*384/t93/1e6/PEAK:.3f}")
