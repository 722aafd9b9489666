"""dev probe: pooled 93-channel front end, rows that are / are not 16-byte aligned chunks (w % 4)"""
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
for (n, h, w) in [(32, 512, 512), (32, 512, 510)]:
    img = shdr.DeviceArray.from_numpy(np.random.default_rng(0).random((n, h, w, 3), dtype=np.float32))
    out = shdr.DeviceArray.empty((n, h, w, 93))
    px = n * h * w
    t = timeit(lambda: N.check(N.lib.shdr_frontend_f32(img.ptr, out.ptr, n, h, w, 16, None)))
    print(f"{n}x{h}x{w}: pooled 93-channel front end {t:.4f} ms = {px*384/t/1e6:.0f} GB/s = {px*384/t/1e6/6458.4:.3f} of the copy peak")
