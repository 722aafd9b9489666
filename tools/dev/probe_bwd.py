"""dev probe: time the backward kernels (CUDA events), random and smooth inputs"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import shdr
from shdr import _native as N
shdr.require_gpu()
z = np.load(os.path.join(os.path.dirname(__file__), "..", "..", "tests", "golden", "invemor_f32.npz"))
shdr.set_emor_table(z["g0"], z["hinv"])
def timeit(fn, reps=10):
    for _ in range(3): fn()
    shdr.synchronize()
    e0, e1 = shdr.Event(), shdr.Event()
    e0.record(None)
    for _ in range(reps): fn()
    e1.record(None)
    return e0.elapsed_ms(e1) / reps
D = shdr.DeviceArray.from_numpy
rng = np.random.default_rng(0)
b, h, w = 16, 1024, 1024
for name, x in (("random", rng.random((b, h, w, 3), dtype=np.float32)),
                ("smooth", np.tile(np.linspace(0, 1, h * w * 3, dtype=np.float32).reshape(1, h, w, 3), (b, 1, 1, 1)))):
    dx = D(x); rf = D(np.tile(np.linspace(0, 1, 1024, dtype=np.float32) ** 2, (b, 1))); gy = D(rng.normal(size=x.shape).astype(np.float32))
    gx = shdr.DeviceArray.empty(x.shape); grf = shdr.DeviceArray.empty((b, 1024))
    per = h * w * 3
    t_both = timeit(lambda: N.check(N.lib.shdr_apply_rf_bwd_f32(dx.ptr, rf.ptr, gy.ptr, gx.ptr, grf.ptr, b, per, 1024, None)))
    t_rf = timeit(lambda: N.check(N.lib.shdr_apply_rf_bwd_f32(dx.ptr, rf.ptr, gy.ptr, None, grf.ptr, b, per, 1024, None)))
    t_x = timeit(lambda: N.check(N.lib.shdr_apply_rf_bwd_f32(dx.ptr, rf.ptr, gy.ptr, gx.ptr, None, b, per, 1024, None)))
    print(f"apply_rf_bwd {name}: both {t_both:.3f} ms  grf only {t_rf:.3f}  gx only {t_x:.3f}   (forward apply ~0.07 ms; 12 B/elem -> {b*per*12/6458.4e6:.3f} ms at roofline)")
img = D(rng.random((8, 512, 512, 3), dtype=np.float32)); gf = D(rng.normal(size=(8, 512, 512, 93)).astype(np.float32)); gi = shdr.DeviceArray.empty((8, 512, 512, 3))
t = timeit(lambda: N.check(N.lib.shdr_frontend_bwd_f32(img.ptr, gf.ptr, gi.ptr, 8, 512, 512, None)))
print(f"frontend_bwd 8x512x512: {t:.3f} ms ({8*512*512*(372+24)/6458.4e6:.3f} ms at roofline)")
w11 = D(rng.normal(0, .5, (16, 11)).astype(np.float32)); gc = D(rng.normal(size=(16, 1024)).astype(np.float32)); gw = shdr.DeviceArray.empty((16, 11))
t = timeit(lambda: N.check(N.lib.shdr_invcrf_build_bwd_f32(w11.ptr, gc.ptr, gw.ptr, 16, 1, None)))
print(f"invcrf_build_bwd (monotone) b=16: {t*1e3:.1f} us")
