"""dev stress: the sliding-window kernel many times over shapes that exercise task boundaries, edge strips and both
variants; every result against the first run (bit-identical: the integer pipeline is deterministic) and one oracle check"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import shdr
shdr.require_gpu()
rng = np.random.default_rng(0)
shapes = [(32, 512, 512), (3, 67, 129), (1, 2160, 3840), (5, 300, 200), (2, 33, 1028), (7, 130, 64)]
bad = 0
for (n, h, w) in shapes:
    img = rng.random((n, h, w, 3), dtype=np.float32)
    d = shdr.DeviceArray.from_numpy(img)
    ref84 = shdr.hist_multi(d, pool=True).numpy()
    ref93 = shdr.frontend(d, pool=True).numpy() if w % 4 == 0 else None
    reps = 40 if n * h * w < 4e6 else 12
    for i in range(reps):
        a = shdr.hist_multi(d, pool=True).numpy()
        if not np.array_equal(a, ref84):
            bad += 1; print("MISMATCH 84", (n, h, w), i, np.abs(a - ref84).max())
        if ref93 is not None:
            b = shdr.frontend(d, pool=True).numpy()
            if not np.array_equal(b, ref93):
                bad += 1; print("MISMATCH 93", (n, h, w), i, np.abs(b - ref93).max())
    print("ok", (n, h, w), reps)
print("STRESS", "FAILED" if bad else "PASSED", bad)
