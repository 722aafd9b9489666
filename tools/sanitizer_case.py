import sys, numpy as np
sys.path.insert(0, '/root/repo')
import shdr, oracle
z = np.load('/root/repo/tests/golden/invemor_f32.npz'); shdr.set_emor_table(z['g0'], z['hinv'])
rng = np.random.default_rng(0)
img = rng.random((2, 40, 130, 3), dtype=np.float32)
d = shdr.DeviceArray.from_numpy(img)
a = shdr.hist_multi(d, pool=True).numpy()          # warp-specialised kernel (even width)
b = shdr.frontend(d).numpy()                       # strip kernel
c = shdr.frontend(d, pool=True).numpy()            # block kernel (93-ch slice)
img2 = rng.random((1, 19, 33, 3), dtype=np.float32)
e = shdr.hist_multi(shdr.DeviceArray.from_numpy(img2), pool=True).numpy()   # odd width -> block kernel
w = rng.normal(0, .5, (2, 11)).astype(np.float32)
y, cv = shdr.linearize(d, shdr.DeviceArray.from_numpy(w))
y = y.numpy()
print('ok', a.shape, b.shape, c.shape, e.shape, y.shape, float(np.abs(a - oracle.hist_multi(img, pool_k=16)).max()))
