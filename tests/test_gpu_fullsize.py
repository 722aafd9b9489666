"""BASELINE.json's full-size configurations on the GPU, checked through size-independent
properties (the oracle is too slow there) plus oracle parity on a sampled image."""
import numpy as np
import pytest

import oracle

from oracle import c_oracle

pytestmark = pytest.mark.gpu

# The C restatement of the oracle (bit-identical to the NumPy one, tests/test_oracle_c.py) is fast enough to check
# EVERY element at BASELINE.json's full sizes; without it the NumPy oracle checks one image per configuration.
HAVE_C = c_oracle.available()


def rnd(shape, seed=0):
    return np.random.default_rng(seed).random(shape, dtype=np.float32)


def test_config2_pooled_hist_full(shdr_gpu):
    """32 x 512 x 512: partition of unity survives the pool; one image checked against the oracle."""
    img = rnd((32, 512, 512, 3), 1)
    img = np.float32(0.125) + img * np.float32(0.75)    # inside [1/2B, 1-1/2B] for B = 4, 8, 16
    out = shdr_gpu.hist_multi(shdr_gpu.DeviceArray.from_numpy(img), pool=True).numpy()
    assert out.shape == (32, 512, 512, 84) and out.min() >= 0 and out.max() <= 1
    for off, B in ((0, 4), (12, 8), (36, 16)):
        s = out[..., off:off + 3 * B].reshape(32, 512, 512, B, 3).sum(3)
        assert np.abs(s - 1).max() < 2e-6          # sum over bins of the pooled votes == 1
    ref = oracle.hist_multi(img[7:8], pool_k=16)
    assert np.all(np.abs(out[7:8] - ref) <= 1e-5 * np.abs(ref))
    if HAVE_C:                                     # all 32 x 512 x 512 x 84 elements, pure relative gate
        full = c_oracle.hist_multi(img, pool_k=16)
        assert np.array_equal(full[7:8], ref)
        assert np.all(np.abs(out - full) <= 1e-5 * np.abs(full))


def test_config3_apply_full(shdr_gpu, emor):
    """16 x 1024 x 1024: identity curve returns x, monotone curve keeps order, sampled oracle parity."""
    _, g0, hinv = emor
    x = rnd((16, 1024, 1024, 3), 2)
    w = np.random.default_rng(3).normal(0, 0.5, (16, 11)).astype(np.float32)
    dx = shdr_gpu.DeviceArray.from_numpy(x)
    ident = np.tile(np.linspace(0, 1, 1024, dtype=np.float32), (16, 1))
    y = shdr_gpu.apply_rf(dx, shdr_gpu.DeviceArray.from_numpy(ident)).numpy()
    assert np.abs(y - x).max() <= 2.4e-7
    y, curve = shdr_gpu.linearize(dx, shdr_gpu.DeviceArray.from_numpy(w))
    y, curve = y.numpy(), curve.numpy()
    rc = oracle.increase(oracle.invcrf_pca_w_2_invcrf(w, g0, hinv))
    assert np.abs(curve - rc).max() <= 5e-6
    for b in (0, 15):
        order = np.argsort(x[b].ravel(), kind="stable")
        assert np.all(np.diff(y[b].ravel()[order]) >= -6e-7)      # monotone curve -> order preserved (up to the 3 roundings of the lerp)
    assert np.abs(y[5] - oracle.apply_rf(x[5:6], rc[5:6])[0]).max() <= 1e-5
    assert np.array_equal(y[5], oracle.apply_rf(x[5:6], curve[5:6])[0])   # bit-exact given the same curve
    if HAVE_C:                                     # every element of the 16 x 1024 x 1024 x 3 batch
        assert np.abs(y - c_oracle.apply_rf(x, rc)).max() <= 1e-5
        assert np.array_equal(y, c_oracle.apply_rf(x, curve))


def test_config4_frontend_full(shdr_gpu):
    """8 x 512 x 512 front end: slices equal the stand-alone layers; one image vs the oracle."""
    img = rnd((8, 512, 512, 3), 4)
    d = shdr_gpu.DeviceArray.from_numpy(img)
    f = shdr_gpu.frontend(d).numpy()
    assert np.array_equal(f[..., :3], img)
    assert np.array_equal(f[..., 3:9], shdr_gpu.sobel_edges6(d).numpy())
    assert np.array_equal(f[..., 9:21], shdr_gpu.histogram_layer(d, 4).numpy())
    assert np.array_equal(f[..., 21:45], shdr_gpu.histogram_layer(d, 8).numpy())
    assert np.array_equal(f[..., 45:], shdr_gpu.histogram_layer(d, 16).numpy())
    assert np.array_equal(f[3], oracle.frontend(img[3:4])[0])
    if HAVE_C:                                     # every element of the 8 x 512 x 512 x 93 tensor, bit for bit
        assert np.array_equal(f, c_oracle.frontend(img))


def test_config5_4k_frame(shdr_gpu, emor):
    """One 3840x2160 frame: front end + linearize; row-tile sharding gives the same bytes."""
    _, g0, hinv = emor
    img = rnd((1, 2160, 3840, 3), 5)
    d = shdr_gpu.DeviceArray.from_numpy(img)
    f = shdr_gpu.frontend(d).numpy()
    rows = slice(1000, 1016)
    ref = oracle.frontend(img[:, 999:1017])[:, 1:-1]          # interior rows: halo rows are real data
    assert np.array_equal(f[:, rows], ref)
    if HAVE_C:                                     # the whole 2160 x 3840 x 93 frame, bit for bit
        assert np.array_equal(f, c_oracle.frontend(img))
    # tile sharding: 4 row tiles with a 1-row read-only halo reproduce the frame
    for (y0, y1, i0, i1) in shdr_gpu.row_tiles(2160, 4, 1, 1)[1:3]:
        part = shdr_gpu.frontend(shdr_gpu.DeviceArray.from_numpy(img[:, i0:i1])).numpy()
        assert np.array_equal(part[:, y0 - i0:y1 - i0, :, 9:], f[:, y0:y1, :, 9:])
        assert np.array_equal(part[:, y0 - i0:y1 - i0, :, :9], f[:, y0:y1, :, :9])
    w = np.random.default_rng(99).normal(0, 0.5, (1, 11)).astype(np.float32)
    y, curve = shdr_gpu.linearize(d, shdr_gpu.DeviceArray.from_numpy(w))
    assert np.array_equal(y.numpy()[:, :64], oracle.apply_rf(img[:, :64], curve.numpy()))


@pytest.mark.parametrize("shape", [(1, 2160, 3840, 3), (2, 1080, 1922, 3), (1, 1000, 66, 3)])
def test_pooled_large_frames(shdr_gpu, shape):
    """Pooled histograms on large / awkward frames (4K; even width that is no multiple of the 64-pixel tile; tall and
    narrow), every element against the C oracle."""
    if not HAVE_C:
        pytest.skip("C oracle not built")
    img = rnd(shape, 17 + shape[2])
    out = shdr_gpu.hist_multi(shdr_gpu.DeviceArray.from_numpy(img), pool=True).numpy()
    ref = c_oracle.hist_multi(img, pool_k=16)
    assert np.all(np.abs(out - ref) <= 1e-5 * np.abs(ref))


def test_config4p_pooled_frontend_full(shdr_gpu):
    """8 x 512 x 512 pooled 93-channel front end (one sliding-window launch): img / Sobel slices bit-exact, pooled
    histograms within 1e-5 pure relative of the C oracle on every element, and equal to the 84-channel kernel."""
    if not HAVE_C:
        pytest.skip("C oracle not built")
    img = rnd((8, 512, 512, 3), 44)
    d = shdr_gpu.DeviceArray.from_numpy(img)
    f = shdr_gpu.frontend(d, pool=True).numpy()
    ref = c_oracle.frontend(img, pool_k=16)
    assert np.array_equal(f[..., :9], ref[..., :9])
    assert np.all(np.abs(f[..., 9:] - ref[..., 9:]) <= 1e-5 * np.abs(ref[..., 9:]))
    assert np.array_equal(f[..., 9:], shdr_gpu.hist_multi(d, pool=True).numpy())


def test_config4_frontend_conv1_full(shdr_gpu):
    """configs[3] shape, 8 x 512 x 512, through the fused front end + conv1 kernel (4096 tiles, 28 iterations of
    every CTA pair).  The convolution is local (an output pixel sees the input within +-6 pixels), so the oracle is
    run on crops: the interior of a crop's result must equal the full-size result, and crops touching the image
    corners check the 'SAME' padding and the REFLECT Sobel at the border."""
    rng = np.random.default_rng(44)
    img = rng.random((8, 512, 512, 3), dtype=np.float32)
    kern = (rng.normal(0, 1, (7, 7, 93, 64)) / 67.5).astype(np.float32)
    bias = rng.normal(0, 0.1, 64).astype(np.float32)
    D = shdr_gpu.DeviceArray.from_numpy
    out = shdr_gpu.frontend_conv1(D(img), shdr_gpu.conv1_pack_weights(D(kern)), bias=D(bias)).numpy()
    assert out.shape == (8, 256, 256, 64) and np.isfinite(out).all()
    scale = np.abs(out).max()
    M = 8                                             # even margin (input pixels) dropped around an interior crop
    boxes = [(0, 0, 0), (3, 0, 416), (5, 416, 0), (7, 416, 416)]                     # (image, y0, x0): the four corners
    boxes += [(int(rng.integers(0, 8)), int(rng.integers(8, 200)) * 2, int(rng.integers(8, 200)) * 2) for _ in range(6)]
    for n, y0, x0 in boxes:
        crop = img[n:n + 1, y0:y0 + 96, x0:x0 + 96]
        ref = oracle.frontend_conv1(crop, kern, bias, half_operands=True)          # [1, 48, 48, 64]
        ya, yb = (0 if y0 == 0 else M // 2), (48 if y0 + 96 == 512 else 48 - M // 2)
        xa, xb = (0 if x0 == 0 else M // 2), (48 if x0 + 96 == 512 else 48 - M // 2)
        got = out[n, y0 // 2 + ya:y0 // 2 + yb, x0 // 2 + xa:x0 // 2 + xb]
        assert np.abs(got - ref[0, ya:yb, xa:xb]).max() <= 2e-5 * scale, (n, y0, x0)
