"""GPU parity of the reverse-mode kernels and of the fused clip / alpha-mask neighbours of
apply_rf (rank 3), through the C ABI, against the oracle (whose gradient formulas are themselves checked against
torch.autograd and finite differences in tests/test_oracle_grad.py).

Tolerances: the kernels accumulate in fp32 (and apply_rf's d/drf with fp32 atomics, so its summation order is not
deterministic); the oracle evaluates in fp64.  gx, front-end and curve gradients: 2e-5 of the gradient's scale;
d/drf: 1e-4 of its largest bin (a bin sums thousands of terms of mixed sign)."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def close(a, b, tol):
    scale = max(1e-30, float(np.abs(b).max()))
    err = float(np.abs(a - b).max())
    assert err <= tol * scale, f"max err {err:.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("shape,k", [((3, 17, 23, 3), 1024), ((2, 64, 64, 3), 1024), ((1, 5, 7, 1), 16), ((2, 1000), 333)])
def test_apply_rf_bwd(shdr_gpu, shape, k):
    rng = np.random.default_rng(sum(shape) + k)
    x = (rng.random(shape) * 1.2 - 0.1).astype(np.float32)
    rf = (np.sort(rng.random((shape[0], k)), axis=1) * 1.1).astype(np.float32)
    gy = rng.normal(size=shape).astype(np.float32)
    D = shdr_gpu.DeviceArray.from_numpy
    gx, grf = shdr_gpu.apply_rf_bwd(D(x), D(rf), D(gy))
    rx, rrf = oracle.apply_rf_grad(x, rf, gy)
    close(gx.numpy(), rx, 2e-5)
    close(grf.numpy(), rrf, 1e-4)
    only_x, none = shdr_gpu.apply_rf_bwd(D(x), D(rf), D(gy), need_grf=False)
    assert none is None and np.array_equal(only_x.numpy(), gx.numpy())
    none, only_rf = shdr_gpu.apply_rf_bwd(D(x), D(rf), D(gy), need_gx=False)
    assert none is None
    close(only_rf.numpy(), rrf, 1e-4)


def test_apply_rf_bwd_smooth_image_contention(shdr_gpu):
    """a smooth ramp makes whole warps hit the same two bins (worst case for the shared-memory atomics)"""
    x = np.tile(np.linspace(0, 1, 256 * 256 * 3, dtype=np.float32).reshape(1, 256, 256, 3), (2, 1, 1, 1))
    rf = np.tile(np.linspace(0, 1, 1024, dtype=np.float32) ** 2, (2, 1))
    gy = np.ones_like(x)
    D = shdr_gpu.DeviceArray.from_numpy
    gx, grf = shdr_gpu.apply_rf_bwd(D(x), D(rf), D(gy))
    rx, rrf = oracle.apply_rf_grad(x, rf, gy)
    close(gx.numpy(), rx, 2e-5)
    close(grf.numpy(), rrf, 1e-4)
    assert abs(float(grf.numpy().sum()) - x.size) <= 1e-3 * x.size      # the two weights of every element sum to 1


@pytest.mark.parametrize("k", [2, 7, 1024, 1025, 4096])
def test_increase_bwd(shdr_gpu, emor, k):
    _, g0, hinv = emor
    rng = np.random.default_rng(k)
    if k == 1024:
        rf = oracle.invcrf_pca_w_2_invcrf(rng.normal(0, 0.5, (5, 11)).astype(np.float32), g0, hinv)
        rf[4] = np.linspace(0, 1, 1024, dtype=np.float32) ** 2            # monotone: min > 0, no shift
    elif k == 2:                                 # a single diff: only a positive one gives a finite curve (else 0/0, as in TF)
        rf = np.array([[0.0, 0.5], [0.2, 0.9], [0.0, 1.0]], np.float32)
    else:
        rf = np.cumsum(rng.normal(0.01, 0.02, (3, k)), axis=1).astype(np.float32)
    gout = rng.normal(size=rf.shape).astype(np.float32)
    D = shdr_gpu.DeviceArray.from_numpy
    got = shdr_gpu._increase_bwd(D(rf), D(gout)).numpy()
    close(got, oracle.increase_grad(rf, gout), 5e-4 if k > 1024 else 5e-5)


def test_increase_bwd_ties(shdr_gpu):
    rf = np.array([[0.0, 0.1, 0.05, 0.3, 0.25, 0.6]], np.float32)       # two equal minima
    gout = (np.arange(6, dtype=np.float32)[None] / 3.0)
    D = shdr_gpu.DeviceArray.from_numpy
    close(shdr_gpu._increase_bwd(D(rf), D(gout)).numpy(), oracle.increase_grad(rf, gout), 1e-5)


@pytest.mark.parametrize("monotone", [False, True])
def test_invcrf_build_bwd(shdr_gpu, emor, monotone):
    _, g0, hinv = emor
    rng = np.random.default_rng(7)
    w = rng.normal(0, 0.5, (6, 11)).astype(np.float32)
    g = rng.normal(size=(6, 1024)).astype(np.float32)
    D = shdr_gpu.DeviceArray.from_numpy
    got = shdr_gpu.invcrf_build_bwd(D(w), D(g), monotone=monotone).numpy()
    if monotone:
        rf = oracle.invcrf_pca_w_2_invcrf(w, g0, hinv, np.float64)
        want = oracle.invcrf_pca_grad(oracle.increase_grad(rf, g), hinv)
    else:
        want = oracle.invcrf_pca_grad(g, hinv)
    close(got, want, 5e-5)


@pytest.mark.parametrize("shape", [(2, 9, 11, 3), (1, 2, 2, 3), (1, 3, 5, 3), (2, 64, 130, 3), (1, 300, 2, 3)])
def test_frontend_bwd(shdr_gpu, shape):
    rng = np.random.default_rng(sum(shape))
    img = rng.random(shape, dtype=np.float32)
    img[0, 0, 0] = [0.125, 0.5, 1.0]
    gfeat = rng.normal(size=shape[:3] + (93,)).astype(np.float32)
    D = shdr_gpu.DeviceArray.from_numpy
    got = shdr_gpu.frontend_bwd(D(img), D(gfeat)).numpy()
    close(got, oracle.frontend_grad(img, gfeat), 2e-5)


@pytest.mark.parametrize("B,c", [(4, 3), (5, 3), (16, 1), (33, 3)])
def test_histogram_layer_bwd(shdr_gpu, B, c):
    rng = np.random.default_rng(B * 10 + c)
    img = rng.random((2, 13, 17, c), dtype=np.float32)
    gh = rng.normal(size=(2, 13, 17, c * B)).astype(np.float32)
    D = shdr_gpu.DeviceArray.from_numpy
    got = shdr_gpu.histogram_layer_bwd(D(img), D(gh), B).numpy()
    close(got, oracle.histogram_layer_grad(img, gh, B), 2e-5)


# ---------------------------------------------------------------- fused neighbours of apply_rf
@pytest.mark.parametrize("shape", [(2, 16, 24, 3), (1, 7, 5, 3), (3, 33, 31, 3)])
@pytest.mark.parametrize("clip", [True, False])
def test_linearize_ex_from_curve(shdr_gpu, shape, clip):
    rng = np.random.default_rng(sum(shape))
    x = (rng.random(shape) * 1.4 - 0.2).astype(np.float32)
    rf = (np.sort(rng.random((shape[0], 1024)), axis=1) * 1.3).astype(np.float32)
    D = shdr_gpu.DeviceArray.from_numpy
    out = shdr_gpu.linearize_ex(D(x), rf=D(rf), clip=clip, alpha_threshold=0.12, want_clipped=True)
    c, y, a = oracle.linearize_ex(x, rf, 0.12, clip=clip)
    assert np.array_equal(out["clipped"].numpy(), c)
    assert np.array_equal(out["y"].numpy(), y)               # same rounding sequence as apply_rf: bit-exact
    assert np.array_equal(out["alpha"].numpy(), a)
    plain = shdr_gpu.linearize_ex(D(x), rf=D(rf), clip=clip)
    assert set(plain) == {"y"} and np.array_equal(plain["y"].numpy(), y)


def test_linearize_ex_from_weights(shdr_gpu, emor):
    _, g0, hinv = emor
    rng = np.random.default_rng(12)
    x = (rng.random((4, 32, 48, 3)) * 1.2 - 0.1).astype(np.float32)
    w = rng.normal(0, 0.5, (4, 11)).astype(np.float32)
    D = shdr_gpu.DeviceArray.from_numpy
    out = shdr_gpu.linearize_ex(D(x), invcrf_pca_w=D(w), alpha_threshold=0.12, want_clipped=True)
    curve = out["curve"].numpy()
    assert np.abs(curve - oracle.increase(oracle.invcrf_pca_w_2_invcrf(w, g0, hinv))).max() <= 5e-6
    c, y, a = oracle.linearize_ex(x, curve, 0.12)
    assert np.array_equal(out["clipped"].numpy(), c) and np.array_equal(out["y"].numpy(), y)
    assert np.array_equal(out["alpha"].numpy(), a)
    y2, _ = shdr_gpu.linearize(D(c), D(w))
    assert np.array_equal(y2.numpy(), y)


# ---------------------------------------------------------------- synthetic-LDR generator
@pytest.mark.parametrize("shape,k", [((3, 20, 28, 3), 1024), ((1, 7, 5, 3), 1024), ((2, 16, 16, 3), 256)])
def test_synth_ldr(shdr_gpu, shape, k):
    rng = np.random.default_rng(sum(shape) + k)
    b = shape[0]
    hdr = (rng.random(shape) ** 3 * 4.0).astype(np.float32)                  # HDR radiance, some above 1 after exposure
    t = (2.0 ** rng.uniform(-3, 3, b)).astype(np.float32)
    ss = (0.08 / 6 * rng.random((b, 3))).astype(np.float32)
    sc = (0.005 * rng.random((b, 3))).astype(np.float32)
    ns = rng.normal(size=shape).astype(np.float32)
    nc = rng.normal(size=shape).astype(np.float32)
    crf = (np.linspace(0, 1, k, dtype=np.float32)[None] ** rng.uniform(0.3, 0.8, (b, 1))).astype(np.float32)
    D = shdr_gpu.DeviceArray.from_numpy
    out = shdr_gpu.synth_ldr(D(hdr), D(t), D(ss), D(sc), D(ns), D(nc), D(crf))
    x, c, l, q = oracle.synth_ldr(hdr, t, ss, sc, ns, nc, crf)
    assert np.array_equal(out["hdr_t"].numpy(), x)          # same op order and roundings: bit-exact
    assert np.array_equal(out["clipped"].numpy(), c)
    assert np.array_equal(out["ldr"].numpy(), l)
    assert np.array_equal(out["quant"].numpy(), q)
    assert q.min() >= 0 and q.max() <= 255
    only = shdr_gpu.synth_ldr(D(hdr), D(t), D(ss), D(sc), D(ns), D(nc), D(crf), outputs=("ldr",))
    assert set(only) == {"ldr"} and np.array_equal(only["ldr"].numpy(), l)
