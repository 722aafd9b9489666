"""GPU parity tests: the sm_100a kernels, called through the C ABI (via the ctypes host layer),
against the oracle on identical seeded inputs.

Tolerances are the ones BASELINE.json's north_star states:
    histogram / edges   <= 1e-6 abs     (the kernels are in fact bit-exact to the oracle)
    pooled histogram    <= 1e-5 rel     (pure relative: |a-b| <= 1e-5*|b|, zeros must be zeros)
    curve               <= 5e-6 abs     (block scan vs sequential cumsum)
    linearised image    <= 1e-5 abs
"""
import ctypes as C

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

ATOL_FEAT = 1e-6
RTOL_POOL = 1e-5
ATOL_CURVE = 5e-6
ATOL_LIN = 1e-5


def rnd(shape, seed=0):
    return np.random.default_rng(seed).random(shape, dtype=np.float32)


def assert_rel(a, b, rtol):
    bad = np.abs(a - b) > rtol * np.abs(b)
    assert not bad.any(), f"{bad.sum()} elements off; worst abs {np.abs(a - b).max()}"


# ---------------------------------------------------------------- front end
@pytest.mark.parametrize("shape", [(1, 2, 2, 3), (1, 256, 256, 3), (3, 17, 31, 3), (2, 64, 130, 3),
                                   (1, 5, 1027, 3), (1, 300, 2, 3)])
def test_frontend_unpooled(shdr_gpu, shape):
    img = rnd(shape, sum(shape))
    got = shdr_gpu.frontend(shdr_gpu.DeviceArray.from_numpy(img)).numpy()
    ref = oracle.frontend(img)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= ATOL_FEAT
    assert np.array_equal(got[..., :3], img)
    assert np.array_equal(got[..., 9:], ref[..., 9:])        # histograms are bit-exact
    assert np.array_equal(got[..., 3:9], ref[..., 3:9])      # same tap order -> edges too


def test_frontend_quantised_input(shdr_gpu):
    """8-bit LDR values (what test_real_refinement.py:126 ingests) hit bin centres/edges exactly."""
    img = np.round(rnd((1, 64, 64, 3), 5) * 255).astype(np.float32) / np.float32(255)
    img[0, 0, :4, 0] = [0.0, 1.0, 0.5, 0.25]
    got = shdr_gpu.frontend(shdr_gpu.DeviceArray.from_numpy(img)).numpy()
    assert np.array_equal(got, oracle.frontend(img))


@pytest.mark.parametrize("c", [1, 3, 4])
def test_sobel(shdr_gpu, c):
    img = rnd((2, 19, 23, c), c)
    got = shdr_gpu.sobel_edges6(shdr_gpu.DeviceArray.from_numpy(img)).numpy()
    ref = oracle.sobel_edges6(img)
    assert got.shape == (2, 19, 23, 2 * c)
    assert np.array_equal(got, ref)
    assert np.abs(got - oracle.sobel_edges6(img, np.float64)).max() <= ATOL_FEAT


@pytest.mark.parametrize("B", [1, 3, 4, 5, 8, 16, 33])
@pytest.mark.parametrize("c", [1, 3])
def test_histogram_layer(shdr_gpu, B, c):
    img = rnd((2, 21, 37, c), B * 10 + c)
    img[0, 0, :3, 0] = [0.0, 1.0, 0.5]
    got = shdr_gpu.histogram_layer(shdr_gpu.DeviceArray.from_numpy(img), B).numpy()
    ref = oracle.histogram_layer(img, B)
    assert got.shape == ref.shape and np.array_equal(got, ref)


def test_histogram_kat_lin2(shdr_gpu, kat_lin2):
    v = kat_lin2["values"].reshape(1, 1, 5, 1)
    got = shdr_gpu.histogram_layer(shdr_gpu.DeviceArray.from_numpy(v), 5).numpy()[0, 0]
    np.testing.assert_allclose(got, kat_lin2["votes"], atol=2e-7)


def test_hist_multi_unpooled(shdr_gpu):
    img = rnd((2, 33, 65, 3), 11)
    got = shdr_gpu.hist_multi(shdr_gpu.DeviceArray.from_numpy(img)).numpy()
    assert np.array_equal(got, oracle.hist_multi(img))


def test_out_of_range_and_nan_inputs(shdr_gpu):
    img = rnd((1, 8, 8, 3), 12)
    img[0, 0, 0] = [-0.5, 1.5, np.nan]
    img[0, 1, 1] = [np.inf, -np.inf, 2.0]
    got = shdr_gpu.hist_multi(shdr_gpu.DeviceArray.from_numpy(img)).numpy()
    ref = oracle.hist_multi(img)
    assert np.array_equal(got, ref)           # NaN / out-of-range vote 0 everywhere, like tf.where
    gp = shdr_gpu.hist_multi(shdr_gpu.DeviceArray.from_numpy(img), pool=True).numpy()
    assert np.isfinite(gp).all()


# ---------------------------------------------------------------- pooled
@pytest.mark.parametrize("shape", [(1, 16, 64, 3), (2, 40, 72, 3), (1, 3, 5, 3), (1, 1, 1, 3),
                                   (1, 15, 200, 3), (2, 67, 129, 3), (1, 128, 128, 3)])
def test_hist_multi_pooled(shdr_gpu, shape):
    img = rnd(shape, sum(shape) + 1)
    got = shdr_gpu.hist_multi(shdr_gpu.DeviceArray.from_numpy(img), pool=True).numpy()
    ref = oracle.hist_multi(img, pool_k=16)
    assert got.shape == ref.shape
    assert_rel(got, ref, RTOL_POOL)
    t = oracle.hist_multi(img, pool_k=16, dtype=np.float64)
    assert_rel(got, t, RTOL_POOL)


@pytest.mark.parametrize("shape", [(8, 48, 512, 3), (3, 80, 1024, 3)])
def test_histogram_layer_pooled_b4_border_interior_mix(shdr_gpu, shape):
    """B = 4 alone is the two-unit case of the tiled kernel (pooled_ws.cu): many tiles per CTA, border and interior
    tiles alternating, so that the double-buffered border-scale table is reused while slow warps may still read it."""
    img = rnd(shape, 99)
    for _ in range(3):
        got = shdr_gpu.histogram_layer(shdr_gpu.DeviceArray.from_numpy(img), 4, pool=True).numpy()
        ref = oracle.avg_pool_same(oracle.histogram_layer(img, 4), 16)
        assert_rel(got, ref, RTOL_POOL)


@pytest.mark.parametrize("B", [4, 5, 16, 21])
def test_histogram_layer_pooled(shdr_gpu, B):
    img = rnd((1, 37, 70, 3), B)
    got = shdr_gpu.histogram_layer(shdr_gpu.DeviceArray.from_numpy(img), B, pool=True).numpy()
    ref = oracle.avg_pool_same(oracle.histogram_layer(img, B), 16)
    assert_rel(got, ref, RTOL_POOL)


@pytest.mark.parametrize("shape", [(1, 16, 2, 3), (2, 17, 66, 3), (1, 33, 126, 3), (3, 50, 190, 3), (1, 9, 64, 3),
                                   (1, 64, 62, 3), (1, 31, 256, 3)])
def test_hist_multi_pooled_ws_tile_edges(shdr_gpu, shape):
    """Even widths take the warp-specialised whole-sector kernel: partial tiles, 1-tile images, many tiles."""
    img = rnd(shape, sum(shape) + 7)
    got = shdr_gpu.hist_multi(shdr_gpu.DeviceArray.from_numpy(img), pool=True).numpy()
    assert_rel(got, oracle.hist_multi(img, pool_k=16), RTOL_POOL)


@pytest.mark.parametrize("B", [8, 16, 32, 64])
def test_histogram_layer_pooled_pow2(shdr_gpu, B):
    """C = 3B channels: 24, 48 (compile-time pitch, every pixel sector-aligned), 96, 192 (runtime pitch)."""
    img = rnd((2, 35, 130, 3), B)
    got = shdr_gpu.histogram_layer(shdr_gpu.DeviceArray.from_numpy(img), B, pool=True).numpy()
    assert_rel(got, oracle.avg_pool_same(oracle.histogram_layer(img, B), 16), RTOL_POOL)


def test_pooled_ws_generic_channel_counts(shdr_gpu):
    """Channel counts other than 84 through the C ABI: C = 60 (4|16: odd half-sector pitch, runtime), C = 36 (4|8)."""
    from shdr import _native as N
    img = rnd((2, 40, 96, 3), 77)
    d = shdr_gpu.DeviceArray.from_numpy(img)
    for bins in ((4, 16), (4, 8), (8, 16)):
        C = 3 * sum(bins)
        out = shdr_gpu.DeviceArray.from_numpy(np.full((2, 40, 96, C), -1, np.float32))
        off = 0
        ref = []
        for B in bins:      # per-histogram calls into channel slices (block kernel) as the reference layout
            ref.append(oracle.avg_pool_same(oracle.histogram_layer(img, B), 16))
        ref = np.concatenate(ref, -1)
        for B in bins:
            N.check(N.lib.shdr_soft_hist_f32(d.ptr, out.ptr, 2, 40, 96, 3, B, 16, C, off, None))
            off += 3 * B
        assert_rel(out.numpy(), ref, RTOL_POOL)


def test_pooled_matches_block_kernel_bitwise(shdr_gpu):
    """The warp-specialised kernel and the generic block kernel sum in the same order: identical interior bits."""
    import os, subprocess, sys
    img = rnd((1, 64, 128, 3), 91)
    a = shdr_gpu.hist_multi(shdr_gpu.DeviceArray.from_numpy(img), pool=True).numpy()
    code = ("import sys, numpy as np; sys.path.insert(0, %r); import shdr; "
            "img = np.random.default_rng(91).random((1, 64, 128, 3), dtype=np.float32); "
            "np.save(sys.argv[1], shdr.hist_multi(shdr.DeviceArray.from_numpy(img), pool=True).numpy())"
            % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    path = "/tmp/shdr_block_kernel.npy"
    env = dict(os.environ, SHDR_POOL_BLOCK="1")
    subprocess.run([sys.executable, "-c", code, path], check=True, env=env)
    b = np.load(path)
    assert np.array_equal(a[:, 8:-8, 8:-8], b[:, 8:-8, 8:-8])
    assert_rel(a, b, 1e-6)


def test_pooled_sparse_image_zeros_stay_zero(shdr_gpu):
    """A window with no vote must give exactly 0 (pure relative gate)."""
    img = np.full((1, 64, 96, 3), 0.03, np.float32)       # only bins near 0 vote
    img[0, 30:34, 40:44] = 0.97
    got = shdr_gpu.hist_multi(shdr_gpu.DeviceArray.from_numpy(img), pool=True).numpy()
    ref = oracle.hist_multi(img, pool_k=16)
    assert np.array_equal(got == 0, ref == 0)
    assert_rel(got, ref, RTOL_POOL)


def test_pool_border_counts(shdr_gpu):
    """Constant image -> pooled histogram equals the un-pooled one everywhere iff the divide uses
    the in-bounds count (81 at the top-left corner, 64 at bottom-right, 256 inside)."""
    img = np.full((1, 40, 50, 3), 0.3, np.float32)
    got = shdr_gpu.hist_multi(shdr_gpu.DeviceArray.from_numpy(img), pool=True).numpy()
    ref = oracle.hist_multi(img)
    np.testing.assert_allclose(got, ref, rtol=1e-6, atol=0)


def test_frontend_pooled(shdr_gpu, golden_small):
    img = golden_small["img"]
    got = shdr_gpu.frontend(shdr_gpu.DeviceArray.from_numpy(img), pool=True).numpy()
    ref = golden_small["frontend_pooled"]
    assert np.array_equal(got[..., :9], ref[..., :9])
    assert_rel(got[..., 9:], ref[..., 9:], RTOL_POOL)


@pytest.mark.parametrize("shape", [(2, 40, 70, 3), (1, 33, 131, 3), (1, 70, 65, 3)])
def test_frontend_pooled_width_not_multiple_of_4(shdr_gpu, shape):
    """93 channels x a width that is not a multiple of 4: output rows are not 16-byte aligned chunks, so the
    sliding-window kernel writes them with cooperative 4-byte stores instead of bulk copies (same kernel, same values)."""
    img = rnd(shape, sum(shape))
    before = shdr_gpu.launch_count()
    got = shdr_gpu.frontend(shdr_gpu.DeviceArray.from_numpy(img), pool=True).numpy()
    assert shdr_gpu.launch_count() - before == 1           # ONE launch, not copy + Sobel + generic pooled
    ref = oracle.frontend(img, pool_k=16)
    assert np.array_equal(got[..., :9], ref[..., :9])
    assert_rel(got[..., 9:], ref[..., 9:], RTOL_POOL)


# ---------------------------------------------------------------- inverse CRF
def test_invcrf_build_and_increase(shdr_gpu, emor):
    _, g0, hinv = emor
    w = np.random.default_rng(3).normal(0, 0.5, (16, 11)).astype(np.float32)
    dw = shdr_gpu.DeviceArray.from_numpy(w)
    pca = shdr_gpu.invcrf_pca_w_2_invcrf(dw).numpy()
    ref_pca = oracle.invcrf_pca_w_2_invcrf(w, g0, hinv)
    assert np.abs(pca - ref_pca).max() <= 1e-6
    curve = shdr_gpu._increase(shdr_gpu.DeviceArray.from_numpy(ref_pca)).numpy()
    ref = oracle.increase(ref_pca)
    assert np.abs(curve - ref).max() <= ATOL_CURVE
    assert (curve[:, 0] == 0).all() and np.all(np.diff(curve, axis=1) >= 0)
    assert np.abs(curve[:, -1] - 1).max() < 2e-6
    fused = shdr_gpu.invcrf_build(dw, monotone=True).numpy()
    assert np.abs(fused - ref).max() <= ATOL_CURVE
    t = oracle.increase(oracle.invcrf_pca_w_2_invcrf(w, g0, hinv, np.float64), np.float64)
    assert np.abs(fused - t).max() <= ATOL_CURVE


def test_w_zero_is_g0(shdr_gpu, emor):
    _, g0, _ = emor
    c = shdr_gpu.invcrf_pca_w_2_invcrf(shdr_gpu.DeviceArray.from_numpy(np.zeros((3, 11), np.float32))).numpy()
    assert np.array_equal(c, np.stack([g0] * 3))


@pytest.mark.parametrize("k", [2, 3, 64, 1000, 1024, 1025, 4099, 49152])
def test_increase_generic_k(shdr_gpu, k):
    rf = np.cumsum(np.random.default_rng(k).normal(0.2, 1.0, (3, k)), axis=1).astype(np.float32)
    if k == 2:
        rf[0] = [0.25, 0.75]          # rising: -> [0, 1]
        rf[1] = [0.75, 0.25]          # falling: g + relu(-g) = 0 -> 0/0 = NaN, like the reference
    got = shdr_gpu._increase(shdr_gpu.DeviceArray.from_numpy(rf)).numpy()
    ref = oracle.increase(rf.astype(np.float64), np.float64)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert np.abs(got[ok] - ref[ok]).max() <= ATOL_CURVE * max(1, k // 4096)
    assert (got[:, 0] == 0).all()


def test_increase_k_limit(shdr_gpu):
    with pytest.raises(shdr_gpu.ShdrError, match="49152"):
        shdr_gpu._increase(shdr_gpu.DeviceArray.from_numpy(np.zeros((1, 49153), np.float32)))


def test_increase_constant_curve_nan_like_reference(shdr_gpu):
    got = shdr_gpu._increase(shdr_gpu.DeviceArray.from_numpy(np.full((1, 1024), 0.3, np.float32))).numpy()
    assert got[0, 0] == 0 and np.isnan(got[0, 1:]).all()


@pytest.mark.parametrize("shape,k", [((2, 16, 16, 3), 1024), ((3, 7, 5, 3), 1024), ((1, 1), 1024),
                                     ((2, 33, 3), 256), ((2, 1000), 7), ((1, 4096), 40000), ((2, 9), 1)])
def test_apply_rf(shdr_gpu, shape, k):
    x = rnd(shape, k) * 1.2 - 0.1                       # includes x < 0 and x > 1
    rf = np.sort(rnd((shape[0], k), k + 1), axis=1)
    got = shdr_gpu.apply_rf(shdr_gpu.DeviceArray.from_numpy(x), shdr_gpu.DeviceArray.from_numpy(rf)).numpy()
    ref = oracle.apply_rf(x, rf)
    assert got.shape == x.shape
    assert np.array_equal(got, ref)                     # same rounding sequence -> bit-exact


def test_apply_rf_special_values(shdr_gpu):
    rf = np.sort(rnd((1, 1024), 3), axis=1)
    x = np.float32([[0.0, 1.0, -0.0, 1023 / 1024, 1 / 1023, 0.5, -3.0, 7.0, 1e30, -1e30, 1e-40]])
    got = shdr_gpu.apply_rf(shdr_gpu.DeviceArray.from_numpy(x), shdr_gpu.DeviceArray.from_numpy(rf)).numpy()
    ref = oracle.apply_rf(x, rf)
    assert np.array_equal(got, ref)
    assert got[0, 0] == rf[0, 0] and got[0, 1] == rf[0, -1]


def test_linearize_config3_inputs_small(shdr_gpu, emor):
    """Config 3's distribution (w ~ N(0, 0.5), U[0,1) pixels) at a size the oracle does in seconds."""
    _, g0, hinv = emor
    x = rnd((4, 128, 128, 3), 2)
    w = np.random.default_rng(3).normal(0, 0.5, (4, 11)).astype(np.float32)
    y, curve = shdr_gpu.linearize(shdr_gpu.DeviceArray.from_numpy(x), shdr_gpu.DeviceArray.from_numpy(w))
    ry, rc = oracle.linearize(x, w, g0, hinv)
    assert np.abs(curve.numpy() - rc).max() <= ATOL_CURVE
    assert np.abs(y.numpy() - ry).max() <= ATOL_LIN


def test_golden_fixtures(shdr_gpu, golden_small):
    g = golden_small
    d = shdr_gpu.DeviceArray.from_numpy
    assert np.array_equal(shdr_gpu.sobel_edges6(d(g["img"])).numpy(), g["edges"])
    assert np.array_equal(shdr_gpu.histogram_layer(d(g["img"]), 4).numpy(), g["hist4"])
    assert np.array_equal(shdr_gpu.histogram_layer(d(g["img"]), 5).numpy(), g["hist5"])
    assert np.abs(shdr_gpu.frontend(d(g["img"])).numpy() - g["frontend"]).max() <= ATOL_FEAT
    assert_rel(shdr_gpu.histogram_layer(d(g["img"]), 16, pool=True).numpy(), g["hist16_pooled"], RTOL_POOL)
    y, curve = shdr_gpu.linearize(d(g["x"]), d(g["w"]))
    assert np.abs(curve.numpy() - g["curve"]).max() <= ATOL_CURVE
    assert np.abs(y.numpy() - g["lin"]).max() <= ATOL_LIN


# ---------------------------------------------------------------- channel-slice outputs (fused concat)
def test_strided_outputs_compose_the_concat(shdr_gpu):
    from shdr import _native as N
    img = rnd((1, 24, 40, 3), 21)
    d_img = shdr_gpu.DeviceArray.from_numpy(img)
    out = shdr_gpu.DeviceArray.from_numpy(np.full((1, 24, 40, 93), -1, np.float32))
    N.check(N.lib.shdr_sobel6_f32(d_img.ptr, out.ptr, 1, 24, 40, 3, 93, 3, None))
    off = 9
    for B in (4, 8, 16):
        N.check(N.lib.shdr_soft_hist_f32(d_img.ptr, out.ptr, 1, 24, 40, 3, B, 0, 93, off, None))
        off += 3 * B
    got = out.numpy()
    ref = oracle.frontend(img)
    assert np.array_equal(got[..., 3:], ref[..., 3:]) and (got[..., :3] == -1).all()


# ---------------------------------------------------------------- host-buffer (e2e) API and DLPack
def test_host_api(shdr_gpu, emor):
    _, g0, hinv = emor
    img = rnd((5, 48, 80, 3), 31)
    assert np.array_equal(shdr_gpu.frontend_host(img), oracle.frontend(img))
    assert_rel(shdr_gpu.hist_multi_host(img, pool=True), oracle.hist_multi(img, pool_k=16), RTOL_POOL)
    w = np.random.default_rng(4).normal(0, 0.5, (5, 11)).astype(np.float32)
    y, curve = shdr_gpu.linearize_host(img, w)
    ry, rc = oracle.linearize(img, w, g0, hinv)
    assert np.abs(curve - rc).max() <= ATOL_CURVE and np.abs(y - ry).max() <= ATOL_LIN
    assert np.array_equal(shdr_gpu.apply_rf_host(img, rc), oracle.apply_rf(img, rc))


def test_host_pipeline_many_chunks(shdr_gpu):
    from shdr import _native as N
    img = rnd((7, 16, 24, 3), 33)
    out = np.empty((7, 16, 24, 93), np.float32)

    def op(d_in, d_out, m, st, _i0):
        N.check(N.lib.shdr_frontend_f32(d_in, d_out, m, 16, 24, 0, st))

    p = shdr_gpu.HostPipeline(op, 16 * 24 * 3 * 4, 16 * 24 * 93 * 4, items_per_chunk=2, slots=2)
    p.run(img, out, 7)
    p.close()
    assert np.array_equal(out, oracle.frontend(img))


def test_dlpack_torch_roundtrip(shdr_gpu):
    """torch CUDA tensors stand in for TF eager tensors: same DLPack capsule protocol."""
    import torch
    img = rnd((2, 32, 48, 3), 41)
    t = torch.from_numpy(img).cuda()
    out = shdr_gpu.frontend(t)                       # borrowed zero-copy via __dlpack__
    back = torch.from_dlpack(out)                    # consumer takes ownership, zero-copy
    assert back.is_cuda and tuple(back.shape) == (2, 32, 48, 93)
    assert np.array_equal(back.cpu().numpy(), oracle.frontend(img))
    with pytest.raises(RuntimeError):
        out.numpy()                                  # exported: one-shot ownership
    cap = torch.utils.dlpack.to_dlpack(t)            # raw capsule path
    e = shdr_gpu.sobel_edges6(cap).numpy()
    assert np.array_equal(e, oracle.sobel_edges6(img))
    nc = t.permute(0, 3, 1, 2)                       # non-compact view must be rejected, not mis-read
    with pytest.raises(shdr_gpu.ShdrError, match="compact"):
        shdr_gpu.frontend(nc)
    with pytest.raises(shdr_gpu.ShdrError, match="float32"):
        shdr_gpu.frontend(t.double())


def test_error_paths_on_gpu(shdr_gpu):
    d = shdr_gpu.DeviceArray.from_numpy
    with pytest.raises(shdr_gpu.ShdrError):
        shdr_gpu.frontend(d(np.zeros((1, 1, 8, 3), np.float32)))            # h < 2: REFLECT undefined
    with pytest.raises(shdr_gpu.ShdrError):
        shdr_gpu.frontend(d(np.zeros((1, 8, 8, 4), np.float32)))            # c != 3
    with pytest.raises(shdr_gpu.ShdrError):
        shdr_gpu.histogram_layer(d(np.zeros((1, 8, 8, 4), np.float32)), 4, pool=True)
    with pytest.raises(shdr_gpu.ShdrError):
        shdr_gpu.apply_rf(d(np.zeros((2, 8), np.float32)), d(np.zeros((3, 16), np.float32)))
    assert shdr_gpu.frontend(d(np.zeros((0, 8, 8, 3), np.float32))).numpy().shape == (0, 8, 8, 93)


def test_streams_and_events(shdr_gpu):
    img = rnd((2, 64, 64, 3), 51)
    s = shdr_gpu.Stream()
    e0, e1 = shdr_gpu.Event(), shdr_gpu.Event()
    d = shdr_gpu.DeviceArray.from_numpy(img)
    e0.record(s)
    out = shdr_gpu.frontend(d, stream=s)
    e1.record(s)
    assert e0.elapsed_ms(e1) >= 0
    s.sync()
    assert np.array_equal(out.numpy(), oracle.frontend(img))


# ---------------------------------------------------------------- out-of-bounds canaries (compute-sanitizer is closed on this pool)
@pytest.mark.parametrize("shape", [(2, 40, 130, 3), (1, 19, 33, 3), (1, 16, 64, 3), (3, 17, 2, 3)])
def test_no_writes_outside_the_output(shdr_gpu, shape):
    """Every kernel writes its output and nothing else: guard bands before and after the tensor stay intact."""
    from shdr import _native as N
    n, h, w, _ = shape
    img = rnd(shape, 123 + w)
    d_img = shdr_gpu.DeviceArray.from_numpy(img)
    wts = np.random.default_rng(5).normal(0, 0.5, (n, 11)).astype(np.float32)
    d_w = shdr_gpu.DeviceArray.from_numpy(wts)
    G = 4096                                  # guard floats on each side (16 KB, keeps 16-byte alignment)

    def run(nfloats, launch):
        buf = shdr_gpu.DeviceArray.from_numpy(np.full(nfloats + 2 * G, -7.0, np.float32))
        launch(buf.ptr + 4 * G)
        got = buf.numpy()
        assert (got[:G] == -7.0).all() and (got[-G:] == -7.0).all(), "kernel wrote outside its output tensor"
        assert not (got[G:-G] == -7.0).any(), "kernel left output elements unwritten"
        return got[G:-G]

    px = n * h * w
    fe = run(px * 93, lambda p: N.check(N.lib.shdr_frontend_f32(d_img.ptr, p, n, h, w, 0, None)))
    assert np.array_equal(fe.reshape(n, h, w, 93), oracle.frontend(img))
    hp = run(px * 84, lambda p: N.check(N.lib.shdr_hist_multi_f32(d_img.ptr, p, n, h, w, 16, None)))
    assert_rel(hp.reshape(n, h, w, 84), oracle.hist_multi(img, pool_k=16), RTOL_POOL)
    fp = run(px * 93, lambda p: N.check(N.lib.shdr_frontend_f32(d_img.ptr, p, n, h, w, 16, None)))
    assert_rel(fp.reshape(n, h, w, 93)[..., 9:], oracle.hist_multi(img, pool_k=16), RTOL_POOL)
    curve = shdr_gpu.DeviceArray.empty((n, 1024))
    run(px * 3, lambda p: N.check(N.lib.shdr_linearize_f32(d_img.ptr, d_w.ptr, p, curve.ptr, n, h * w * 3, None)))


def _to_bf16_bits(a):
    """float32 -> bfloat16 bit patterns, round to nearest even (what __float2bfloat16_rn does for finite values)"""
    u = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)


@pytest.mark.parametrize("shape", [(2, 24, 40, 3), (1, 5, 1027, 3), (3, 17, 31, 3)])
def test_frontend_bf16_is_the_rounded_fp32_front_end(shdr_gpu, shape):
    """SURVEY 8(f) rank 2 (reduced-precision output flag): exactly the fp32 front end rounded to bfloat16"""
    img = rnd(shape, sum(shape) + 5)
    got = shdr_gpu.frontend_bf16(shdr_gpu.DeviceArray.from_numpy(img))
    assert got.itemsize == 2 and got.shape == shape[:3] + (93,)
    assert np.array_equal(got.numpy(), _to_bf16_bits(oracle.frontend(img)))


@pytest.mark.parametrize("shape", [(2, 24, 40, 3), (1, 5, 1027, 3)])
def test_frontend_f16_is_the_rounded_fp32_front_end(shdr_gpu, shape):
    """the same flag with IEEE half precision: exactly float16(fp32 front end)"""
    img = rnd(shape, sum(shape) + 6)
    got = shdr_gpu.frontend_f16(shdr_gpu.DeviceArray.from_numpy(img))
    assert got.itemsize == 2 and got.shape == shape[:3] + (93,)
    out = got.numpy()
    assert out.dtype == np.float16 and np.array_equal(out, oracle.frontend(img).astype(np.float16))


@pytest.mark.parametrize("value", [0.0, 1.0, 0.5, 0.125, 0.375, 0.0625, 1.0 / 32.0, 31.0 / 32.0, 0.3])
def test_pooled_constant_images(shdr_gpu, value):
    """Constant images: at a bin centre every vote of that bin is exactly 1.0, so every interior 16x16 window sums to
    256 * 2^24 = 2^32 in the integer pipeline -- the overflow guard of pooled_slide.cu (one capped column per window)
    must give exactly 1.0 (within fp32 rounding), and the other bins must be exactly 0."""
    img = np.full((2, 40, 136, 3), value, np.float32)
    got = shdr_gpu.hist_multi(shdr_gpu.DeviceArray.from_numpy(img), pool=True).numpy()
    ref = oracle.hist_multi(img, pool_k=16)
    assert_rel(got, ref, RTOL_POOL)
    assert np.array_equal(got == 0, ref == 0)
    full = shdr_gpu.frontend(shdr_gpu.DeviceArray.from_numpy(img), pool=True).numpy()
    assert_rel(full[..., 9:], ref, RTOL_POOL)
    assert np.array_equal(full[..., :3], img) and np.all(full[..., 3:9] == 0)


def test_pooled_sparse_votes_stay_relative(shdr_gpu):
    """One grazing pixel in an otherwise empty bin: the pooled value is ~1e-9, and still within 1e-5 RELATIVE (the
    integer window sums are exact; an fp32 sliding sum would lose it)."""
    img = np.full((1, 48, 80, 3), 0.9, np.float32)
    img[0, 20, 40] = [0.375 - 2.0 ** -20, 0.625 + 2.0 ** -21, 0.125 - 2.0 ** -22]      # votes of 2^-18 .. 2^-20 for B = 4
    got = shdr_gpu.hist_multi(shdr_gpu.DeviceArray.from_numpy(img), pool=True).numpy()
    ref = oracle.hist_multi(img, pool_k=16)
    t64 = oracle.hist_multi(img, pool_k=16, dtype=np.float64)
    assert_rel(got, ref, RTOL_POOL)
    assert_rel(got, t64, RTOL_POOL)
    assert ((ref > 0) & (ref < 1e-6)).any()               # the case really contains tiny non-zero pooled values


def test_pooled_is_bit_deterministic(shdr_gpu):
    """The integer pipeline has no floating-point reduction whose order could vary: repeated launches (task order,
    warp timing and strip assignment differ) must give bit-identical tensors, for both variants."""
    img = rnd((5, 150, 200, 3), 77)
    d = shdr_gpu.DeviceArray.from_numpy(img)
    a84, a93 = shdr_gpu.hist_multi(d, pool=True).numpy(), shdr_gpu.frontend(d, pool=True).numpy()
    for _ in range(5):
        assert np.array_equal(shdr_gpu.hist_multi(d, pool=True).numpy(), a84)
        assert np.array_equal(shdr_gpu.frontend(d, pool=True).numpy(), a93)
    assert np.array_equal(a93[..., 9:], a84)
