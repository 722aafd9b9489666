"""CPU tests of the oracle: the reference's known answers, structural properties, an fp64
evaluation of the same formulas, stock CPU PyTorch ops with the same published semantics, and
the committed golden vectors (regression anchor).  PARITY UNPINNED -- see oracle/__init__.py."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F
from hypothesis import given, settings, strategies as st

import oracle


def rnd(shape, seed=0):
    return np.random.default_rng(seed).random(shape, dtype=np.float32)


# ---------------------------------------------------------------- known answers of the reference
def test_kat_lin2_figure(kat_lin2):
    """figure/lin2.png: B=5 votes for .63 .65 .32 .84 .15."""
    v = kat_lin2["values"].reshape(1, 1, 5, 1)
    h = oracle.histogram_layer(v, 5)[0, 0]            # [value, bin]
    np.testing.assert_allclose(h, kat_lin2["votes"], atol=2e-7)


def test_emor_table_structure(emor):
    b, g0, hinv = emor
    assert g0[0] == 0 and g0[-1] == 1 and np.all(np.diff(g0) > 0)
    assert np.all(hinv[0] == 0) and np.all(hinv[-1] == 0)
    np.testing.assert_allclose(np.linalg.norm(hinv.astype(np.float64), axis=0), 1.0, atol=1e-6)
    np.testing.assert_allclose(b, np.linspace(0, 1, 1024), atol=1e-7)


def test_w_zero_gives_g0(emor):
    _, g0, hinv = emor
    c = oracle.invcrf_pca_w_2_invcrf(np.zeros((2, 11), np.float32), g0, hinv)
    assert np.array_equal(c, np.stack([g0, g0]))
    np.testing.assert_allclose(oracle.increase(c), c, atol=2e-6)


# ---------------------------------------------------------------- histogram properties
@pytest.mark.parametrize("B", [1, 3, 4, 5, 8, 16])
def test_hist_properties(B):
    img = rnd((1, 9, 7, 3), B)
    h = oracle.histogram_layer(img, B).reshape(1, 9, 7, B, 3)
    assert h.min() >= 0 and h.max() <= 1
    assert (np.count_nonzero(h, axis=3) <= 2).all()
    inner = (img >= 1 / (2 * B)) & (img <= 1 - 1 / (2 * B))
    np.testing.assert_allclose(h.sum(3)[inner], 1.0, atol=4e-7)      # partition of unity


def test_hist_range_ends():
    for B in (4, 8, 16):
        h = oracle.histogram_layer(np.float32([[[[0.0], [1.0]]]]), B)[0, 0]
        assert h[0, 0] == 0.5 and h[0, 1:].sum() == 0
        assert h[1, B - 1] == 0.5 and h[1, :B - 1].sum() == 0


def test_hist_channel_order():
    img = np.float32([0.125, 0.375, 0.875]).reshape(1, 1, 1, 3)     # bin 1, 2, 4 centres for B=4
    h = oracle.histogram_layer(img, 4)[0, 0, 0]
    want = np.zeros(12, np.float32)
    want[0 * 3 + 0] = want[1 * 3 + 1] = want[3 * 3 + 2] = 1.0       # channel = (bin-1)*3 + c
    assert np.array_equal(h, want)


@settings(max_examples=60, deadline=None)
@given(st.floats(-0.5, 1.5, width=32), st.integers(1, 33))
def test_hist_scalar_formula(v, B):
    h = oracle.histogram_layer(np.float32(v).reshape(1, 1, 1, 1), B).ravel()
    for i in range(1, B + 1):
        c = np.float32(np.float32(2 * i - 1) / np.float32(2 * B))
        d = np.abs(np.float32(v) - c)
        want = np.float32(1) - d * np.float32(B) if d < np.float32(1.0 / B) else np.float32(0)
        assert h[i - 1] == want


def test_hist_fp64_truth():
    img = rnd((2, 16, 16, 3), 3)
    for B in (4, 5, 8, 16):
        a = oracle.histogram_layer(img, B)
        t = oracle.histogram_layer(img, B, np.float64)
        assert np.abs(a - t).max() <= 1e-6


# ---------------------------------------------------------------- Sobel
def test_sobel_constant_and_reflect():
    assert not oracle.sobel_edges6(np.full((1, 5, 6, 3), 0.37, np.float32)).any()
    e = oracle.sobel_edges6(rnd((2, 8, 9, 3), 1)).reshape(2, 8, 9, 3, 2)
    assert not e[:, :, 0, :, 1].any() and not e[:, :, -1, :, 1].any()    # dx = 0 on left/right columns
    # dy on top/bottom rows cancels only up to the rounding of the sequential tap sum
    assert np.abs(e[:, 0, :, :, 0]).max() < 5e-7 and np.abs(e[:, -1, :, :, 0]).max() < 5e-7


def test_sobel_vs_torch_conv():
    img = rnd((2, 11, 13, 3), 2)
    e = oracle.sobel_edges6(img)
    t = torch.from_numpy(img).permute(0, 3, 1, 2)
    p = F.pad(t, (1, 1, 1, 1), mode="reflect")
    ky = torch.tensor([[-1., -2, -1], [0, 0, 0], [1, 2, 1]])
    kx = ky.t().contiguous()
    wgt = torch.stack([ky, kx] * 3)[:, None]                  # out channel = c*2 + k
    ref = F.conv2d(p, wgt, groups=3).permute(0, 2, 3, 1).numpy()
    np.testing.assert_allclose(e, ref, atol=1e-6)
    t64 = oracle.sobel_edges6(img, np.float64)
    assert np.abs(e - t64).max() <= 1e-6


def test_sobel_ramp():
    x = np.arange(6, dtype=np.float32)[None, None, :, None] * np.ones((1, 5, 1, 3), np.float32)
    e = oracle.sobel_edges6(x).reshape(1, 5, 6, 3, 2)
    assert np.all(e[0, :, 1:-1, :, 1] == 8) and not e[..., 0].any()      # dx of a unit ramp = 8


# ---------------------------------------------------------------- pool
def test_pool_counts_and_torch():
    x = rnd((1, 20, 23, 5), 4)
    p = oracle.avg_pool_same(x, 16)
    ones = oracle.avg_pool_same(np.ones((1, 20, 23, 1), np.float32), 16)
    np.testing.assert_allclose(ones, 1.0, atol=1e-6)
    t = torch.from_numpy(x).permute(0, 3, 1, 2)
    ref = F.avg_pool2d(F.pad(t, (7, 8, 7, 8)), 16, 1, count_include_pad=True)
    # torch has no asymmetric 'same': emulate, then re-normalise by the in-bounds count
    cnt = F.avg_pool2d(F.pad(torch.ones_like(t), (7, 8, 7, 8)), 16, 1)
    ref = (ref / cnt).permute(0, 2, 3, 1).numpy()
    np.testing.assert_allclose(p, ref, rtol=2e-6, atol=1e-7)
    # corner windows: 8x8 rows/cols at top-left -> 64? no: rows [0,8], cols [0,8] = 9x9 = 81
    assert np.isclose(p[0, 0, 0, 0], x[0, :9, :9, 0].astype(np.float64).mean(), rtol=1e-6)
    assert np.isclose(p[0, -1, -1, 0], x[0, -8:, -8:, 0].astype(np.float64).mean(), rtol=1e-6)


def test_pool_small_image():
    x = rnd((1, 3, 5, 2), 5)
    p = oracle.avg_pool_same(x, 16)
    np.testing.assert_allclose(p, np.broadcast_to(x.mean((1, 2), keepdims=True), x.shape), rtol=1e-6)


def test_frontend_layout(golden_small):
    img = golden_small["img"]
    f = oracle.frontend(img)
    assert f.shape[-1] == 93
    assert np.array_equal(f[..., :3], img)
    assert np.array_equal(f[..., 3:9], oracle.sobel_edges6(img))
    assert np.array_equal(f[..., 9:21], oracle.histogram_layer(img, 4))
    assert np.array_equal(f[..., 21:45], oracle.histogram_layer(img, 8))
    assert np.array_equal(f[..., 45:93], oracle.histogram_layer(img, 16))


# ---------------------------------------------------------------- inverse CRF
def test_increase_properties(emor):
    _, g0, hinv = emor
    w = np.random.default_rng(3).normal(0, 0.5, (16, 11)).astype(np.float32)
    pca = oracle.invcrf_pca_w_2_invcrf(w, g0, hinv)
    assert (np.diff(pca, axis=1).min(1) < 0).all()          # every curve needs the enforcement
    c = oracle.increase(pca)
    assert c.shape == (16, 1024) and (c[:, 0] == 0).all()
    assert np.all(np.diff(c, axis=1) >= 0)
    assert np.abs(c[:, -1] - 1).max() < 2e-6
    t = oracle.increase(oracle.invcrf_pca_w_2_invcrf(w, g0, hinv, np.float64), np.float64)
    assert np.abs(c - t).max() < 5e-6
    tc = torch.cumsum(torch.from_numpy(np.diff(c, axis=1)), 1).numpy()
    np.testing.assert_allclose(c[:, 1:], tc, atol=2e-6)


def test_increase_constant_curve_is_nan():
    """0/0 is not guarded by the reference (:376 is commented out)."""
    assert np.isnan(oracle.increase(np.full((1, 8), 0.3, np.float32))[0, 1:]).all()


def test_apply_rf_identity_and_edges():
    k = 1024
    rf = np.linspace(0, 1, k, dtype=np.float32)[None]
    x = rnd((1, 50, 3), 6)
    np.testing.assert_allclose(oracle.apply_rf(x, rf), x, atol=1.2e-7 * 2)
    rf2 = np.sort(rnd((1, k), 7))
    xs = np.float32([[0.0, 1.0, -0.3, 1.7, 0.5]])
    y = oracle.apply_rf(xs, rf2)[0]
    assert y[0] == rf2[0, 0] and y[1] == rf2[0, -1]
    assert np.isclose(y[2], rf2[0, 0], atol=1e-6) and np.isclose(y[3], rf2[0, -1], atol=1e-6)


def test_apply_rf_vs_numpy_interp():
    rf = np.sort(rnd((3, 64), 8), axis=1)
    x = rnd((3, 7, 5, 3), 9)
    y = oracle.apply_rf(x, rf)
    grid = np.linspace(0, 1, 64)
    for b in range(3):
        ref = np.interp(x[b].astype(np.float64), grid, rf[b].astype(np.float64))
        np.testing.assert_allclose(y[b], ref, atol=2e-6)


def test_parse_roundtrip(tmp_path, emor):
    b, g0, hinv = emor
    cols = [("B =", b), ("g0 =", g0)] + [(f"hinv({i + 1})=", hinv[:, i]) for i in range(11)]
    cols.append(("hinv(12)=", np.zeros(1024, np.float32)))
    lines = []
    for tag, v in cols:
        lines.append(tag + " ")
        for r in v.reshape(256, 4):
            lines.append("   ".join(f"{float(t):.9e}" for t in r))
    p = tmp_path / "invemor.txt"
    p.write_text("\n".join(lines) + "\n")
    b2, g2, h2 = oracle.parse_invemor(str(p))
    assert np.array_equal(b2, b) and np.array_equal(g2, g0) and np.array_equal(h2, hinv)


# ---------------------------------------------------------------- golden regression
def test_golden_regression(golden_small, emor):
    g = golden_small
    _, g0, hinv = emor
    assert np.array_equal(oracle.sobel_edges6(g["img"]), g["edges"])
    assert np.array_equal(oracle.histogram_layer(g["img"], 4), g["hist4"])
    assert np.array_equal(oracle.histogram_layer(g["img"], 5), g["hist5"])
    assert np.array_equal(oracle.frontend(g["img"]), g["frontend"])
    np.testing.assert_allclose(oracle.frontend(g["img"], pool_k=16), g["frontend_pooled"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(oracle.invcrf_pca_w_2_invcrf(g["w"], g0, hinv), g["pca"], atol=1e-6)
    np.testing.assert_allclose(oracle.increase(g["pca"]), g["curve"], atol=2e-6)
    np.testing.assert_allclose(oracle.apply_rf(g["x"], g["curve"]), g["lin"], atol=1e-6)


# ---------------------------------------------------------------- conv1 restatement
@pytest.mark.parametrize("h,w", [(12, 16), (13, 9), (7, 7), (2, 2), (31, 18)])
def test_conv_same_s2_vs_torch(h, w):
    """tf.keras.layers.Conv2D(64, (7,7), strides 2, 'SAME') semantics (linearization_net.py:91): out = ceil(in/2), pad
    = max((out-1)*2 + 7 - in, 0) with the smaller half in front, cross-correlation -- against stock CPU PyTorch with
    the same explicit padding."""
    rng = np.random.default_rng(h * 100 + w)
    x, k, b = rng.normal(size=(2, h, w, 5)), rng.normal(size=(7, 7, 5, 4)), rng.normal(size=4)
    got = oracle.conv2d_same_s2(x, k, b)
    oh, ow = (h + 1) // 2, (w + 1) // 2
    ph, pw = max((oh - 1) * 2 + 7 - h, 0), max((ow - 1) * 2 + 7 - w, 0)
    xt = F.pad(torch.from_numpy(x).permute(0, 3, 1, 2), (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
    ref = F.conv2d(xt, torch.from_numpy(k).permute(3, 2, 0, 1), torch.from_numpy(b), stride=2).permute(0, 2, 3, 1).numpy()
    assert got.shape == (2, oh, ow, 4)
    assert np.abs(got - ref).max() < 1e-12


def test_bf16_round_vs_torch():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.normal(size=4096), rng.random(4096), [0.0, 1.0, 1.00390625, 1.01171875, -2.5e-3, 3.0e38]]).astype(np.float32)
    assert np.array_equal(oracle.bf16_round(x), torch.from_numpy(x).bfloat16().float().numpy())


def test_half_round_vs_torch():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.normal(size=4096), rng.random(4096) * 1e-6, [0.0, 1.0, 1.00048828125, 6.0e-8, -2.5e-3]]).astype(np.float32)
    assert np.array_equal(oracle.half_round(x), torch.from_numpy(x).half().float().numpy())
    assert np.array_equal(oracle.half_round(np.float32([1e6, -3e38])), np.float32([65504.0, -65504.0]))   # saturates


def test_frontend_conv1_composition():
    """frontend_conv1 == conv2d_same_s2(frontend(img)); the bf16-operand variant differs by the rounding only."""
    rng = np.random.default_rng(1)
    img = rng.random((1, 20, 14, 3), dtype=np.float32)
    kern = (rng.normal(size=(7, 7, 93, 64)) / 67.5).astype(np.float32)
    bias = rng.normal(0, 0.1, 64).astype(np.float32)
    full = oracle.frontend_conv1(img, kern, bias)
    assert np.array_equal(full, oracle.conv2d_same_s2(oracle.frontend(img), kern, bias))
    b16 = oracle.frontend_conv1(img, kern, bias, half_operands=True)
    assert full.shape == b16.shape == (1, 10, 7, 64)
    assert 0 < np.abs(full - b16).max() < 3e-3 * np.abs(full).max()
