"""The gradient oracle (oracle/np_oracle.py section D) against two independent references that need no GPU:
``torch.autograd`` on a float64 torch restatement of the forward ops, and central finite differences.  The oracle
restates what TensorFlow's autodiff does for the reference's op sequences; torch's autodiff rules for the same ops
(floor: no gradient, abs -> sign, where -> taken branch, min -> evenly among ties... [amin], index gather ->
scatter-add) are the published ones TF shares."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle

torch.set_num_threads(2)


def t64(a, grad=False):
    return torch.tensor(np.asarray(a, dtype=np.float64), requires_grad=grad)


def torch_apply_rf(x, rf):
    b, k = rf.shape
    y = (k - 1) * x.reshape(b, -1)
    y0 = torch.floor(y).detach()
    y1 = y0 + 1
    i0 = y0.long().clamp(0, k - 1)
    i1 = y1.long().clamp(0, k - 1)
    v0 = torch.gather(rf, 1, i0)
    v1 = torch.gather(rf, 1, i1)
    return ((y1 - y) * v0 + (y - y0) * v1).reshape(x.shape)


def torch_increase(rf):
    g = rf[:, 1:] - rf[:, :-1]
    m = torch.amin(g, dim=-1, keepdim=True)          # amin spreads the gradient evenly among ties, like tf.reduce_min
    u = g + torch.relu(-m)
    n = u / u.sum(-1, keepdim=True)
    return F.pad(torch.cumsum(n, -1), (1, 0))


def torch_hist(img, B):
    outs = []
    for i in range(1, B + 1):
        d = torch.abs(img - (2.0 * i - 1.0) / (2.0 * B))
        outs.append(torch.where(d < 1.0 / B, 1.0 - d * B, torch.zeros_like(d)))
    return torch.cat(outs, -1)


def torch_sobel6(img):
    n, h, w, c = img.shape
    x = F.pad(img.permute(0, 3, 1, 2), (1, 1, 1, 1), mode="reflect")
    ky = torch.tensor([[-1., -2., -1.], [0., 0., 0.], [1., 2., 1.]], dtype=img.dtype)
    k = torch.stack([ky, ky.t()])[:, None]                         # [2,1,3,3]: dy, dx
    e = F.conv2d(x, k.repeat(c, 1, 1, 1), groups=c)                # channel c*2 + k
    return e.permute(0, 2, 3, 1)


def torch_frontend(img):
    return torch.cat([img, torch_sobel6(img), torch_hist(img, 4), torch_hist(img, 8), torch_hist(img, 16)], -1)


def test_apply_rf_grad_matches_autograd():
    rng = np.random.default_rng(0)
    x = rng.random((3, 7, 5, 3)) * 1.2 - 0.1                       # some values outside [0,1]: clipped indices
    rf = np.sort(rng.random((3, 64)), axis=1)
    gy = rng.normal(size=x.shape)
    tx, trf = t64(x, True), t64(rf, True)
    torch_apply_rf(tx, trf).backward(t64(gy))
    gx, grf = oracle.apply_rf_grad(x, rf, gy, index_dtype=np.float64)
    np.testing.assert_allclose(gx, tx.grad.numpy(), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(grf, trf.grad.numpy(), rtol=1e-12, atol=1e-12)


def test_increase_grad_matches_autograd_and_fd(emor):
    _, g0, hinv = emor
    rng = np.random.default_rng(1)
    w = rng.normal(0, 0.5, (4, 11))
    rf = oracle.invcrf_pca_w_2_invcrf(w, g0, hinv, np.float64)      # non-monotone curves: the relu(-min) path is live
    rf[3] = np.linspace(0, 1, 1024) ** 2                           # a monotone one: min > 0, no shift
    gout = rng.normal(size=rf.shape)
    trf = t64(rf, True)
    torch_increase(trf).backward(t64(gout))
    got = oracle.increase_grad(rf, gout)
    np.testing.assert_allclose(got, trf.grad.numpy(), rtol=1e-9, atol=1e-9)
    # central differences on a few coordinates
    f = lambda r: float((oracle.increase(r, np.float64) * gout).sum())
    for (b, j) in [(0, 5), (1, 512), (2, 1023), (3, 100)]:
        e = np.zeros_like(rf); e[b, j] = 1e-6
        fd = (f(rf + e) - f(rf - e)) / 2e-6
        assert abs(fd - got[b, j]) <= 1e-4 * max(1.0, abs(fd))


def test_increase_grad_ties():
    rf = np.array([[0.0, 0.1, 0.05, 0.3, 0.25, 0.6]])              # two equal minima (-0.05)
    gout = np.arange(6, dtype=np.float64)[None] / 3.0
    trf = t64(rf, True)
    torch_increase(trf).backward(t64(gout))
    np.testing.assert_allclose(oracle.increase_grad(rf, gout), trf.grad.numpy(), rtol=1e-12, atol=1e-12)


def test_pca_grad(emor):
    _, g0, hinv = emor
    g = np.random.default_rng(2).normal(size=(3, 1024))
    tw = t64(np.zeros((3, 11)), True)
    (t64(g0)[None] + tw @ t64(hinv).t()).backward(t64(g))
    np.testing.assert_allclose(oracle.invcrf_pca_grad(g, hinv), tw.grad.numpy(), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("shape", [(2, 9, 11, 3), (1, 2, 2, 3), (1, 3, 5, 3)])
def test_frontend_grad_matches_autograd(shape):
    rng = np.random.default_rng(3)
    img = rng.random(shape, dtype=np.float32)
    img[0, 0, 0] = [0.125, 0.5, 1.0]                               # a bin centre (sign 0), interior, upper edge
    gfeat = rng.normal(size=shape[:3] + (93,))
    ti = t64(img, True)
    torch_frontend(ti).backward(t64(gfeat))
    np.testing.assert_allclose(oracle.frontend_grad(img, gfeat), ti.grad.numpy(), rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("B", [4, 5, 16])
def test_histogram_layer_grad(B):
    rng = np.random.default_rng(B)
    img = rng.random((1, 6, 7, 3), dtype=np.float32)
    gh = rng.normal(size=(1, 6, 7, 3 * B))
    ti = t64(img, True)
    torch_hist(ti, B).backward(t64(gh))
    np.testing.assert_allclose(oracle.histogram_layer_grad(img, gh, B), ti.grad.numpy(), rtol=1e-10, atol=1e-10)


def test_linearize_ex_matches_the_chain():
    rng = np.random.default_rng(5)
    x = (rng.random((2, 5, 6, 3)) * 1.4 - 0.2).astype(np.float32)
    rf = np.sort(rng.random((2, 1024)), axis=1).astype(np.float32) * 1.3
    c, y, a = oracle.linearize_ex(x, rf, 0.12)
    assert c.min() >= 0 and c.max() <= 1 and np.array_equal(y, oracle.apply_rf(c, rf))
    m = y.max(axis=3)
    want = np.minimum(1, np.maximum(0, m - 1 + 0.12) / 0.12)
    np.testing.assert_allclose(a[..., 0], want, atol=1e-6)
    assert np.array_equal(a[..., 0], a[..., 1]) and np.array_equal(a[..., 0], a[..., 2])
