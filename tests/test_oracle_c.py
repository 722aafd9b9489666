"""The two independent restatements of the reference path -- NumPy (oracle/np_oracle.py) and plain C
(oracle/shdr_oracle.c) -- agree: bit for bit where the operation order is the same (Sobel, histogram, pool, apply_rf),
within a few ulp where a reduction order differs (PCA dot product, _increase sum)."""
import subprocess
import os

import numpy as np
import pytest

import oracle
from oracle import c_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def built():
    if not c_oracle.available():
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s"], check=True)
    assert c_oracle.available() and c_oracle.threads() >= 1


def rnd(shape, seed=0):
    return np.random.default_rng(seed).random(shape, dtype=np.float32)


def test_sobel_hist_bit_exact():
    img = rnd((2, 19, 23, 3), 1)
    img[0, 0, :4, 0] = [0.0, 1.0, 0.5, 0.125]
    assert np.array_equal(c_oracle.sobel_edges6(img), oracle.sobel_edges6(img))
    for B in (1, 3, 4, 5, 8, 16, 33):
        assert np.array_equal(c_oracle.histogram_layer(img, B), oracle.histogram_layer(img, B)), B
    assert np.array_equal(c_oracle.frontend(img), oracle.frontend(img))


@pytest.mark.parametrize("shape", [(1, 20, 23, 3), (2, 3, 5, 3), (1, 40, 17, 3)])
def test_pool_bit_exact(shape):
    img = rnd(shape, 2)
    h = oracle.histogram_layer(img, 8)
    assert np.array_equal(c_oracle.avg_pool_same(h, 16), oracle.avg_pool_same(h, 16))
    assert np.array_equal(c_oracle.hist_multi(img, pool_k=16), oracle.hist_multi(img, pool_k=16))


def test_inverse_crf(emor):
    _, g0, hinv = emor
    w = np.random.default_rng(3).normal(0, 0.5, (4, 11)).astype(np.float32)
    pca_c, pca_n = c_oracle.invcrf_pca_w_2_invcrf(w, g0, hinv), oracle.invcrf_pca_w_2_invcrf(w, g0, hinv)
    assert np.abs(pca_c - pca_n).max() <= 5e-7
    cur_c, cur_n = c_oracle.increase(pca_n), oracle.increase(pca_n)
    assert np.abs(cur_c - cur_n).max() <= 2e-6
    x = rnd((4, 9, 11, 3), 4) * 1.2 - 0.1
    x[0, 0, 0] = [0.0, 1.0, -3.0]
    assert np.array_equal(c_oracle.apply_rf(x, cur_n), oracle.apply_rf(x, cur_n))
    y, _ = c_oracle.linearize(x, w, g0, hinv)
    assert np.abs(y - oracle.linearize(x, w, g0, hinv)[0]).max() <= 3e-6


def test_kat_lin2(kat_lin2):
    v = kat_lin2["values"].reshape(1, 1, 5, 1)
    np.testing.assert_allclose(c_oracle.histogram_layer(v, 5)[0, 0], kat_lin2["votes"], atol=2e-7)
