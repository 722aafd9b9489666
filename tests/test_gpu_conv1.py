"""GPU parity of the front end fused into crfFeatureNet.conv1 (tcgen05 tensor cores).

Reference: tf.concat([img, edge6, hist4, hist8, hist16]) -> Conv2D(64, (7,7), strides 2, 'SAME', bias)
(linearization_net.py:312-322, :91, :107).  The kernel multiplies fp16 operands and accumulates in fp32, so two gates:

    TOL_EXACT   vs the oracle convolution of the SAME fp16-rounded features and weights (fp64 sum): only the
                fp32 summation order differs -> 2e-5 of the output's largest magnitude.  This is the parity gate: a
                wrong tap, channel, padding or tile border is an O(1) error.
    TOL_FP32    vs the fp32 reference convolution (what TensorFlow computes on a CPU): the fp16 rounding of
                features and weights, <= 2^-11 relative per operand -> 1e-3 of the output's largest magnitude (measured 2.7e-4).
"""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

TOL_EXACT = 2e-5
TOL_FP32 = 1e-3


def _case(shape, seed, quant=False):
    rng = np.random.default_rng(seed)
    img = rng.random(shape, dtype=np.float32)
    if quant:
        img = np.round(img * 255).astype(np.float32) / np.float32(255)
    kern = (rng.normal(0, 1, (7, 7, 93, 64)) / np.sqrt(7 * 7 * 93)).astype(np.float32)   # Glorot-like scale
    bias = rng.normal(0, 0.1, 64).astype(np.float32)
    return img, kern, bias


def _run(shdr, img, kern, bias=None, scale=None, relu=False):
    D = shdr.DeviceArray.from_numpy
    packed = shdr.conv1_pack_weights(D(kern))
    return shdr.frontend_conv1(D(img), packed, bias=None if bias is None else D(bias),
                               scale=None if scale is None else D(scale), relu=relu).numpy()


@pytest.mark.parametrize("shape", [(1, 32, 16, 3), (2, 64, 48, 3), (1, 33, 17, 3), (3, 2, 2, 3), (1, 7, 100, 3),
                                   (1, 70, 6, 3), (2, 45, 51, 3), (2, 101, 77, 3), (1, 64, 258, 3), (5, 18, 34, 3),
                                   (1, 3, 3, 3), (1, 31, 130, 3)])
def test_frontend_conv1_matches_oracle(shdr_gpu, shape):
    img, kern, bias = _case(shape, sum(shape))
    got = _run(shdr_gpu, img, kern, bias)
    ref_b = oracle.frontend_conv1(img, kern, bias, half_operands=True)
    ref_f = oracle.frontend_conv1(img, kern, bias)
    assert got.shape == ref_f.shape == (shape[0], (shape[1] + 1) // 2, (shape[2] + 1) // 2, 64)
    scale = np.abs(ref_f).max()
    assert np.abs(got - ref_b).max() <= TOL_EXACT * scale, np.abs(got - ref_b).max() / scale
    assert np.abs(got - ref_f).max() <= TOL_FP32 * scale, np.abs(got - ref_f).max() / scale


def test_frontend_conv1_tap_and_channel_map(shdr_gpu):
    """One-hot kernels: every (tap, channel) pair lands on the right feature of the right neighbour (49 x 93 checks
    folded into 64-channel batches would be slow; probe the corners of the tap window and every channel once)."""
    rng = np.random.default_rng(3)
    img = rng.random((1, 40, 24, 3), dtype=np.float32)
    feat = oracle.half_round(oracle.frontend(img))
    taps = [(0, 0), (0, 6), (6, 0), (6, 6), (3, 3), (2, 5), (5, 2)]
    for t0 in range(0, 93, 64):
        kern = np.zeros((7, 7, 93, 64), np.float32)
        picks = []
        for o in range(64):
            ch = t0 + o
            if ch >= 93:
                break
            ky, kx = taps[ch % len(taps)]
            kern[ky, kx, ch, o] = 1.0
            picks.append((ky, kx, ch, o))
        got = _run(shdr_gpu, img, kern)
        ref = oracle.conv2d_same_s2(feat, kern)
        for ky, kx, ch, o in picks:
            assert np.array_equal(got[..., o], ref[..., o].astype(np.float32)), (ky, kx, ch)


def test_frontend_conv1_scale_shift_relu(shdr_gpu):
    """Folded inference-mode batch norm + ReLU behind the convolution (linearization_net.py:108-109)."""
    img, kern, bias = _case((2, 48, 40, 3), 11, quant=True)
    rng = np.random.default_rng(12)
    scale = rng.uniform(0.5, 2.0, 64).astype(np.float32)
    got = _run(shdr_gpu, img, kern, bias, scale, relu=True)
    conv = oracle.frontend_conv1(img, kern, None, half_operands=True)
    ref = np.maximum(conv * scale + bias, 0.0)
    assert np.abs(got - ref).max() <= TOL_EXACT * np.abs(conv).max() * 2.0
    assert (got >= 0).all() and (got == 0).any()


def test_frontend_conv1_many_tiles_and_repeat(shdr_gpu):
    """More tiles than SMs (persistent loop, both accumulators, ring wrap-around) and run-to-run identity."""
    img, kern, bias = _case((3, 256, 208, 3), 21)
    D = shdr_gpu.DeviceArray.from_numpy
    packed = shdr_gpu.conv1_pack_weights(D(kern))
    d_img, d_b = D(img), D(bias)
    a = shdr_gpu.frontend_conv1(d_img, packed, bias=d_b).numpy()
    b = shdr_gpu.frontend_conv1(d_img, packed, bias=d_b).numpy()
    assert np.array_equal(a, b)
    ref = oracle.frontend_conv1(img, kern, bias, half_operands=True)
    assert np.abs(a - ref).max() <= TOL_EXACT * np.abs(ref).max()


def test_frontend_conv1_pair_odd_tile_count(shdr_gpu):
    """CTA-pair kernel with an ODD number of tiles spread over several iterations: the last pair runs a dummy tile in
    its second CTA (481 tiles = 13 x 37 on 74 clusters), and partial tiles on the right and bottom edges."""
    img, kern, bias = _case((1, 400, 584, 3), 31)
    got = _run(shdr_gpu, img, kern, bias)
    ref = oracle.frontend_conv1(img, kern, bias, half_operands=True)
    assert got.shape == ref.shape == (1, 200, 292, 64)
    assert np.abs(got - ref).max() <= TOL_EXACT * np.abs(ref).max()


def test_frontend_conv1_host_pipeline(shdr_gpu):
    """numpy in / numpy out through the pipelined host API (several chunks), with scale + bias + ReLU."""
    img, kern, bias = _case((5, 96, 80, 3), 41)
    scale = np.random.default_rng(42).uniform(0.5, 2.0, 64).astype(np.float32)
    got = shdr_gpu.frontend_conv1_host(img, kern, bias=bias, scale=scale, relu=True)
    conv = oracle.frontend_conv1(img, kern, None, half_operands=True)
    ref = np.maximum(conv * scale + bias, 0.0)
    assert got.shape == (5, 48, 40, 64)
    assert np.abs(got - ref).max() <= TOL_EXACT * np.abs(conv).max() * 2.0


def test_frontend_conv1_rejects_bad_arguments(shdr_gpu):
    D = shdr_gpu.DeviceArray.from_numpy
    with pytest.raises(ValueError):
        shdr_gpu.conv1_pack_weights(D(np.zeros((7, 7, 92, 64), np.float32)))
    packed = shdr_gpu.conv1_pack_weights(D(np.zeros((7, 7, 93, 64), np.float32)))
    with pytest.raises(ValueError):
        shdr_gpu.frontend_conv1(D(np.zeros((1, 8, 8, 4), np.float32)), packed)
    with pytest.raises(shdr_gpu.ShdrError):
        shdr_gpu.frontend_conv1(D(np.zeros((1, 1, 8, 3), np.float32)), packed)      # REFLECT needs h >= 2
