"""tf_adapter.patch exercised end to end on the GPU with a FAKE `tensorflow` module.

TensorFlow cannot be installed in this project's image, so the adapter's eager path is driven through a minimal
stand-in that implements exactly the TF surface the adapter touches (`tf.experimental.dlpack.to_dlpack/from_dlpack`,
`tf.executing_eagerly`, `tf.py_function`, `tf.float32`) on top of torch CUDA tensors -- the DLPack capsule protocol
is the same one real TF eager tensors use.  The patched objects are stand-ins for the reference's
`linearization_net` / `tf_utils` modules with the same attribute names.  This proves the plumbing (borrow via
DLPack, run the native kernel, hand the result back via DLPack, same signatures), not TF itself.
"""
import sys
import types

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


SUMMARIES = []
PY_FUNCTION_CALLS = []


class FakeTFTensor:
    """Looks like a TF eager tensor to the adapter: no __dlpack__, module name starts with 'tensorflow'."""
    __module__ = "tensorflow.python.framework.ops"

    def __init__(self, t):
        self._t = t
        self.shape = tuple(t.shape)

    def numpy(self):
        return self._t.cpu().numpy()

    def __getitem__(self, idx):
        return FakeTFTensor(self._t[idx].contiguous())

    def set_shape(self, shape):
        assert tuple(shape) == self.shape, (tuple(shape), self.shape)      # the adapter's static-shape functions


def _install_fake_tf(eager=True):
    import torch
    tf = types.ModuleType("tensorflow")
    tf.float32 = "float32"
    tf.executing_eagerly = lambda: eager

    def py_function(fn, args, dtype):
        """graph mode: the adapter wraps every native call (forward and gradient) in tf.py_function and restores the
        static shape with set_shape; here the wrapped function simply runs, and the shapes set are recorded"""
        PY_FUNCTION_CALLS.append(dtype)
        out = fn(*args)
        assert isinstance(dtype, list) == isinstance(out, tuple)
        return list(out) if isinstance(out, tuple) else out
    tf.py_function = py_function
    tf.cast = lambda x, dtype: x
    tf.summary = types.SimpleNamespace(image=lambda name, t: SUMMARIES.append((name, t.shape)))

    def custom_gradient(f):
        """records the gradient function on the output, the way a GradientTape would hold it"""
        def wrapper(*xs):
            y, grad = f(*xs)
            y._grad_fn = grad
            return y
        return wrapper
    tf.custom_gradient = custom_gradient
    dl = types.SimpleNamespace(
        to_dlpack=lambda x: torch.utils.dlpack.to_dlpack(x._t),
        from_dlpack=lambda cap: FakeTFTensor(torch.utils.dlpack.from_dlpack(cap)))
    tf.experimental = types.SimpleNamespace(dlpack=dl)
    sys.modules["tensorflow"] = tf
    return tf


@pytest.mark.parametrize("eager", [True, False], ids=["eager", "graph"])
def test_patch_reference_modules(shdr_gpu, emor, tmp_path, monkeypatch, eager):
    import torch
    from shdr import tf_adapter
    _install_fake_tf(eager)
    PY_FUNCTION_CALLS.clear()
    try:
        _, g0, hinv = emor
        # the adapter parses 'invemor.txt' relative to the CWD, like the reference (linearization_net.py:219)
        cols = [("B =", np.linspace(0, 1, 1024, dtype=np.float32)), ("g0 =", g0)] + \
               [(f"hinv({i + 1})=", hinv[:, i]) for i in range(11)]
        lines = []
        for tag, v in cols:
            lines.append(tag + " ")
            lines += ["   ".join(f"{float(t):.9e}" for t in r) for r in v.reshape(256, 4)]
        (tmp_path / "invemor.txt").write_text("\n".join(lines) + "\n")
        monkeypatch.chdir(tmp_path)

        # stand-ins for the reference's modules (same attribute names as linearization_net.py / tf_utils.py)
        lin = types.ModuleType("linearization_net")

        class model:            # noqa: N801  (reference spelling)
            def call(self, img, training="training"):
                raise AssertionError("stock TF path must have been replaced")

            def histogram_layer(self, img, max_bin):
                raise AssertionError("stock TF path must have been replaced")

            @staticmethod
            def _increase(rf):
                raise AssertionError("stock TF path must have been replaced")

        class AEInvcrfDecodeNet:
            def invcrf_pca_w_2_invcrf(self, w):
                raise AssertionError("stock TF path must have been replaced")

        lin.model, lin.AEInvcrfDecodeNet = model, AEInvcrfDecodeNet
        tfu = types.ModuleType("tf_utils")
        tfu.apply_rf = lambda x, rf: (_ for _ in ()).throw(AssertionError("not replaced"))
        assert tf_adapter.patch(lin, tfu) is True

        rng = np.random.default_rng(11)
        img = rng.random((2, 24, 40, 3), dtype=np.float32)
        w = rng.normal(0, 0.5, (2, 11)).astype(np.float32)
        t_img = FakeTFTensor(torch.from_numpy(img).cuda())
        t_w = FakeTFTensor(torch.from_numpy(w).cuda())

        hist = lin.model().histogram_layer(t_img, 8)
        assert isinstance(hist, FakeTFTensor) and hist.shape == (2, 24, 40, 24)
        assert np.array_equal(hist.numpy(), oracle.histogram_layer(img, 8))

        feat = lin.shdr_frontend(t_img)
        assert np.array_equal(feat.numpy(), oracle.frontend(img))

        pca = lin.AEInvcrfDecodeNet().invcrf_pca_w_2_invcrf(t_w)
        ref_pca = oracle.invcrf_pca_w_2_invcrf(w, g0, hinv)
        assert np.abs(pca.numpy() - ref_pca).max() <= 1e-6
        curve = lin.model._increase(pca)
        ref_curve = oracle.increase(ref_pca)
        assert np.abs(curve.numpy() - ref_curve).max() <= 5e-6
        lin_img = tfu.apply_rf(t_img, curve)
        assert lin_img.shape == img.shape
        assert np.abs(lin_img.numpy() - oracle.apply_rf(img, ref_curve)).max() <= 1e-5

        # ---- gradients ride on tf.custom_gradient (the training steps differentiate through these ops)
        gy = rng.normal(size=img.shape).astype(np.float32)
        gx, grf = lin_img._grad_fn(FakeTFTensor(torch.from_numpy(gy).cuda()))
        rx, rrf = oracle.apply_rf_grad(img, curve.numpy(), gy)
        assert np.abs(gx.numpy() - rx).max() <= 2e-5 * np.abs(rx).max()
        assert np.abs(grf.numpy() - rrf).max() <= 1e-4 * np.abs(rrf).max()
        gc = rng.normal(size=(2, 1024)).astype(np.float32)
        g_inc = curve._grad_fn(FakeTFTensor(torch.from_numpy(gc).cuda()))
        want = oracle.increase_grad(pca.numpy(), gc)
        assert np.abs(g_inc.numpy() - want).max() <= 5e-5 * np.abs(want).max()
        g_w = pca._grad_fn(FakeTFTensor(torch.from_numpy(gc).cuda()))
        want = oracle.invcrf_pca_grad(gc, hinv)
        assert np.abs(g_w.numpy() - want).max() <= 5e-5 * np.abs(want).max()
        gf = rng.normal(size=(2, 24, 40, 93)).astype(np.float32)
        g_img = feat._grad_fn(FakeTFTensor(torch.from_numpy(gf).cuda()))
        want = oracle.frontend_grad(img, gf)
        assert np.abs(g_img.numpy() - want).max() <= 2e-5 * np.abs(want).max()
        gh = rng.normal(size=(2, 24, 40, 24)).astype(np.float32)
        g_img = hist._grad_fn(FakeTFTensor(torch.from_numpy(gh).cuda()))
        want = oracle.histogram_layer_grad(img, gh, 8)
        assert np.abs(g_img.numpy() - want).max() <= 2e-5 * np.abs(want).max()

        # ---- model.call itself is replaced: fused front end -> stock sub-networks -> native _increase
        proj = rng.normal(0, 0.05, (93, 11)).astype(np.float32)
        seen = {}

        def crf_feature_net(feat93, training):          # stand-in backbone: global mean of the 93 channels
            seen["feat"], seen["training"] = feat93, training
            return FakeTFTensor(feat93._t.mean(dim=(1, 2)))

        def ae_invcrf_decode_net(feature):              # stand-in Dense(11) + the (patched) PCA method
            w11 = FakeTFTensor(feature._t @ torch.from_numpy(proj).cuda())
            seen["w"] = w11
            return lin.AEInvcrfDecodeNet().invcrf_pca_w_2_invcrf(w11)

        net = lin.model()
        net.crf_feature_net, net.ae_invcrf_decode_net = crf_feature_net, ae_invcrf_decode_net
        SUMMARIES.clear()
        out = net.call(t_img, training=False)
        assert seen["training"] is False and seen["feat"].shape == (2, 24, 40, 93)
        assert np.array_equal(seen["feat"].numpy(), oracle.frontend(img))
        assert SUMMARIES == [("edge0", (2, 24, 40, 3)), ("edge1", (2, 24, 40, 3))]
        w_ref = oracle.frontend(img).mean(axis=(1, 2)) @ proj
        assert np.abs(seen["w"].numpy() - w_ref).max() <= 1e-5
        ref_out = oracle.increase(oracle.invcrf_pca_w_2_invcrf(seen["w"].numpy(), g0, hinv))
        assert out.shape == (2, 1024) and np.abs(out.numpy() - ref_out).max() <= 5e-6

        # graph mode goes through tf.py_function for every forward and gradient call; eager mode never does
        assert (len(PY_FUNCTION_CALLS) > 10) == (not eager)

        monkeypatch.setenv("SHDR_NATIVE", "0")          # A/B switch leaves the stock ops alone
        assert tf_adapter.patch(lin, tfu) is False
    finally:
        sys.modules.pop("tensorflow", None)


@pytest.mark.parametrize("eager", [True, False], ids=["eager", "graph"])
def test_patch_fuse_conv1(shdr_gpu, emor, tmp_path, monkeypatch, eager):
    """patch(..., fuse_conv1=True): inference calls run front end + conv1 + folded norm1 + act1 as ONE tensor-core
    kernel on the 3-channel image (linearization_net.py:312-322 -> :107-109); training calls keep the stock layers."""
    import torch
    from shdr import tf_adapter
    tf = _install_fake_tf(eager)
    tf.convert_to_tensor = lambda v: v
    tf.zeros = lambda shape, dtype: FakeTFTensor(torch.zeros(shape, device="cuda"))
    tf.reduce_mean = lambda x, axes, keepdims=False: FakeTFTensor(x._t.mean(dim=tuple(axes), keepdim=keepdims))
    PY_FUNCTION_CALLS.clear()
    try:
        _, g0, hinv = emor
        cols = [("B =", np.linspace(0, 1, 1024, dtype=np.float32)), ("g0 =", g0)] + \
               [(f"hinv({i + 1})=", hinv[:, i]) for i in range(11)]
        lines = []
        for tag, v in cols:
            lines.append(tag + " ")
            lines += ["   ".join(f"{float(t):.9e}" for t in r) for r in v.reshape(256, 4)]
        (tmp_path / "invemor.txt").write_text("\n".join(lines) + "\n")
        monkeypatch.chdir(tmp_path)

        rng = np.random.default_rng(5)
        kern = (rng.normal(0, 1, (7, 7, 93, 64)) / 67.5).astype(np.float32)
        vec = {k: v.astype(np.float32) for k, v in dict(
            bias=rng.normal(0, 0.1, 64), gamma=rng.uniform(0.5, 1.5, 64), beta=rng.normal(0, 0.1, 64),
            mean=rng.normal(0, 0.2, 64), var=rng.uniform(0.5, 2.0, 64)).items()}
        T = lambda a: FakeTFTensor(torch.from_numpy(np.ascontiguousarray(a)).cuda())   # noqa: E731
        seen = {}
        lin = types.ModuleType("linearization_net")

        class crfFeatureNet:     # noqa: N801  (reference spelling; attributes of linearization_net.py:88-99)
            def __init__(self):
                self.conv1 = types.SimpleNamespace(kernel=T(kern), bias=T(vec["bias"]), use_bias=True)
                self.norm1 = types.SimpleNamespace(gamma=T(vec["gamma"]), beta=T(vec["beta"]), moving_mean=T(vec["mean"]),
                                                   moving_variance=T(vec["var"]), epsilon=1e-3)

                def pool1(x):
                    seen["act1"] = x
                    return x
                self.pool1 = pool1
                self.res1 = self.res2 = self.res3 = self.res4 = self.res5 = lambda x, training: x

            def call(self, ldr, training="training"):
                seen["stock_input"] = ldr
                return FakeTFTensor(ldr._t.mean(dim=(1, 2))[:, :64].contiguous())

            def __call__(self, x, training="training"):
                return type(self).call(self, x, training)

        class AEInvcrfDecodeNet:   # noqa: N801
            def invcrf_pca_w_2_invcrf(self, w):
                raise AssertionError("stock TF path must have been replaced")

        class model:            # noqa: N801
            def call(self, img, training="training"):
                raise AssertionError("stock TF path must have been replaced")

            def histogram_layer(self, img, max_bin):
                raise AssertionError("stock TF path must have been replaced")

            @staticmethod
            def _increase(rf):
                raise AssertionError("stock TF path must have been replaced")

        lin.crfFeatureNet, lin.AEInvcrfDecodeNet, lin.model = crfFeatureNet, AEInvcrfDecodeNet, model
        assert tf_adapter.patch(lin, None, fuse_conv1=True) is True

        img = rng.random((2, 40, 56, 3), dtype=np.float32)
        proj = rng.normal(0, 0.05, (64, 11)).astype(np.float32)
        net = lin.model()
        net.crf_feature_net = crfFeatureNet()

        def ae_invcrf_decode_net(feature):
            seen["feature"] = feature
            return lin.AEInvcrfDecodeNet().invcrf_pca_w_2_invcrf(FakeTFTensor(feature._t @ torch.from_numpy(proj).cuda()))
        net.ae_invcrf_decode_net = ae_invcrf_decode_net

        out = net.call(T(img), training=False)
        scale = vec["gamma"] / np.sqrt(vec["var"] + np.float32(1e-3))
        shift = (vec["bias"] - vec["mean"]) * scale + vec["beta"]
        conv = oracle.frontend_conv1(img, kern, None, half_operands=True)
        ref_act = np.maximum(conv * scale + shift, 0.0)
        assert "stock_input" not in seen
        assert seen["act1"].shape == (2, 20, 28, 64)
        assert np.abs(seen["act1"].numpy() - ref_act).max() <= 4e-5 * np.abs(conv).max()
        assert np.abs(seen["feature"].numpy() - ref_act.mean(axis=(1, 2))).max() <= 1e-4
        assert out.shape == (2, 1024)

        net.call(T(img), training=True)                 # training: fp32 front end + the stock (differentiable) layers
        assert seen["stock_input"].shape == (2, 40, 56, 93)
        assert np.array_equal(seen["stock_input"].numpy(), oracle.frontend(img))
        assert (len(PY_FUNCTION_CALLS) > 0) == (not eager)
    finally:
        sys.modules.pop("tensorflow", None)
