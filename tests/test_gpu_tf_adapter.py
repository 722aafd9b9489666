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


class FakeTFTensor:
    """Looks like a TF eager tensor to the adapter: no __dlpack__, module name starts with 'tensorflow'."""
    __module__ = "tensorflow.python.framework.ops"

    def __init__(self, t):
        self._t = t
        self.shape = tuple(t.shape)

    def numpy(self):
        return self._t.cpu().numpy()


def _install_fake_tf():
    import torch
    tf = types.ModuleType("tensorflow")
    tf.float32 = "float32"
    tf.executing_eagerly = lambda: True
    tf.py_function = lambda fn, args, dtype: fn(*args)
    dl = types.SimpleNamespace(
        to_dlpack=lambda x: torch.utils.dlpack.to_dlpack(x._t),
        from_dlpack=lambda cap: FakeTFTensor(torch.utils.dlpack.from_dlpack(cap)))
    tf.experimental = types.SimpleNamespace(dlpack=dl)
    sys.modules["tensorflow"] = tf
    return tf


def test_patch_reference_modules(shdr_gpu, emor, tmp_path, monkeypatch):
    import torch
    from shdr import tf_adapter
    _install_fake_tf()
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

        monkeypatch.setenv("SHDR_NATIVE", "0")          # A/B switch leaves the stock ops alone
        assert tf_adapter.patch(lin, tfu) is False
    finally:
        sys.modules.pop("tensorflow", None)
