"""Parity against vectors produced by the UNMODIFIED reference source (tools/make_tf_golden.py).

``tests/golden/ref_standin.npz`` -- the reference's Python executed on the NumPy stand-in for TensorFlow
(oracle/standin/tensorflow): pins the oracle's restatement of the reference's op order, constants and indexing.
``tests/golden/ref_tf.npz`` -- the same from real TensorFlow; absent until a TensorFlow install exists (it is not
installable in this project's image), and the tests that need it skip with that reason.  Every ``ref_*.npz`` present
is checked: the oracle on the CPU, the CUDA kernels on the GPU.  A further CPU test re-runs the reference source live
when ``/root/reference`` exists (it does in the build container, not on the GPU box)."""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
FILES = sorted(glob.glob(os.path.join(GOLD, "ref_*.npz")))
FRONT = ("c1_frontend", "c4_frontend", "c5_frontend", "edge_quantised", "edge_out_of_range")
# real TensorFlow may order the additions inside a primitive differently from the oracle (reduce_sum, matmul, Sobel
# taps): the north_star tolerances apply there; the stand-in shares the oracle's primitives, so it must match bit for bit
TOL = {"standin": dict(feat=0.0, pca=0.0, curve=0.0, lin=0.0), "tf": dict(feat=1e-6, pca=1e-6, curve=5e-6, lin=1e-5)}


def backend_of(path):
    return os.path.basename(path)[4:-4]


def near(a, b, tol):
    if tol == 0.0:
        return np.array_equal(a, b)
    return a.shape == b.shape and float(np.abs(a - b).max()) <= tol


def test_standin_golden_is_committed():
    assert any(backend_of(f) == "standin" for f in FILES), "tests/golden/ref_standin.npz missing: run tools/make_tf_golden.py"


def test_real_tf_golden_present_or_explained():
    if not any(backend_of(f) == "tf" for f in FILES):
        pytest.skip("tests/golden/ref_tf.npz absent: TensorFlow is not installable in this image (no wheel, no network); "
                    "`python tools/make_tf_golden.py --backend tf` creates it and these tests then gate on it")


@pytest.mark.parametrize("path", FILES, ids=backend_of)
def test_oracle_matches_reference_vectors(path):
    z = np.load(path)
    t = TOL[backend_of(path)]
    for name in FRONT:
        assert near(oracle.frontend(z[f"in_{name}"]), z[f"out_frontend_{name}"], t["feat"]), name
    assert near(oracle.hist_multi(z["in_c2_hist"]), z["out_hist_multi_c2_hist"], t["feat"])
    for b in (1, 3, 4, 5, 8, 16, 33):
        assert near(oracle.histogram_layer(z["in_c2_hist"], b), z[f"out_hist{b}_c2_hist"], t["feat"]), b
    pca = oracle.invcrf_pca_w_2_invcrf(z["in_w"], z["g0"], z["hinv"])
    assert near(pca, z["out_pca"], t["pca"])
    assert near(oracle.increase(z["out_pca"]), z["out_increase"], t["curve"])
    assert near(oracle.apply_rf(z["in_x_apply"], z["out_increase"]), z["out_apply"], t["lin"])
    assert near(oracle.increase(z["in_rf_small"]), z["out_increase_small"], t["curve"])
    assert near(oracle.apply_rf(z["in_x_small"], z["out_increase_small"]), z["out_apply_small"], t["lin"])


def test_table_parse_matches_reference(emor):
    z = np.load(os.path.join(GOLD, "ref_standin.npz"))
    assert np.array_equal(z["g0"], emor[1]) and np.array_equal(z["hinv"], emor[2])


def test_reference_source_live_equals_golden_and_oracle():
    """Re-run the reference's own source (NumPy stand-in backend) in a fresh interpreter and compare with the committed
    vectors and the oracle -- proves the probe path of oracle/tf_reference.py and that the golden file is current."""
    from oracle import tf_reference as R
    if not R.have_reference_source():
        pytest.skip(f"reference source not under {R.reference_dir()} (GPU box)")
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "from oracle import tf_reference as R; import oracle\n"
        "z = np.load(%r)\n"
        "assert R.available('standin')\n"
        "for n in %r:\n"
        "    f = R.frontend(z['in_' + n], 'standin')\n"
        "    assert np.array_equal(f, z['out_frontend_' + n]) and np.array_equal(f, oracle.frontend(z['in_' + n])), n\n"
        "y, c = R.linearize(z['in_x_apply'], z['in_w'], 'standin')\n"
        "assert np.array_equal(c, z['out_increase']) and np.array_equal(y, z['out_apply'])\n"
        "print('OK')\n" % (ROOT, os.path.join(GOLD, "ref_standin.npz"), FRONT))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


def test_bench_cpu_arm_uses_the_reference_when_it_can():
    """bench.py's CPU legs time the reference's own functions (kind 'reference') iff real TensorFlow and the reference
    source are both present; otherwise the oracle port (kind 'port')."""
    sys.path.insert(0, ROOT)
    import bench
    from oracle import tf_reference as R
    fn, px, what, kind = bench.cpu_step_fn("config1", 1)
    assert kind == ("reference" if R.available("tf") else "port")
    assert px == 256 * 256 and callable(fn)
    if kind == "port":
        assert "TensorFlow" in what


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=backend_of)
def test_kernels_match_reference_vectors(shdr_gpu, path):
    z = np.load(path)
    D = shdr_gpu.DeviceArray.from_numpy
    for name in FRONT:
        got = shdr_gpu.frontend(D(z[f"in_{name}"])).numpy()
        assert float(np.abs(got - z[f"out_frontend_{name}"]).max()) <= 1e-6, name
    got = shdr_gpu.hist_multi(D(z["in_c2_hist"])).numpy()
    assert float(np.abs(got - z["out_hist_multi_c2_hist"]).max()) <= 1e-6
    for b in (1, 3, 4, 5, 8, 16, 33):
        got = shdr_gpu.histogram_layer(D(z["in_c2_hist"]), b).numpy()
        assert float(np.abs(got - z[f"out_hist{b}_c2_hist"]).max()) <= 1e-6, b
    shdr_gpu.set_emor_table(z["g0"], z["hinv"])
    pca = shdr_gpu.invcrf_pca_w_2_invcrf(D(z["in_w"])).numpy()
    assert float(np.abs(pca - z["out_pca"]).max()) <= 1e-6
    curve = shdr_gpu._increase(D(z["out_pca"])).numpy()
    assert float(np.abs(curve - z["out_increase"]).max()) <= 5e-6
    y = shdr_gpu.apply_rf(D(z["in_x_apply"]), D(z["out_increase"])).numpy()
    assert float(np.abs(y - z["out_apply"]).max()) <= 1e-5
    y2, c2 = shdr_gpu.linearize(D(z["in_x_apply"]), D(z["in_w"]))
    assert float(np.abs(c2.numpy() - z["out_increase"]).max()) <= 5e-6
    assert float(np.abs(y2.numpy() - z["out_apply"]).max()) <= 1e-5
    cs = shdr_gpu._increase(D(z["in_rf_small"])).numpy()
    assert float(np.abs(cs - z["out_increase_small"]).max()) <= 5e-6
    ys = shdr_gpu.apply_rf(D(z["in_x_small"]), D(z["out_increase_small"])).numpy()
    assert float(np.abs(ys - z["out_apply_small"]).max()) <= 1e-5
