#!/usr/bin/env python
"""Golden vectors produced by the UNMODIFIED reference source for the per-pixel path.

    python tools/make_tf_golden.py                    # real TensorFlow if importable, else the NumPy stand-in
    python tools/make_tf_golden.py --backend standin  # oracle/standin/tensorflow (published TF semantics in NumPy)
    python tools/make_tf_golden.py --backend tf       # real TensorFlow (CPU; GPUs hidden) -- pins the oracle for good

Runs in the BUILD container (it imports /root/reference/{linearization_net,tf_utils}.py through
``oracle/tf_reference.py``; the reference does not exist on the GPU box) and writes
``tests/golden/ref_<backend>.npz``: seeded small inputs for every BASELINE.json configuration plus edge cases, and the
reference's outputs for them.  ``tests/test_reference_golden.py`` checks the oracle (CPU) and the CUDA kernels (GPU)
against every ``ref_*.npz`` it finds -- so committing ``ref_tf.npz`` once TensorFlow is available turns "parity
unpinned" into "parity pinned" without touching a test.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import tf_reference as R  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def cases():
    """name -> inputs.  Small versions of configs[0..4] + the edge cases the reference's arithmetic can hit."""
    rng = np.random.default_rng(2026)
    c = {}
    c["c1_frontend"] = rng.random((1, 24, 24, 3), dtype=np.float32)                      # configs[0] front end
    c["c2_hist"] = rng.random((2, 20, 36, 3), dtype=np.float32)                          # configs[1] (un-pooled part)
    c["c4_frontend"] = rng.random((2, 16, 20, 3), dtype=np.float32)                      # configs[3]
    c["c5_frontend"] = rng.random((1, 18, 32, 3), dtype=np.float32)                      # configs[4], a frame strip
    q = np.round(rng.random((1, 12, 12, 3)) * 255).astype(np.float32) / np.float32(255)   # 8-bit LDR values
    q[0, 0, :6, 0] = [0.0, 1.0, 0.5, 0.25, 0.125, 0.375]                                 # bin edges / centres
    c["edge_quantised"] = q
    wide = (rng.random((1, 6, 6, 3)) * 1.6 - 0.3).astype(np.float32)                     # outside [0, 1]
    c["edge_out_of_range"] = wide
    c["w"] = rng.normal(0, 0.5, (4, 11)).astype(np.float32)                              # configs[2]: non-monotone curves
    c["x_apply"] = (rng.random((4, 10, 14, 3)) * 1.2 - 0.1).astype(np.float32)
    c["rf_small"] = np.cumsum(rng.normal(0.02, 0.05, (3, 17)), axis=1).astype(np.float32)  # generic k, needs the shift
    c["x_small"] = rng.random((3, 33), dtype=np.float32)
    return c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", choices=["tf", "standin"], default=None)
    args = ap.parse_args()
    backend = args.backend or ("tf" if R.have_real_tf() else "standin")
    if not R.available(backend):
        sys.exit(f"cannot run the reference with backend {backend!r}: "
                 f"source under {R.reference_dir()}: {R.have_reference_source()}, real TensorFlow: {R.have_real_tf()}")
    tf, _, _ = R.load(backend)
    c = cases()
    out = {f"in_{k}": v for k, v in c.items()}
    for name in ("c1_frontend", "c4_frontend", "c5_frontend", "edge_quantised", "edge_out_of_range"):
        out[f"out_frontend_{name}"] = R.frontend(c[name], backend)
    out["out_hist_multi_c2_hist"] = R.hist_multi(c["c2_hist"], backend)
    for b in (1, 3, 4, 5, 8, 16, 33):
        out[f"out_hist{b}_c2_hist"] = R.histogram_layer(c["c2_hist"], b, backend)
    _, g0, hinv = R.parse_invemor(backend)
    out["g0"], out["hinv"] = g0, hinv
    pca = R.invcrf_pca_w_2_invcrf(c["w"], backend)
    out["out_pca"] = pca
    out["out_increase"] = R.increase(pca, backend)
    out["out_apply"] = R.apply_rf(c["x_apply"], out["out_increase"], backend)
    out["out_increase_small"] = R.increase(c["rf_small"], backend)
    out["out_apply_small"] = R.apply_rf(c["x_small"], out["out_increase_small"], backend)
    out["meta"] = np.array([backend, getattr(tf, "__version__", "?")])
    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, f"ref_{backend}.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, backend {backend} ({getattr(tf, '__version__', '?')})")


if __name__ == "__main__":
    main()
