#!/usr/bin/env python
"""Randomised parity sweep on the GPU: random shapes (tiny, odd, ragged) through the C ABI against the oracle, with
the tolerances of tests/test_gpu_parity.py / test_gpu_conv1.py.  Not part of the pytest suite (run time grows with
--cases); run it after touching a kernel:   python tools/fuzz_gpu.py --cases 60 --seed 1"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
import shdr  # noqa: E402


def rel_ok(a, b, rtol):
    return not (np.abs(a - b) > rtol * np.abs(b)).any()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=40)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--max-side", type=int, default=160)
    args = ap.parse_args()
    shdr.require_gpu()
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "invemor_f32.npz"))
    shdr.set_emor_table(z["g0"], z["hinv"])
    rng = np.random.default_rng(args.seed)
    D = shdr.DeviceArray.from_numpy
    fails, t0 = [], time.time()
    for i in range(args.cases):
        n = int(rng.integers(1, 4))
        h, w = (int(rng.integers(2, args.max_side + 1)) for _ in range(2))
        kind = rng.integers(0, 4)
        img = rng.random((n, h, w, 3), dtype=np.float32)
        if kind == 1:
            img = np.round(img * 255).astype(np.float32) / np.float32(255)       # 8-bit LDR values: bin edges / centres
        elif kind == 2:
            img = np.clip(img * 1.4 - 0.2, -0.1, 1.1).astype(np.float32)         # out-of-range values vote 0
        elif kind == 3:
            img = (np.linspace(0, 1, h * w, dtype=np.float32).reshape(1, h, w, 1) * np.ones((n, 1, 1, 3), np.float32))
        d = D(img)
        tag = f"case {i}: n={n} h={h} w={w} kind={kind}"
        if not np.array_equal(shdr.frontend(d).numpy(), oracle.frontend(img)):
            fails.append(tag + " frontend")
        ref = oracle.hist_multi(img, pool_k=16)
        if not rel_ok(shdr.hist_multi(d, pool=True).numpy(), ref, 1e-5):
            fails.append(tag + " hist_multi pooled")
        got = shdr.frontend(d, pool=True).numpy()
        if not (np.array_equal(got[..., :9], oracle.frontend(img)[..., :9]) and rel_ok(got[..., 9:], ref, 1e-5)):
            fails.append(tag + " frontend pooled")
        wts = rng.normal(0, 0.5, (n, 11)).astype(np.float32)
        y, curve = shdr.linearize(d, D(wts))
        ry, rc = oracle.linearize(img, wts, z["g0"], z["hinv"])
        if not (np.abs(curve.numpy() - rc).max() <= 5e-6 and np.abs(y.numpy() - ry).max() <= 1e-5):
            fails.append(tag + " linearize")
        kern = (rng.normal(0, 1, (7, 7, 93, 64)) / 67.5).astype(np.float32)
        bias = rng.normal(0, 0.1, 64).astype(np.float32)
        c1 = shdr.frontend_conv1(d, shdr.conv1_pack_weights(D(kern)), bias=D(bias)).numpy()
        r1 = oracle.frontend_conv1(img, kern, bias, half_operands=True)
        if not np.abs(c1 - r1).max() <= 2e-5 * max(np.abs(r1).max(), 1e-3):
            fails.append(tag + f" conv1 ({np.abs(c1 - r1).max() / np.abs(r1).max():.2e})")
    print(f"{args.cases} random cases x 5 ops in {time.time() - t0:.0f} s: {len(fails)} failures")
    for f in fails:
        print("FAIL", f)
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
