#!/usr/bin/env python
"""Generate the committed fixtures under tests/golden/.

Run in the BUILD container (it reads /root/reference/invemor.txt, which does
not exist on the GPU box):

    python tools/make_golden.py

Outputs
  tests/golden/invemor_f32.npz   B[1024], g0[1024], hinv[1024,11] float32, parsed
                                 from the reference's data table invemor.txt with
                                 the oracle's restatement of parse_invemor
                                 (linearization_net.py:217-227,255-268).
  tests/golden/oracle_small.npz  seeded small inputs and the fp32 oracle's outputs
                                 for every stage (regression anchor for the oracle
                                 and fixture for the GPU parity tests).
  tests/golden/kat_lin2.npz      the B=5 soft-histogram known answers printed in
                                 the reference's figure/lin2.png (typed in by hand
                                 from the figure; NOT produced by any code here).

The reference cannot be imported to produce vectors (it needs TensorFlow, which
is not installable here), so no fixture in this directory is an output of the
reference itself -- parity is unpinned, see oracle/__init__.py.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REF_TABLE = "/root/reference/invemor.txt"


def main():
    os.makedirs(GOLD, exist_ok=True)
    b, g0, hinv = oracle.parse_invemor(REF_TABLE)
    assert b.shape == (1024,) and g0.shape == (1024,) and hinv.shape == (1024, 11)
    np.savez_compressed(os.path.join(GOLD, "invemor_f32.npz"), B=b, g0=g0, hinv=hinv)

    rng = np.random.default_rng(20261018)
    img = rng.random((2, 20, 24, 3), dtype=np.float32)
    # exercise exact bin centres / edges / range ends
    img[0, 0, :8, 0] = np.float32([0, 1, 0.125, 0.25, 0.5, 0.375, 0.0625, 0.9375])
    w = rng.normal(0, 0.5, (2, 11)).astype(np.float32)
    x = rng.random((2, 20, 24, 3), dtype=np.float32)
    x[1, 0, :6, 1] = np.float32([0, 1, -0.25, 1.5, 1023.0 / 1024.0, 0.5])
    pca = oracle.invcrf_pca_w_2_invcrf(w, g0, hinv)
    curve = oracle.increase(pca)
    np.savez_compressed(
        os.path.join(GOLD, "oracle_small.npz"),
        img=img, w=w, x=x,
        edges=oracle.sobel_edges6(img),
        hist4=oracle.histogram_layer(img, 4),
        hist5=oracle.histogram_layer(img, 5),
        hist16_pooled=oracle.avg_pool_same(oracle.histogram_layer(img, 16)),
        frontend=oracle.frontend(img),
        frontend_pooled=oracle.frontend(img, pool_k=16),
        pca=pca, curve=curve, lin=oracle.apply_rf(x, curve),
    )

    # figure/lin2.png: B = 5, bins centred .1 .3 .5 .7 .9
    vals = np.float32([.63, .65, .32, .84, .15])
    votes = np.zeros((5, 5), np.float32)          # [value, bin]
    votes[0, 2], votes[0, 3] = .35, .65
    votes[1, 2], votes[1, 3] = .25, .75
    votes[2, 1], votes[2, 2] = .9, .1
    votes[3, 3], votes[3, 4] = .3, .7
    votes[4, 0], votes[4, 1] = .75, .25
    np.savez_compressed(os.path.join(GOLD, "kat_lin2.npz"), values=vals, votes=votes)
    print("wrote", sorted(os.listdir(GOLD)))


if __name__ == "__main__":
    main()
