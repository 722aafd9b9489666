"""CPU oracle for the Linearization-Net per-pixel path -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED w.r.t. TensorFlow's primitives: the reference
(ShinYwings/SingleHDR-tf2) delegates every arithmetic step of this path to
TensorFlow >= 2.4 (un-vendored, unpinned, ``README.md:100``), TensorFlow cannot
be installed in the build image, and the reference ships no tests, golden
vectors or fixtures.  The oracle restates the reference's Python op-for-op (same
fp32 rounding points, same op order) from the call sites cited in every
function.  It IS pinned, bit for bit, against the unmodified reference source
files executed on a NumPy stand-in for TensorFlow (``tf_reference.py``,
``standin/tensorflow``, ``tests/golden/ref_standin.npz``,
``tests/test_reference_golden.py``), which fixes op order, constants, indexing
and channel order; ``tools/make_tf_golden.py --backend tf`` adds the real-TF
vectors the day a TensorFlow install exists.  Further anchors:

* the known-answer material the reference does hold (``figure/lin2.png`` B=5
  soft-histogram example; structural facts of ``invemor.txt``), and
* two independent cross-checks: an fp64 "truth" evaluation of the same
  formulas, and stock CPU PyTorch ops with the same published semantics
  (reflect-pad depthwise conv, ``avg_pool2d(count_include_pad=False)``,
  ``cumsum``) -- see ``tests/test_oracle.py``.

Two independent restatements live here and are cross-checked against each other bit for
bit (``tests/test_oracle_c.py``): ``np_oracle.py`` (NumPy, one full-tensor op per TF op) and
``shdr_oracle.c`` (plain C, OpenMP; built by ``make -C oracle`` / ``__graft_entry__.build()``,
wrapped by ``c_oracle.py``; also the CPU baseline that ``bench.py`` times).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product package
(``singlehdr-tf2_b200``) never does: it fails loudly when its CUDA library is
missing instead of falling back to anything here.
"""
from .np_oracle import (  # noqa: F401
    BINS, S, NCOMP,
    sobel_edges6, histogram_layer, avg_pool_same, frontend, hist_multi,
    parse_invemor, parse_table, invcrf_pca_w_2_invcrf, increase, apply_rf,
    linearize, hist_centers,
    clip01, alpha_mask, linearize_ex, synth_ldr,
    bf16_round, half_round, conv2d_same_s2, frontend_conv1,
    apply_rf_grad, increase_grad, invcrf_pca_grad, histogram_layer_grad, sobel_edges6_grad, frontend_grad,
)
