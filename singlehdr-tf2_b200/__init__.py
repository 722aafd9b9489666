"""singlehdr-tf2_b200 -- B200-native (sm_100a) Linearization-Net per-pixel path of
SingleHDR-tf2: Sobel + spatial-aware soft histogram front end (optionally fused with the
16x16 'same' average pool) and the EMoR inverse-CRF stage (PCA reconstruction, monotonic
enforcement, per-pixel curve lookup), behind the reference's own Python call surface.

Import name: ``shdr`` (see ``shdr.py`` at the repo root; the directory name carries a hyphen).
Importing this package loads ``libshdr.so`` and fails loudly if it has not been built.
"""
from . import _native
from ._native import ShdrError, device_count, launch_count, require_gpu   # noqa: F401
from .device import DeviceArray, PinnedArray, Stream, Event, synchronize  # noqa: F401
from .layers import (                                                      # noqa: F401
    BINS, POOL_K, sobel_edges6, histogram_layer, frontend, frontend_bf16, frontend_f16, hist_multi, conv1_pack_weights, frontend_conv1, parse_invemor,
    set_emor_table, invcrf_pca_w_2_invcrf, invcrf_build, _increase, apply_rf, linearize, linearize_ex, synth_ldr,
    apply_rf_bwd, _increase_bwd, invcrf_build_bwd, frontend_bwd, histogram_layer_bwd,
    AEInvcrfDecodeNet, model,
)
from .host import (                                                        # noqa: F401
    HostPipeline, frontend_host, frontend_conv1_host, hist_multi_host, linearize_host, apply_rf_host,
)
from .sharding import shard_range, row_tiles                              # noqa: F401

__version__ = "0.2.0"
