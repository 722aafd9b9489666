"""Batch / row-tile sharding of the per-pixel path across the GPUs of one box.

Every batch item is independent in every stage (front end: per-image stencil,
linearization_net.py:312-350; curve: per batch row, :231-253, 369-392; apply_rf uses
rf[b] for image b only, tf_utils.py:61-68), so ranks get disjoint item ranges and the
hot path needs NO collective.  An optional gather of the small outputs (curves) for a
caller that wants them in one place is launcher plumbing, not product code: see
``tools/dist_gather.py`` (``torch.distributed`` all-gather: NCCL on GPUs, gloo in the CPU
tests), which this package does not import.
"""
from __future__ import annotations


def shard_range(n_items: int, world: int, rank: int):
    """Contiguous, balanced split: the first ``n_items % world`` ranks get one extra item."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(int(n_items), world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def row_tiles(h: int, parts: int, halo_before: int, halo_after: int):
    """Split ``h`` image rows into ``parts`` tiles for a single huge frame.

    Returns ``[(out_y0, out_y1, in_y0, in_y1), ...]``: rows a rank writes and the rows it must
    read (its tile plus a read-only halo clipped to the image: 1/1 for Sobel, 7/8 for the
    pooled histogram).  The halo is re-read from the replicated input, so no exchange step.
    NOTE: REFLECT / border-count semantics apply at TRUE image borders only; a rank whose
    tile is interior passes the halo rows as real data.
    """
    tiles = []
    for r in range(parts):
        y0, y1 = shard_range(h, parts, r)
        tiles.append((y0, y1, max(0, y0 - halo_before), min(h, y1 + halo_after)))
    return tiles
