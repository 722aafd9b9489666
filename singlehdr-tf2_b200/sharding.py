"""Batch / row-tile sharding of the per-pixel path across the GPUs of one box.

Every batch item is independent in every stage (front end: per-image stencil,
linearization_net.py:312-350; curve: per batch row, :231-253, 369-392; apply_rf uses
rf[b] for image b only, tf_utils.py:61-68), so ranks get disjoint item ranges and the
hot path needs NO collective.  The only exchange offered is an optional gather of the
small outputs (curves) for a caller that wants them in one place; it goes through
whatever ``torch.distributed`` process group the launcher created (NCCL on GPUs, gloo
in the CPU tests) and is off the timed path.
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


def gather_to_all(local, group=None):
    """All-gather equally-shaped per-rank tensors (e.g. curves ``[b_local,1024]``) with
    ``torch.distributed`` -- the optional NCCL gather of outputs.  ``local`` is a torch tensor
    (CUDA for NCCL, CPU for gloo).  Ragged shards are padded to the largest and trimmed."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    n_local = torch.tensor([local.shape[0]], device=local.device, dtype=torch.int64)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
    m = max(counts)
    pad = local
    if local.shape[0] < m:
        pad = torch.cat([local, local.new_zeros((m - local.shape[0],) + tuple(local.shape[1:]))], 0)
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad.contiguous(), group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], 0)
