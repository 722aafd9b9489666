"""Optional gather of per-rank outputs: launcher-side plumbing for bench.py and the world-size-2
CPU test, deliberately OUTSIDE the product package (which imports neither torch nor a process group).  Goes through
whatever ``torch.distributed`` process group the launcher created: NCCL all-gather on GPUs, gloo in the CPU tests."""


def gather_to_all(local, group=None):
    """All-gather equally-shaped per-rank tensors (e.g. curves ``[b_local,1024]``).  ``local`` is a torch tensor (CUDA
    for NCCL, CPU for gloo).  Ragged shards are padded to the largest and trimmed."""
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
