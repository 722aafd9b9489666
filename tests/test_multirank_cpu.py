"""world_size-2 gloo test of the N>1 host logic: every rank takes its shard of the batch,
produces its part (here with the oracle standing in for the kernels -- no GPU in this
container), and the optional gather reassembles exactly the single-rank result."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, out_q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    import shdr
    z = np.load(os.path.join(ROOT, "tests", "golden", "invemor_f32.npz"))
    w = np.random.default_rng(99).normal(0, 0.5, (n_items, 11)).astype(np.float32)
    a, b = shdr.shard_range(n_items, world, rank)
    local = oracle.increase(oracle.invcrf_pca_w_2_invcrf(w[a:b], z["g0"], z["hinv"])) if b > a \
        else np.zeros((0, 1024), np.float32)
    from tools.dist_gather import gather_to_all
    full = gather_to_all(torch.from_numpy(local))
    dist.barrier()
    if rank == 0:
        out_q.put(full.numpy())
    dist.destroy_process_group()


def test_sharded_curves_gather_gloo():
    import oracle
    world, n_items = 2, 5      # ragged: 3 + 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    z = np.load(os.path.join(ROOT, "tests", "golden", "invemor_f32.npz"))
    w = np.random.default_rng(99).normal(0, 0.5, (n_items, 11)).astype(np.float32)
    want = oracle.increase(oracle.invcrf_pca_w_2_invcrf(w, z["g0"], z["hinv"]))
    assert got.shape == want.shape and np.array_equal(got, want)
