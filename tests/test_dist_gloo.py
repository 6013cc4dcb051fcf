"""world_size-2 gloo test of the multi-rank path: walker sharding + trace gather (CPU)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, W, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from smol_b200.dist import gather_walker_axis, shard_walkers
    start, count = shard_walkers(W, world, rank)
    S = 3
    # a fake local trace whose values encode (sample, global walker id)
    ids = torch.arange(start, start + count, dtype=torch.float64)
    local = ids[None, :, None] + 1000.0 * torch.arange(S, dtype=torch.float64)[:, None, None]
    local = local.expand(S, count, 2).contiguous()
    full = gather_walker_axis(local, W, axis=1)
    ok = full.shape == (S, W, 2) and bool(torch.equal(full[1, :, 0], 1000.0 + torch.arange(W, dtype=torch.float64)))
    occ = torch.full((S, count, 5), rank, dtype=torch.int8)
    full_occ = gather_walker_axis(occ, W, axis=1)
    counts = [shard_walkers(W, world, r)[1] for r in range(world)]
    want = torch.cat([torch.full((c,), r, dtype=torch.int8) for r, c in enumerate(counts)])
    ok = ok and bool(torch.equal(full_occ[0, :, 0], want))
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_walker_axis_two_ranks_uneven_shards():
    W, world = 7, 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, W, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
