"""Multi-GPU: walkers shard across ranks, traces are gathered -- nothing else is exchanged.

Walkers are independent Markov chains (one kernel + RNG each in the reference,
``smol/moca/sampler/sampler.py:111-116``), so the path is embarrassingly parallel: rank ``r``
owns a contiguous block of walkers, every rank holds a full replica of the model tables, and
the RNG counter carries the GLOBAL walker id so results do not depend on the number of ranks.
The only collective is an ``all_gather`` of the per-sample traces (NCCL over NVLink on GPUs,
gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_walkers(nwalkers: int, world_size: int, rank: int):
    """Contiguous block ``(start, count)`` of rank ``rank``; earlier ranks take the remainder."""
    base, rem = divmod(int(nwalkers), int(world_size))
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def gather_walker_axis(local, nwalkers: int, group=None, axis: int = 1):
    """All-gather ``local`` (torch tensor, walkers along ``axis``) from every rank.

    Shards may be uneven: every rank pads its block to the largest shard, one
    ``all_gather_into_tensor`` moves the data, the padding is dropped afterwards.
    """
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    counts = [shard_walkers(nwalkers, world, r)[1] for r in range(world)]
    wmax = max(counts)
    x = local.movedim(axis, 0).contiguous()
    if x.shape[0] < wmax:
        pad = torch.zeros((wmax - x.shape[0], *x.shape[1:]), dtype=x.dtype, device=x.device)
        x = torch.cat([x, pad], dim=0)
    out = torch.empty((world * wmax, *x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x, group=group)
    pieces = [out[r * wmax:r * wmax + counts[r]] for r in range(world)]
    return torch.cat(pieces, dim=0).movedim(0, axis).contiguous()


class ShardedSampler:
    """One ``Sampler`` per rank over its block of the global walkers (``torch.distributed``)."""

    def __init__(self, ensemble, nwalkers, seeds, *args, rank=None, world_size=None, group=None,
                 **kwargs):
        import torch.distributed as dist
        from .sampler import Sampler

        if world_size is None:
            world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        if rank is None:
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        if len(seeds) != nwalkers:
            raise ValueError("Number of seeds does not match number of kernels!")
        self.rank, self.world_size, self.group = rank, world_size, group
        self.nwalkers = int(nwalkers)
        self.start, self.count = shard_walkers(nwalkers, world_size, rank)
        self.local = Sampler.from_ensemble(ensemble, *args, nwalkers=self.count,
                                           seeds=list(seeds[self.start:self.start + self.count]),
                                           walker_id_base=self.start, **kwargs)

    def run(self, nsteps, initial_occupancies=None, thin_by=1, resident=None, **kw):
        """Advance this rank's walkers.  ``resident`` (default: on when the process group's backend is NCCL): the
        traces stay on the device (``Sampler.run_device``) and ``gather`` all-gathers them GPU to GPU; otherwise they
        go to the local ``samples`` container as in ``Sampler.run``."""
        import torch.distributed as dist
        if resident is None:
            resident = dist.is_initialized() and dist.get_backend(self.group) == "nccl"
        occ = None
        if initial_occupancies is not None:
            occ = initial_occupancies[self.start:self.start + self.count]
        if resident:
            self._resident = self.local.run_device(nsteps, occ, thin_by=thin_by)
        else:
            self._resident = None
            self.local.run(nsteps, occ if occ is None else np.asarray(occ), thin_by=thin_by, **kw)

    def gather(self, name, device=None):
        """Global ``[S, W, ...]`` trace ``name`` on every rank (a CUDA tensor after a resident run: device
        buffers -> NCCL -> device, nothing passes through the host; ``occupancy`` then holds int8 codes)."""
        import torch
        res = getattr(self, "_resident", None)
        if res is not None:
            t = res[name]
            if name in ("enthalpy", "accepted", "bias", "mod_factor"):
                t = t[:, :, None]                   # the reference's trailing axis (trace.py)
            return gather_walker_axis(t, self.nwalkers, self.group, axis=1)
        arr = self.local.samples.get_trace_value(name, flat=False)
        t = torch.from_numpy(np.ascontiguousarray(arr))
        if device is not None:
            t = t.to(device)
        return gather_walker_axis(t, self.nwalkers, self.group, axis=1)
