"""torch.distributed plumbing (one process per GPU, NCCL over NVLink; gloo on CPU for the
tests): process-group bring-up from the torchrun environment, the round-robin batch shard of
the covariance accumulation, and the layer-sharded computation the projector build uses
(``SGDNSCL.get_eigens``).  The collective of the accumulation itself - one SUM all-reduce of
the flat accumulator arena (nsrunner_roi_replay.py:746-749) - is ``CovarianceHooks.all_reduce``;
the class-sharded prototype build is ``prototypes.build_prototypes_sharded``."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise from RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 or dist.is_initialized():
        return int(os.environ.get("RANK", "0")), world
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if backend == "nccl":
        torch.cuda.set_device(local)
    dist.init_process_group(backend=backend)
    return dist.get_rank(), dist.get_world_size()


def shard_batches(n_batches: int, rank: int, world: int):
    """Batch indices of this rank: r, r+W, ... (DefaultSampler-style round robin,
    SURVEY.md 8e)."""
    return list(range(rank, n_batches, world))


def shard_by_cost(costs, world: int):
    """Owner rank of every item, longest-processing-time first: items sorted by descending
    cost, each to the currently least-loaded rank (ties -> lowest rank).  Deterministic, so
    every rank computes the same plan without communication (projector build
    shards by layer, balance by d^3)."""
    load = [0.0] * max(1, world)
    owner = [0] * len(costs)
    for i in sorted(range(len(costs)), key=lambda k: (-float(costs[k]), k)):
        r = min(range(len(load)), key=lambda q: (load[q], q))
        owner[i] = r
        load[r] += float(costs[i])
    return owner


def sharded_compute(shapes, costs, compute, device, dtype=torch.float32, group=None):
    """``compute(i) -> tuple of tensors`` (shapes[i] = their shapes) evaluated by item i's owner
    only; every rank ends up with every result.  One flat buffer per owner travels in ONE
    broadcast (NCCL over NVLink; gloo in the CPU tests) - ``world`` collectives in total,
    all in flight together.  Returns list[tuple[Tensor]] (views of the flat buffers)."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    owner = shard_by_cost(costs, world)
    numel = [[int(torch.Size(sh).numel()) for sh in shp] for shp in shapes]
    flat = [torch.empty(sum(sum(numel[i]) for i in range(len(shapes)) if owner[i] == r),
                        dtype=dtype, device=device) for r in range(world)]
    off = [0] * world
    out = []
    for i, shp in enumerate(shapes):
        r = owner[i]
        views = []
        for sh, n in zip(shp, numel[i]):
            views.append(flat[r][off[r]:off[r] + n].view(sh))
            off[r] += n
        if r == rank:
            for v, t in zip(views, compute(i)):
                v.copy_(t)
        out.append(tuple(views))
    if world > 1:
        works = [dist.broadcast(flat[r], src=dist.get_global_rank(group, r) if group is not None
                                else r, group=group, async_op=True)
                 for r in range(world) if flat[r].numel()]
        for w in works:
            w.wait()
    return out
