"""Data-parallel host logic for the rollout: one process per GPU, episodes sharded by rank, weights
replicated, ONE all-reduce(sum) of the flat gradient buffers per optimizer step; no data-path collective.

The reference has no distributed path for r2r_src (single process, README.md:82); its loss normalises the summed
cross-entropy by the batch size (agent_dg.py:1024), so with the global batch split over `world` ranks each rank scales
its loss by ml_weight / (B_local * world) and the gradients are SUMMED — identical to one process running the whole batch.
"""
import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend=None, device=None):
    world, rank, local = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        kw = {}
        if backend == "nccl" and device is not None:
            kw["device_id"] = torch.device(device)
        dist.init_process_group(backend or ("nccl" if torch.cuda.is_available() else "gloo"), **kw)
    return world, rank, local


def shard(n_items, rank, world):
    """Contiguous block of episodes owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def loss_scale(ml_weight, global_batch):
    """Per-rank loss factor: sum-reduced gradients then equal the single-process big-batch gradients."""
    return ml_weight / float(global_batch)


def allreduce_sum_(buffers, world):
    """In-place sum over ranks of every flat buffer (gradients). Returns the number of collectives issued."""
    if world <= 1:
        return 0
    for b in buffers:
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
    return len(buffers)


def broadcast_(buffers, world, src=0):
    if world <= 1:
        return
    for b in buffers:
        dist.broadcast(b, src)


def max_over_ranks(value, device, world):
    t = torch.tensor([float(value)], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
