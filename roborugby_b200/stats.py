"""Sharding and the optional episode-statistics reduction.

Episodes are independent (the reference runs one env per process), so the batch
shards trivially: rank r owns a contiguous slice of envs and there is NO collective on the step
path.  The only exchange is an optional all-reduce(sum) of the 8-double statistics vector
(include/rr_b200.h RR_STAT_*), issued through torch.distributed — NCCL over NVLink 5 / NVSwitch
on GPUs, gloo in the CPU tests.  64 bytes: latency-bound, bandwidth irrelevant.
"""
import torch

from ._lib import STAT_NAMES


def shard_envs(total_envs, rank, world_size):
    """Contiguous slice of `total_envs` owned by `rank`: returns (n_local, env_offset).

    The offset is fed to rr_config.env_offset so that the Philox stream of every env depends on its
    GLOBAL index only: results are identical however many ranks the batch is split over."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(int(total_envs), int(world_size))
    n_local = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return n_local, offset


def allreduce_stats(stats, group=None):
    """Sum a statistics vector over all ranks; returns a dict of global statistics.

    `stats` is the per-rank tensor (CUDA for NCCL, CPU for gloo); it is not modified."""
    import torch.distributed as dist
    out = stats.detach().clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return summarize(out)


def summarize(vec):
    v = [float(x) for x in (vec.tolist() if isinstance(vec, torch.Tensor) else vec)]
    d = dict(zip(STAT_NAMES, v))
    ep = d["episodes"]
    d["mean_return_happy"] = d["return_happy"] / ep if ep else 0.0
    d["mean_return_grumpy"] = d["return_grumpy"] / ep if ep else 0.0
    d["mean_length"] = d["length"] / ep if ep else 0.0
    return d
