"""Multi-GPU plumbing: environments shard trivially (one process per GPU, no data-path collective); the only
exchange is an all-reduce of episode statistics.  torch.distributed is used as is
(NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import torch

STAT_KEYS = ("episodes", "return_sum", "length_sum", "nonfinite", "overflow")


def shard_seed(base_seed: int, rank: int, envs_per_rank: int) -> int:
    """Env i of rank r is global env r*envs_per_rank + i and is seeded base_seed + global index, so the union of
    all shards is independent of the number of GPUs."""
    return int(base_seed) + int(rank) * int(envs_per_rank)


def global_env_ids(rank: int, envs_per_rank: int) -> range:
    return range(rank * envs_per_rank, (rank + 1) * envs_per_rank)


def allreduce_stats(stats: dict, device=None, group=None) -> dict:
    """Sum the per-rank episode statistics over all ranks (one <= 64-byte message)."""
    import torch.distributed as dist

    vec = torch.tensor([float(stats[k]) for k in STAT_KEYS], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    out = {k: float(v) for k, v in zip(STAT_KEYS, vec.tolist())}
    out["mean_return"] = out["return_sum"] / out["episodes"] if out["episodes"] else None
    out["mean_length"] = out["length_sum"] / out["episodes"] if out["episodes"] else None
    return out


def max_over_ranks(value: float, device=None, group=None) -> float:
    import torch.distributed as dist

    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
