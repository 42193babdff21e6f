"""Env-batch sharding across the GPUs of one box (one process per GPU, torch.distributed plumbing).

The step path has no cross-GPU traffic: rank r owns the contiguous global env ids [r*n, (r+1)*n) and the
device RNG is keyed by the GLOBAL env id, so results do not depend on the number of ranks.  The only
collective is an all-gather of the 8-double episode-statistics vector (NCCL on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import torch
import torch.distributed as dist

STAT_NAMES = ("env_steps", "episodes", "successes", "return_sum", "length_sum", "nonfinite_resets", "solver_iters", "contacts")


@dataclass(frozen=True)
class Shard:
    rank: int
    world: int
    n_local: int
    first_env_id: int

    @property
    def n_global(self):
        return self.n_local * self.world


def env_shard(total_envs: int | None = None, envs_per_rank: int | None = None, rank: int | None = None, world: int | None = None) -> Shard:
    """Partition either a fixed total (strong scaling; must divide evenly) or a fixed per-rank count (weak scaling)."""
    rank = int(os.environ.get("RANK", "0")) if rank is None else rank
    world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else world
    if (total_envs is None) == (envs_per_rank is None):
        raise ValueError("give exactly one of total_envs / envs_per_rank")
    if total_envs is not None:
        if total_envs % world:
            raise ValueError(f"total_envs={total_envs} is not divisible by world size {world}")
        n = total_envs // world
    else:
        n = envs_per_rank
    if n <= 0 or not 0 <= rank < world:
        raise ValueError("bad shard")
    return Shard(rank, world, n, rank * n)


def gather_stats(local_stats: torch.Tensor) -> torch.Tensor:
    """All-gather the per-rank statistics vector -> (world, 8).  Works without an initialised process group (world = 1)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_stats.reshape(1, -1).clone()
    out = torch.empty((dist.get_world_size(), local_stats.numel()), dtype=local_stats.dtype, device=local_stats.device)
    dist.all_gather_into_tensor(out, local_stats.reshape(1, -1).contiguous())
    return out


def summarize(stats_all: torch.Tensor) -> dict:
    s = stats_all.sum(0).tolist()
    d = dict(zip(STAT_NAMES, s))
    d["success_rate"] = d["successes"] / d["episodes"] if d["episodes"] else 0.0
    d["mean_episode_length"] = d["length_sum"] / d["episodes"] if d["episodes"] else 0.0
    return d
