"""Sharding helpers of the data-parallel PPO path.

The reference has no collective anywhere: its multi-GPU story is independent processes time-sharing GPUs
(``utils/device_pool.py:45-72``).  Here envs are sharded across ranks and the only exchanges are sums:
the flat gradient per optimizer step, three doubles per update for the advantage normalisation
(``ppo/agent.py:204`` normalises over the WHOLE buffer) and the metric accumulators.  The functions work on
any backend (NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def backend() -> str:
    return dist.get_backend() if dist.is_available() and dist.is_initialized() else ""


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def allreduce_sum_(t: torch.Tensor) -> torch.Tensor:
    """In-place sum over ranks (no-op for a single process)."""
    if world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def shard_envs(total_envs: int, rank_: int, world: int) -> Tuple[int, int]:
    """(env_id_base, num_envs) of a rank: contiguous, equal shards; the global env id keys the Philox draws,
    so the union of the shards reproduces the single-process episodes."""
    if total_envs % world != 0:
        raise ValueError(f"{total_envs} envs do not divide over {world} ranks")
    per = total_envs // world
    return rank_ * per, per


def loss_scale(local_batch: int, world: int) -> float:
    """Weight of one sample's loss term on a rank, so that the SUM over ranks of the local gradients is the
    gradient of the global minibatch mean (equal shard sizes)."""
    return 1.0 / (local_batch * world)


def global_mean_std(stats: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(mean, unbiased std) from all-reduced (sum, sum of squares, count), as ``hrp_adv_normalize`` forms them."""
    s, q, n = stats[0], stats[1], stats[2]
    mean = s / n
    var = (q - n * mean * mean) / (n - 1)
    return mean, torch.sqrt(torch.clamp(var, min=0))
