"""Multi-GPU plumbing: env-index sharding and the end-of-rollout statistics gather.

Envs never read each other's state, so a batch shards by env index with no collective in ``step``/``reset``
(SURVEY.md section 8e).  Philox reset streams are keyed by *global* env id (``env_id_offset + i``), so results do not
depend on how many ranks the batch is split over.  The only collective is the optional all-reduce of episode
statistics at the end of a rollout (``torch.distributed``; NCCL on GPUs, gloo in the CPU tests).
"""

from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist


def shard_envs(total_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous shard [offset, offset + count) of ``total_envs`` for ``rank``; counts differ by at most one."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    base, rem = divmod(int(total_envs), int(world_size))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def gather_episode_stats(returns: Optional[torch.Tensor], group: Optional[dist.ProcessGroup] = None, moments=None) -> Dict[str, float]:
    """Global mean / std / min / max / count of episode returns across all ranks (one small all-reduce each).

    Either ``returns`` (a tensor of per-episode returns held by this rank) or ``moments`` = (count, sum, sum of
    squares, min, max) accumulated on the device by a running monitor (``rollout.VecJitterbugEnv``)."""
    if moments is None:
        r = returns.detach().double().reshape(-1)
        dev = r.device
        cnt, s1, s2 = torch.tensor(float(r.numel()), dtype=torch.float64, device=dev), r.sum(), (r * r).sum()
        lo = r.min() if r.numel() else torch.tensor(float("inf"), dtype=torch.float64, device=dev)
        hi = r.max() if r.numel() else torch.tensor(float("-inf"), dtype=torch.float64, device=dev)
    else:
        cnt, s1, s2, lo, hi = [m.detach().double().clone() for m in moments]
    sums = torch.stack([s1, s2, cnt])
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    n = float(sums[2])
    mean = float(sums[0]) / max(n, 1.0)
    var = max(float(sums[1]) / max(n, 1.0) - mean * mean, 0.0)
    return {"mean": mean, "std": var ** 0.5, "min": float(lo), "max": float(hi), "count": n}


def make_sharded_env(task, total_envs: int, device=None, seed: int = 0, **kwargs):
    """This rank's shard of a ``total_envs``-env batch (RANK / WORLD_SIZE / LOCAL_RANK from the environment)."""
    import os
    from .batched_env import BatchedJitterbugEnv
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    offset, count = shard_envs(total_envs, rank, world)
    if not isinstance(task, str):
        task = list(task)[offset:offset + count]
    return BatchedJitterbugEnv(task, num_envs=count, device=device or f"cuda:{local}", seed=seed,
                               env_id_offset=offset, **kwargs)
