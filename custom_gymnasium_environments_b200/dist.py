"""Multi-GPU sharding of env batches: one process per GPU, contiguous global env-id slices, and
ONE tiny all-reduce of the integer episode statistics per reporting interval.

The envs are independent: there is no data-path collective.  The RNG is keyed on the GLOBAL env id,
so env i has the same trajectory at world sizes 1, 2, 4 and 8.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

STAT_NAMES = ("n_episodes", "sum_return", "sum_length", "sum_score", "max_score")


def shard_range(total_envs: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous slice [start, start+count) of the global env ids owned by `rank` (ragged tail spread
    over the first ranks)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(int(total_envs), int(world_size))
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def env_from_rank():
    """(rank, local_rank, world_size) from the torchrun environment (1-process defaults)."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))


def init_process_group(backend: str | None = None):
    """Initialise torch.distributed from the torchrun environment if WORLD_SIZE > 1."""
    rank, local_rank, world = env_from_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kwargs["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, **kwargs)
    return rank, local_rank, world


def all_reduce_episode_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """Reduce the int64 vector {n_episodes, sum_return, sum_length, sum_score, max_score} over ranks:
    SUM for the first four, MAX for the last.  Returns a new tensor; a no-op without a process group."""
    out = stats.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        sums = out[:4].contiguous()
        mx = out[4:5].contiguous()
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
        out[:4] = sums
        out[4:5] = mx
    return out


def summarize(stats: torch.Tensor) -> dict:
    v = stats.tolist()
    n = max(v[0], 1)
    return {"episodes": v[0], "episode_return_mean": v[1] / n, "episode_len_mean": v[2] / n,
            "score_mean": v[3] / n, "score_max": v[4] if v[0] else 0}
