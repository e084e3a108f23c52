"""Multi-GPU plumbing: one process per GPU, envs sharded by rank, one collective per minibatch.

Rank r owns global envs [r*N, (r+1)*N); parameters are replicated (identical init from the same torch
seed); after each minibatch's backward the flat fp32 gradient is summed across ranks with one NCCL
all-reduce over NVLink and divided by `world` inside the clip+Adam kernel, so that clipping sees the
averaged gradient exactly like single-process `clip_grad_norm_` (deep_rl/ppo.py:191).  On CPU-only
hosts the same helpers run over gloo (tests).
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as td


def init_from_env(backend: str | None = None) -> Tuple[int, int]:
    """Initialise torch.distributed from torchrun's env (RANK/LOCAL_RANK/WORLD_SIZE/MASTER_*).
    Returns (rank, world).  A single process (no WORLD_SIZE or WORLD_SIZE=1) stays uninitialised."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world <= 1:
        return 0, 1
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
    if not td.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            td.init_process_group(backend=backend, rank=rank, world_size=world,
                                  device_id=torch.device("cuda", local_rank))
        else:
            td.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


def shard_envs(total_envs: int, rank: int, world: int) -> Tuple[int, int]:
    """(first global env id, env count) of `rank` when `total_envs` are split evenly."""
    if total_envs % world != 0:
        raise ValueError(f"total_envs={total_envs} is not divisible by world={world}")
    n = total_envs // world
    return rank * n, n


def all_reduce_sum(t: torch.Tensor) -> torch.Tensor:
    """In-place SUM all-reduce on the current stream (NCCL) / blocking (gloo)."""
    if td.is_initialized() and td.get_world_size() > 1:
        td.all_reduce(t, op=td.ReduceOp.SUM)
    return t


def all_reduce_max(x: float, device) -> float:
    if td.is_initialized() and td.get_world_size() > 1:
        t = torch.tensor([x], dtype=torch.float64, device=device)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t.item())
    return x


def barrier() -> None:
    if td.is_initialized() and td.get_world_size() > 1:
        td.barrier()


def shutdown() -> None:
    if td.is_initialized():
        td.destroy_process_group()
