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


def merge_minibatch_stats(stats: torch.Tensor, counts: torch.Tensor) -> torch.Tensor:
    """Turn rank-local (mean, unbiased std) of minibatches into those of their union over all ranks, in place.
    stats [..., 2] float32, counts [...] samples per rank-local minibatch.  ONE SUM all-reduce of (n, sum x, sum x^2) in
    fp64 for all minibatches of the update (global advantage statistics, ppo.py:169 on the global
    minibatch)."""
    n = counts.to(torch.float64)
    mean, std = stats[..., 0].to(torch.float64), stats[..., 1].to(torch.float64)
    buf = torch.stack([n, mean * n, std * std * (n - 1.0) + n * mean * mean])
    all_reduce_sum(buf)
    big_n, s1, s2 = buf[0], buf[1], buf[2]
    gmean = s1 / big_n
    gvar = torch.clamp((s2 - s1 * gmean) / (big_n - 1.0), min=0.0)
    stats[..., 0] = gmean.to(stats.dtype)
    stats[..., 1] = torch.sqrt(gvar).to(stats.dtype)
    return stats


def all_reduce_max(x: float, device) -> float:
    if td.is_initialized() and td.get_world_size() > 1:
        t = torch.tensor([x], dtype=torch.float64, device=device)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t.item())
    return x


def barrier() -> None:
    if td.is_initialized() and td.get_world_size() > 1:
        td.barrier()


class PeerComm:
    """Symmetric peer-memory buffers for the fused multi-GPU minibatch step (drl_comm_t): every rank allocates one
    IPC-exportable buffer, the handles travel through torch.distributed, and each rank maps all peers' buffers so
    the gradient kernel can read them over NVLink.  One node only."""

    def __init__(self, net, rank: int, world: int, device):
        import ctypes as C

        from . import _lib
        if world > 8:
            raise ValueError("PeerComm supports up to 8 ranks on one node")
        self.L = _lib.lib()
        self.rank, self.world = rank, world
        nbytes = int(self.L.drl_comm_bytes(C.byref(net)))
        self.own = C.c_void_p()
        handle = (C.c_char * 64)()
        _lib.check(self.L.drl_comm_alloc(nbytes, C.byref(self.own), handle))
        handles = [None] * world
        td.all_gather_object(handles, bytes(handle))
        self.opened = []
        peers = (C.c_void_p * 8)()
        for r in range(world):
            if r == rank:
                peers[r] = self.own.value
            else:
                p = C.c_void_p()
                buf = (C.c_char * 64).from_buffer_copy(handles[r])
                _lib.check(self.L.drl_comm_open(buf, C.byref(p)))
                peers[r] = p.value
                self.opened.append(p)
        self.error_flag = torch.zeros(1, dtype=torch.int32, device=device)
        self.struct = _lib.CommT(world, rank, peers, 0, self.error_flag.data_ptr())
        self.seq = 0
        td.barrier()

    def next(self):
        """The drl_comm_t for the next minibatch step (sequence number advanced identically on every rank)."""
        self.seq += 1
        self.struct.seq = self.seq
        return self.struct

    def close(self):
        from . import _lib
        torch.cuda.synchronize()
        td.barrier()
        for p in self.opened:
            self.L.drl_comm_close(p)
        self.opened = []
        if self.own:
            self.L.drl_comm_free(self.own)
            self.own = None


def shutdown() -> None:
    if td.is_initialized():
        td.destroy_process_group()
