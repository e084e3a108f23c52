"""Worker for tests/test_gpu_multi.py: launched by torchrun with one rank per GPU.
Checks that the in-kernel NVLink all-reduce path (grad_allreduce="peer") and the NCCL path give the same training
trajectory, and that all ranks stay bit-identical."""
import os
import sys

import numpy as np
import torch
import torch.distributed as td

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deep_rl_b200 as drl  # noqa: E402
from deep_rl_b200 import dist  # noqa: E402


def run(mode, rank, world, envs, updates, global_stats=False):
    cfg = drl.PPOConfig(num_envs=envs, num_steps=32, seed=3, total_timesteps=envs * 32 * world * 16, grad_allreduce=mode,
                        global_adv_stats=global_stats, update_precision="bf16")      # the in-kernel exchange is part of the tcgen05 kernel
    tr = drl.PPOTrainer(cfg, rank=rank, world=world)
    assert (tr.peer is not None) == (mode == "peer")
    for _ in range(updates):
        tr.update()
    m = tr.metrics()
    torch.cuda.synchronize()
    p = tr.agent.flat_params.clone()
    g = tr.grad.clone()
    if global_stats:     # the statistics every rank used are those of the union of the ranks' minibatches of the last epoch
        B, M = cfg.batch_size, cfg.minibatch_size
        idx = tr.idx[-1].long()
        adv = tr.advantages.flatten()[:B][idx].view(tr.n_mb, M).double()
        mine = torch.stack([adv.sum(1), (adv * adv).sum(1)])
        td.all_reduce(mine)
        n = float(M * world)
        mean = mine[0] / n
        std = torch.sqrt((mine[1] - mine[0] * mean) / (n - 1.0))
        got = tr.adv_stats[-1].double()
        assert torch.allclose(got[:, 0], mean, rtol=1e-4, atol=1e-6) and torch.allclose(got[:, 1], std, rtol=1e-4), (got, mean, std)
    if tr.peer is not None:
        tr.peer.close()
    return p, g, m


def main():
    rank, world = dist.init_from_env()
    envs = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    p_nccl, g_nccl, m_nccl = run("nccl", rank, world, envs, 3)
    p_peer, g_peer, m_peer = run("peer", rank, world, envs, 3)
    # (1) every rank holds the same parameters, bit for bit
    for name, p in (("nccl", p_nccl), ("peer", p_peer)):
        gathered = [torch.empty_like(p) for _ in range(world)]
        td.all_gather(gathered, p)
        for r in range(1, world):
            assert torch.equal(gathered[0], gathered[r]), f"{name}: rank {r} diverged from rank 0"
    # (2) the two exchange paths agree (summation order differs -> fp32 rounding only, amplified by Adam's normalisation)
    d = (p_peer - p_nccl).abs().max().item()
    gd = (g_peer - g_nccl).abs().max().item() / (g_nccl.abs().max().item() + 1e-30)
    assert d < 5e-5, f"peer vs nccl parameters differ by {d}"
    assert abs(m_peer["loss"] - m_nccl["loss"]) < 1e-3 * max(1.0, abs(m_nccl["loss"]))
    # (3) global advantage statistics: ranks stay bit-identical and use the statistics of the union
    p_glob, _, m_glob = run("peer", rank, world, envs, 2, global_stats=True)
    gathered = [torch.empty_like(p_glob) for _ in range(world)]
    td.all_gather(gathered, p_glob)
    for r in range(1, world):
        assert torch.equal(gathered[0], gathered[r]), f"global stats: rank {r} diverged from rank 0"
    assert np.isfinite(m_glob["loss"])
    if rank == 0:
        print(f"MULTI-GPU OK world={world} envs/rank={envs}: max |dp| peer-vs-nccl {d:.2e}, rel |dg| {gd:.2e}, loss {m_peer['loss']:.5f}")
    dist.shutdown()


if __name__ == "__main__":
    main()
