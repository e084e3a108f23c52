"""N-env CPU port of the PPO update -- CPU ORACLE / BASELINE, test infrastructure only.

The reference script (`deep_rl/ppo.py`) drives ONE env; this is its algorithm on the N-env workload the CUDA path is
benchmarked on, assembled from the pinned oracle pieces: batched model forward and loss with stock PyTorch
CPU ops on all host threads (`ppo_oracle.py`, ppo.py:34-59,166-187), env dynamics / sampler / GAE / permutation in C
(`drl_oracle.c`, ppo.py:110-151,155), clip + Adam through `torch.optim.Adam` (ppo.py:189-192).  Used by
`bench.py --impl reference` as the all-host-threads CPU arm on the benchmarked configuration; never imported by the product.
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import clib
from . import ppo_oracle as po


class VectorPort:
    def __init__(self, env_id: str = "CartPole-v1", num_envs: int = 4096, num_steps: int = 128, seed: int = 1, hidden: int = 64,
                 update_epochs: int = 4, num_minibatches: int = 4, learning_rate: float = 2.5e-4, gamma: float = 0.99,
                 gae_lambda: float = 0.95, max_grad_norm: float = 0.5):
        self.env = clib.OracleVecEnv(env_id, num_envs, seed)
        self.N, self.T, self.H, self.seed = num_envs, num_steps, hidden, seed
        self.O, self.A = self.env.obs_dim, self.env.num_actions
        self.E, self.n_mb = update_epochs, num_minibatches
        self.lr, self.gamma, self.lam, self.max_norm = learning_rate, gamma, gae_lambda, max_grad_norm
        self.params = po.init_params(self.O, self.H, self.A, seed).numpy().copy()
        self.m = np.zeros_like(self.params)
        self.v = np.zeros_like(self.params)
        self.adam_step = 0
        self.update_idx = 0
        self.obs = self.env.reset()
        self.carry = None
        self.env_steps = 0

    def update(self, lr: float | None = None) -> dict:
        N, T, O, A, H = self.N, self.T, self.O, self.A, self.H
        lr = self.lr if lr is None else lr
        ro = po.rollout(self.params, self.env, self.obs, T, H, self.carry)            # ppo.py:110-141
        self.obs, self.carry = ro["obs"][T], (ro["rew"][T], ro["done"][T])
        adv, ret = clib.gae(ro["rew"], ro["done"], ro["val"], self.gamma, self.lam)   # ppo.py:144-151
        B = N * T
        M = (B + self.n_mb - 1) // self.n_mb
        obs = ro["obs"][:T].reshape(B, O)
        act, logp, val = ro["act"][:T].reshape(B), ro["logp"][:T].reshape(B), ro["val"][:T].reshape(B)
        advf, retf = adv[:T].reshape(B), ret[:T].reshape(B)
        terms = None
        for epoch in range(self.E):
            idx = clib.permutation(B, self.seed, self.update_idx * self.E + epoch, 0).astype(np.int64)   # ppo.py:155
            for k in range(self.n_mb):
                sel = idx[k * M:(k + 1) * M]
                terms, grad = po.minibatch_loss_and_grad(self.params, obs[sel], act[sel], logp[sel], advf[sel], retf[sel], val[sel],
                                                         O, H, A)                    # ppo.py:159-190
                self.adam_step += 1
                self.params, self.m, self.v, _ = po.clip_adam(self.params, grad, self.m, self.v, self.adam_step, lr, self.max_norm)
        self.update_idx += 1
        self.env_steps += B
        return {"loss": terms[0], "episodes": len(ro["episodes"])}


def time_vector_port(env_id: str, num_envs: int, num_steps: int, updates: int, warmup: int = 1, threads: int | None = None) -> dict:
    if threads:
        torch.set_num_threads(threads)
    vp = VectorPort(env_id, num_envs, num_steps)
    for _ in range(warmup):
        vp.update()
    per = []
    for _ in range(updates):
        t0 = time.perf_counter()
        vp.update()
        per.append(time.perf_counter() - t0)
    return {"env_steps_per_update": num_envs * num_steps, "seconds": per, "threads": torch.get_num_threads()}
