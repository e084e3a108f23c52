"""REINFORCE on the device-resident CartPole -- the B200 counterpart of deep_rl/reinforce.py.

Names and defaults are the script's (reinforce.py:26-47): env_id "CartPole-v1", gamma 0.99, seed 1, policy
`nn.Sequential(nn.Linear(4, 128), nn.Dropout(p=0.6), nn.ReLU(), nn.Linear(128, 2), nn.Softmax(-1))`, Adam(lr=1e-2), 100 episodes.
`num_envs` is the one addition: N environments run one episode each per iteration and share one optimizer step (gradient =
mean over the N episodes); N = 1 is the reference's schedule.  Per iteration:

    episodes (reset, act, step until done)   reinforce.py:55-67   drl_reinforce_episodes   (1 launch)
    discounted reward-to-go                  reinforce.py:67      drl_gae(gae_lambda=1, V=0): the PPO path's reverse-time scan
    normalise, loss, backward                reinforce.py:71-76   drl_reinforce_grad       (2 launches)
    Adam                                     reinforce.py:77      drl_adam_step
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional

import torch
from torch import nn

from . import _lib
from .envs import VecEnv


@dataclass
class ReinforceConfig:
    env_id: str = "CartPole-v1"
    gamma: float = 0.99
    learning_rate: float = 1e-2
    seed: int = 1
    num_episodes: int = 100          # reinforce.py:49 (`for episode_idx in range(100)`): iterations of the loop
    num_envs: int = 1


def make_policy() -> nn.Sequential:
    """reinforce.py:38-44."""
    return nn.Sequential(nn.Linear(4, 128), nn.Dropout(p=0.6), nn.ReLU(), nn.Linear(128, 2), nn.Softmax(-1))


class ReinforceTrainer:
    def __init__(self, cfg: ReinforceConfig, device: Optional[torch.device] = None, debug_masks: bool = False):
        _lib.require_cuda()
        if cfg.env_id != "CartPole-v1":
            raise ValueError("the reference's REINFORCE script is written for CartPole-v1 (reinforce.py:26,39,42)")
        self.cfg, self.L = cfg, _lib.lib()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        N = cfg.num_envs
        self.env = VecEnv(cfg.env_id, num_envs=N, seed=cfg.seed, device=self.device)
        self.T = 500                                   # env.spec.max_episode_steps (reinforce.py:53-54)
        torch.manual_seed(cfg.seed)                    # reinforce.py:35-36: the same init draws as the script
        self.agent = make_policy()
        flat = torch.cat([p.detach().reshape(-1) for p in self.agent.parameters()]).to(self.device, torch.float32).contiguous()
        P = int(self.L.drl_reinforce_param_count())
        assert flat.numel() == P
        self.params = flat
        off = 0
        for mod in self.agent:
            if isinstance(mod, nn.Linear):
                for name in ("weight", "bias"):
                    old = getattr(mod, name)
                    setattr(mod, name, nn.Parameter(flat[off:off + old.numel()].view(old.shape)))
                    off += old.numel()
        d, f32 = self.device, torch.float32
        T = self.T
        self.observations = torch.zeros((T + 1, N, 4), dtype=f32, device=d)
        self.actions = torch.zeros((T + 1, N), dtype=torch.uint8, device=d)
        self.rewards = torch.zeros((T + 1, N), dtype=f32, device=d)
        self.dones = torch.zeros((T + 1, N), dtype=torch.uint8, device=d)
        self.zeros = torch.zeros((T + 1, N), dtype=f32, device=d)          # the value plane of the lambda = 1 scan
        self.returns = torch.zeros((T + 1, N), dtype=f32, device=d)
        self._ret_plus_v = torch.zeros((T + 1, N), dtype=f32, device=d)
        self.ep_len = torch.zeros(N, dtype=torch.int32, device=d)
        self.mask_bits = torch.zeros((T, N, 4), dtype=torch.int32, device=d) if debug_masks else None
        self.grad = torch.zeros(P, dtype=f32, device=d)
        self.exp_avg = torch.zeros(P, dtype=f32, device=d)
        self.exp_avg_sq = torch.zeros(P, dtype=f32, device=d)
        self.loss = torch.zeros(1, dtype=f32, device=d)
        self._grad_part = torch.zeros((N, (P + 3) // 4 * 4), dtype=f32, device=d)     # rows padded to 16 bytes
        self._loss_part = torch.zeros(N, dtype=f32, device=d)
        self.net = _lib.NetT(4, 64, 2, 4)              # shape argument of drl_gae (plane strides only)
        self.buf = _lib.RolloutBufT(0, 0, 0, self.zeros.data_ptr(), self.rewards.data_ptr(), self.dones.data_ptr(), 0)
        self.step0 = 1                                  # global step index of the next episode's first step
        self.adam_step = 0
        self.global_step = 0
        self.kernel_launches = 0

    def episodes(self) -> None:
        _lib.check(self.L.drl_reinforce_episodes(C.byref(self.env.struct), self.params.data_ptr(), self.T, self.step0,
                                                 self.observations.data_ptr(), self.actions.data_ptr(), self.rewards.data_ptr(),
                                                 self.dones.data_ptr(), self.ep_len.data_ptr(), _lib.ptr(self.mask_bits),
                                                 C.byref(self.env.log.struct), _lib.stream_ptr()))
        self.kernel_launches += 1

    def compute_returns(self) -> None:
        _lib.check(self.L.drl_gae(C.byref(self.buf), C.byref(self.net), self.T, self.cfg.num_envs, self.cfg.gamma, 1.0,
                                  self.returns.data_ptr(), self._ret_plus_v.data_ptr(), 0, _lib.stream_ptr()))
        self.kernel_launches += 1

    def optimize(self, teacher_masks: Optional[torch.Tensor] = None) -> None:
        N = self.cfg.num_envs
        masks = teacher_masks if teacher_masks is not None else None
        _lib.check(self.L.drl_reinforce_grad(self.params.data_ptr(), self.observations.data_ptr(), self.actions.data_ptr(),
                                             self.returns.data_ptr(), self.ep_len.data_ptr(), _lib.ptr(masks), N, self.cfg.seed,
                                             self.env.env_gid0, self.step0, 1.0 / N, self.grad.data_ptr(), self.loss.data_ptr(),
                                             self._grad_part.data_ptr(), self._loss_part.data_ptr(), _lib.stream_ptr()))
        self.adam_step += 1
        _lib.check(self.L.drl_adam_step(self.params.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                        self.params.numel(), self.adam_step, self.cfg.learning_rate, 0.9, 0.999, 1e-8, _lib.stream_ptr()))
        self.kernel_launches += 3

    def iteration(self) -> None:
        """One pass of the script's outer loop for every env (asynchronous)."""
        self.episodes()
        self.compute_returns()
        self.optimize()
        self.step0 += self.T

    def metrics(self) -> Dict[str, object]:
        n, sum_ret, sum_len, entries = self.env.log.drain()
        self.global_step += int(sum_len)
        return {"episodes": n, "mean_return": (sum_ret / n) if n else float("nan"), "loss": float(self.loss.item()),
                "episode_log": entries}


def train(cfg: ReinforceConfig, quiet: bool = False) -> ReinforceTrainer:
    """The training loop of the reference script; prints its `global_step=..., episodic_return=...` lines (reinforce.py:69)."""
    tr = ReinforceTrainer(cfg)
    for _ in range(cfg.num_episodes):
        tr.iteration()
        m = tr.metrics()
        if not quiet:
            if cfg.num_envs == 1:
                for _step, _env, ret, _len in m["episode_log"]:
                    print(f"global_step={tr.global_step}, episodic_return={ret:.2f}")
            else:
                print(f"global_step={tr.global_step}, episodic_return={m['mean_return']:.2f} (mean of {m['episodes']} episodes)")
    return tr
