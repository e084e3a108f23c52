"""Single-env CPU port of the reference training loop -- CPU ORACLE, test infrastructure only.

A restatement (not a copy) of qgallouedec/deep_rl `deep_rl/ppo.py:79-197` with the same stock
PyTorch ops at the same granularity (one env, 1-D tensors, torch CPU mt19937 for sampling, numpy
legacy MT19937 for the minibatch permutation, Adam over twelve parameters), so that

  * tests can pin it against outputs of the UNMODIFIED reference (tests/golden/ref_ppo_seed1.npz),
  * bench.py can time "the reference's own CPU implementation" on a box where /root/reference and
    gym do not exist (cpu_baseline.kind == "port", and `bench.py --impl reference`).

The env is the gym-0.21 shim's CartPole/Acrobot (oracle/gym_shim), i.e. the C oracle physics.
"""
from __future__ import annotations

import os
import sys
import time
from dataclasses import dataclass, field
from typing import Callable, List, Optional

import numpy as np
import torch
from torch import nn

_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gym_shim")


def _shim_gym():
    if _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)
    import gym  # the shim, never a real gym (none is installed)
    return gym


@dataclass
class PortConfig:
    env_id: str = "CartPole-v1"
    total_timesteps: int = 20_000
    num_steps: int = 128
    update_epochs: int = 4
    gamma: float = 0.99
    gae_lambda: float = 0.95
    learning_rate: float = 2.5e-4
    clip_coef: float = 0.2
    ent_coef: float = 0.01
    vf_coef: float = 0.5
    max_grad_norm: float = 0.5
    seed: int = 1
    hidden: int = 64

    @property
    def num_updates(self) -> int:
        return self.total_timesteps // self.num_steps

    @property
    def minibatch_size(self) -> int:
        return self.num_steps // 4


def _ortho(layer: nn.Linear, gain: float) -> nn.Linear:
    nn.init.orthogonal_(layer.weight, gain)
    nn.init.constant_(layer.bias, 0.0)
    return layer


class PortAgent(nn.Module):
    """Two separate tanh MLPs, ppo.py:31-59."""

    def __init__(self, obs_dim: int, num_actions: int, hidden: int = 64):
        super().__init__()
        g = float(np.sqrt(2))
        self.actor = nn.Sequential(_ortho(nn.Linear(obs_dim, hidden), g), nn.Tanh(),
                                   _ortho(nn.Linear(hidden, hidden), g), nn.Tanh(),
                                   _ortho(nn.Linear(hidden, num_actions), 0.01))
        self.critic = nn.Sequential(_ortho(nn.Linear(obs_dim, hidden), g), nn.Tanh(),
                                    _ortho(nn.Linear(hidden, hidden), g), nn.Tanh(),
                                    _ortho(nn.Linear(hidden, 1), 1.0))

    def value(self, obs):
        return self.critic(obs).squeeze(-1)

    def dist(self, obs):
        return torch.distributions.Categorical(logits=self.actor(obs))


@dataclass
class PortTrace:
    episodes: List[tuple] = field(default_factory=list)      # (global_step, return)
    seconds: float = 0.0
    env_steps: int = 0


def run(cfg: PortConfig = PortConfig(), max_updates: Optional[int] = None,
        on_update: Optional[Callable[[int, dict], None]] = None, quiet: bool = True) -> PortTrace:
    gym = _shim_gym()
    env = gym.wrappers.RecordEpisodeStatistics(gym.make(cfg.env_id))
    env.seed(cfg.seed)
    np.random.seed(cfg.seed)
    torch.manual_seed(cfg.seed)

    obs_dim = int(np.prod(env.observation_space.shape))
    agent = PortAgent(obs_dim, env.action_space.n, cfg.hidden)
    opt = torch.optim.Adam(agent.parameters(), lr=cfg.learning_rate, eps=1e-5)

    T = cfg.num_steps
    buf_obs = torch.zeros((T + 1, obs_dim))
    buf_val = torch.zeros(T + 1)
    buf_act = torch.zeros(T + 1, dtype=torch.long)
    buf_logp = torch.zeros(T + 1)
    buf_rew = torch.zeros(T + 1)
    buf_done = torch.zeros(T + 1)

    trace = PortTrace()
    ob = torch.tensor(env.reset())
    gstep = 0
    n_updates = cfg.num_updates if max_updates is None else min(max_updates, cfg.num_updates)
    t_start = time.perf_counter()
    for upd in range(n_updates):
        opt.param_groups[0]["lr"] = (1.0 - upd / cfg.num_updates) * cfg.learning_rate

        # ---- rollout (ppo.py:110-141) ----
        buf_obs[0] = ob
        with torch.no_grad():
            buf_val[0] = agent.value(ob)
        for t in range(T):
            with torch.no_grad():
                pi = agent.dist(buf_obs[t])
                a = pi.sample()
                buf_act[t] = a
                buf_logp[t] = pi.log_prob(a)
            o_np, r, d, info = env.step(a.cpu().numpy())
            ob = torch.tensor(o_np)
            if d:
                ob = torch.tensor(env.reset())
                trace.episodes.append((gstep, float(info["episode"]["r"])))
                if not quiet:
                    print(f"global_step={gstep}, episodic_return={info['episode']['r']:.2f}")
            gstep += 1
            buf_obs[t + 1] = ob
            with torch.no_grad():
                buf_val[t + 1] = agent.value(buf_obs[t + 1])
            buf_rew[t + 1] = r
            buf_done[t + 1] = d

        # ---- GAE (ppo.py:144-151) ----
        adv = torch.zeros_like(buf_rew)
        last = 0
        for t in reversed(range(T)):
            adv[t] = buf_rew[t + 1] + cfg.gamma * (1.0 - buf_done[t + 1]) * (buf_val[t + 1] + cfg.gae_lambda * last) - buf_val[t]
            last = adv[t]
        ret = adv + buf_val

        # ---- minibatch SGD (ppo.py:154-192) ----
        mbs = cfg.minibatch_size
        for _ in range(cfg.update_epochs):
            order = np.random.permutation(T)
            for lo in range(0, T, mbs):
                sel = order[lo:lo + mbs]
                pi = agent.dist(buf_obs[sel])
                A_mb = adv[sel]
                A_mb = (A_mb - torch.mean(A_mb)) / (torch.std(A_mb) + 1e-8)
                ratio = torch.exp(pi.log_prob(buf_act[sel]) - buf_logp[sel])
                pg = torch.mean(torch.max(-A_mb * ratio, -A_mb * torch.clamp(ratio, 1 - cfg.clip_coef, 1 + cfg.clip_coef)))
                ent = torch.mean(pi.entropy())
                v_new = agent.value(buf_obs[sel])
                v_clip = buf_val[sel] + torch.clamp(v_new - buf_val[sel], -cfg.clip_coef, cfg.clip_coef)
                v_loss = 0.5 * torch.mean(torch.max((v_new - ret[sel]) ** 2, (v_clip - ret[sel]) ** 2))
                loss = pg - cfg.ent_coef * ent + v_loss * cfg.vf_coef
                opt.zero_grad()
                loss.backward()
                nn.utils.clip_grad_norm_(agent.parameters(), cfg.max_grad_norm)
                opt.step()

        if on_update is not None:
            on_update(upd, dict(obs=buf_obs, val=buf_val, act=buf_act, logp=buf_logp, rew=buf_rew, done=buf_done,
                                adv=adv, ret=ret, agent=agent))
    trace.seconds = time.perf_counter() - t_start
    trace.env_steps = gstep
    env.close()
    return trace


def time_port(seconds_budget: float = 15.0, cfg: PortConfig = PortConfig()) -> dict:
    """Bounded CPU-baseline sample: run whole updates of the port until ~seconds_budget is spent
    (after a one-update warm-up that absorbs torch's lazy imports).  Single-threaded, like the
    reference."""
    run(cfg, max_updates=1)
    t0 = time.perf_counter()
    probe = run(cfg, max_updates=2)
    per_update = (time.perf_counter() - t0) / 2
    n = int(max(2, min(cfg.num_updates, seconds_budget / max(per_update, 1e-6))))
    tr = run(cfg, max_updates=n)
    return {"env_steps": tr.env_steps, "seconds": tr.seconds, "steps_per_s": tr.env_steps / tr.seconds,
            "updates": n, "probe_steps": probe.env_steps}
