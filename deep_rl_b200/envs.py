"""Device-resident batched CartPole-v1 / Acrobot-v1 with the gym-0.21 4-tuple API.

Replaces `gym.wrappers.RecordEpisodeStatistics(gym.make(env_id))` + `TorchWrapper`
(deep_rl/ppo.py:10-22,79-80) and the manual auto-reset of ppo.py:127-129 for `num_envs`
environments.  With num_envs == 1 a step returns what the reference's wrapper chain returns (obs
tensor, reward, done, info with info["episode"]["r"|"l"] on episode end), except that the reset
after `done` has already happened inside the kernel (SyncVectorEnv semantics): the returned
observation is the first observation of the next episode, so the caller must NOT call reset() again.
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace
from typing import Optional

import torch

from . import _lib

ENV_KINDS = {"CartPole-v1": 0, "Acrobot-v1": 1, "MountainCar-v0": 2}
MAX_EPISODE_STEPS = {"CartPole-v1": 500, "Acrobot-v1": 500, "MountainCar-v0": 200}      # TimeLimit of the gym registrations


class EpisodeLog:
    """Finished-episode log written by the kernels (drl_ep_log_t)."""

    def __init__(self, cap: int, device):
        self.cap = int(cap)
        # count (int32 @0) and the two float64 sums (@8, @16) share one 24-byte buffer: one fill clears them,
        # one 24-byte device->host copy reads them.
        self._hdr = torch.zeros(24, dtype=torch.uint8, device=device)
        self.count = self._hdr[0:4].view(torch.int32)
        self.sums = self._hdr[8:24].view(torch.float64)      # sum_ret, sum_len
        self._hdr_host = torch.zeros(24, dtype=torch.uint8).pin_memory()
        self.ret = torch.zeros(self.cap, dtype=torch.float32, device=device)
        self.len = torch.zeros(self.cap, dtype=torch.int32, device=device)
        self.env = torch.zeros(self.cap, dtype=torch.int32, device=device)
        self.step = torch.zeros(self.cap, dtype=torch.int64, device=device)
        self.struct = _lib.EpLogT(self.count.data_ptr(), self.sums.data_ptr(), self.sums.data_ptr() + 8,
                                  self.ret.data_ptr(), self.len.data_ptr(), self.env.data_ptr(), self.step.data_ptr(),
                                  self.cap)

    def read_header(self):
        """(count, sum_ret, sum_len): one 24-byte D2H copy (synchronises the current stream)."""
        self._hdr_host.copy_(self._hdr, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        n = int(self._hdr_host[0:4].view(torch.int32).item())
        s = self._hdr_host[8:24].view(torch.float64).tolist()
        return n, s[0], s[1]

    def clear(self):
        self._hdr.zero_()

    def drain_arrays(self):
        """D2H read + clear without Python-object conversion: (count, sum_ret, sum_len, {"step","env","ret","len"} numpy views
        of pinned host buffers, unsorted, valid until the next drain).  ONE synchronisation in the steady state: the header
        and a speculative prefix of the entries (1.25 x the previous count) are copied together; only when more episodes
        finished than guessed does a second copy + synchronisation fetch the rest."""
        if not hasattr(self, "_host"):
            self._host = {"step": torch.zeros(self.cap, dtype=torch.int64).pin_memory(), "env": torch.zeros(self.cap, dtype=torch.int32).pin_memory(),
                          "ret": torch.zeros(self.cap, dtype=torch.float32).pin_memory(), "len": torch.zeros(self.cap, dtype=torch.int32).pin_memory()}
            self._guess = 0
        srcs = (("step", self.step), ("env", self.env), ("ret", self.ret), ("len", self.len))
        g = min(self.cap, self._guess)
        self._hdr_host.copy_(self._hdr, non_blocking=True)
        if g:
            for name, src in srcs:
                self._host[name][:g].copy_(src[:g], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        n = int(self._hdr_host[0:4].view(torch.int32).item())
        s = self._hdr_host[8:24].view(torch.float64).tolist()
        k = min(n, self.cap)
        if k > g:
            for name, src in srcs:
                self._host[name][g:k].copy_(src[g:k], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        self.clear()
        self.last_d2h_bytes = 24 + 20 * max(k, g)
        self._guess = k + k // 4 + 64
        return n, s[0], s[1], {name: buf[:k].numpy() for name, buf in self._host.items()}

    def drain(self, with_entries: bool = True):
        """D2H read + clear.  Returns (count, sum_ret, sum_len, entries sorted by (step, env))."""
        n, sum_ret, sum_len = self.read_header()
        k = min(n, self.cap)
        entries = []
        if k and with_entries:
            st, ev = self.step[:k].tolist(), self.env[:k].tolist()
            rt, ln = self.ret[:k].tolist(), self.len[:k].tolist()
            entries = sorted(zip(st, ev, rt, ln))
        self.clear()
        return n, sum_ret, sum_len, entries


class VecEnv:
    """`num_envs` independent environments stepping in one kernel launch."""

    def __init__(self, env_id: str = "CartPole-v1", num_envs: int = 1, seed: int = 1, env_gid0: int = 0,
                 device: Optional[torch.device] = None, log_capacity: Optional[int] = None):
        if env_id not in ENV_KINDS:
            raise ValueError(f"unsupported env_id {env_id!r}; supported: {sorted(ENV_KINDS)}")
        _lib.require_cuda()
        self.L = _lib.lib()
        self.env_id, self.kind = env_id, ENV_KINDS[env_id]
        self.num_envs, self.env_gid0 = int(num_envs), int(env_gid0)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.obs_dim = self.L.drl_env_obs_dim(self.kind)
        self.num_actions = self.L.drl_env_num_actions(self.kind)
        self.obs_stride = self.L.drl_env_obs_stride(self.kind)
        self.observation_space = SimpleNamespace(shape=(self.obs_dim,), dtype=torch.float32)
        self.action_space = SimpleNamespace(n=self.num_actions, shape=(), dtype=torch.int64)
        N = self.num_envs
        self.state = torch.zeros((4, N), dtype=torch.float64, device=self.device)
        self.elapsed = torch.zeros(N, dtype=torch.int32, device=self.device)
        self.ep_ret = torch.zeros(N, dtype=torch.float32, device=self.device)
        self.ep_len = torch.zeros(N, dtype=torch.int32, device=self.device)
        if log_capacity is None:
            # enough for every episode that can finish between two drains of a 256-step rollout (an untrained CartPole policy
            # lasts >= 8 steps), 24 B per entry, bounded at 4 Mi entries; beyond it entries are dropped and COUNTED
            # (metrics()["episodes_dropped"])
            log_capacity = min(max(1 << 16, N * 32), 1 << 22)
        self.log = EpisodeLog(log_capacity, self.device)
        self.step_count = 0          # global step index fed to the Philox counter
        self._seed = int(seed)
        self._obs = torch.zeros((N, self.obs_stride), dtype=torch.float32, device=self.device)
        self._rew = torch.zeros(N, dtype=torch.float32, device=self.device)
        self._done = torch.zeros(N, dtype=torch.uint8, device=self.device)
        self._struct = None

    # -- gym surface ------------------------------------------------------------------------
    def seed(self, seed: int):
        self._seed = int(seed)
        self._struct = None
        return [seed]

    @property
    def struct(self) -> _lib.EnvT:
        if self._struct is None:
            self._struct = _lib.EnvT(self.kind, self.num_envs, self._seed, self.env_gid0, MAX_EPISODE_STEPS[self.env_id],
                                     self.state.data_ptr(), self.elapsed.data_ptr(), self.ep_ret.data_ptr(),
                                     self.ep_len.data_ptr())
        return self._struct

    def _view(self, obs: torch.Tensor) -> torch.Tensor:
        o = obs[:, : self.obs_dim]
        return o[0] if self.num_envs == 1 else o

    def reset(self) -> torch.Tensor:
        _lib.check(self.L.drl_env_reset(C.byref(self.struct), self._obs.data_ptr(), _lib.stream_ptr()))
        self.step_count = 0
        return self._view(self._obs.clone())

    def observe(self) -> torch.Tensor:
        _lib.check(self.L.drl_env_observe(C.byref(self.struct), self._obs.data_ptr(), _lib.stream_ptr()))
        return self._view(self._obs.clone())

    def observe_padded(self) -> torch.Tensor:
        _lib.check(self.L.drl_env_observe(C.byref(self.struct), self._obs.data_ptr(), _lib.stream_ptr()))
        return self._obs

    def set_state(self, state: torch.Tensor) -> torch.Tensor:
        """Inject float64 states [N, 4] (parity tests teacher-force the oracle's states)."""
        self.state.copy_(torch.as_tensor(state, dtype=torch.float64).reshape(self.num_envs, 4).t())
        return self.observe()

    def get_state(self) -> torch.Tensor:
        return self.state.t().contiguous()

    def step(self, action: torch.Tensor):
        a = torch.as_tensor(action, device=self.device).reshape(self.num_envs).to(torch.int32).contiguous()
        before = int(self.log.count.item()) if self.num_envs == 1 else None
        _lib.check(self.L.drl_env_step(C.byref(self.struct), self.step_count, a.data_ptr(), self._obs.data_ptr(),
                                       self._rew.data_ptr(), self._done.data_ptr(), C.byref(self.log.struct),
                                       _lib.stream_ptr()))
        self.step_count += 1
        obs, rew, done = self._view(self._obs.clone()), self._rew.clone(), self._done.to(torch.bool)
        info = {}
        if self.num_envs == 1:
            d = bool(done.item())
            if d:
                k = min(before, self.log.cap - 1)
                info["episode"] = {"r": float(self.log.ret[k].item()), "l": int(self.log.len[k].item())}
            return obs, float(rew.item()), d, info
        return obs, rew, done, info

    def close(self):
        pass


def make(env_id: str = "CartPole-v1", num_envs: int = 1, seed: int = 1, **kw) -> VecEnv:
    """Counterpart of `gym.wrappers.RecordEpisodeStatistics(gym.make(env_id))` wrapped in `TorchWrapper`."""
    return VecEnv(env_id, num_envs=num_envs, seed=seed, **kw)
