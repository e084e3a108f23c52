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

import numpy as np
import torch

from . import _lib

ENV_KINDS = {"CartPole-v1": 0, "Acrobot-v1": 1, "MountainCar-v0": 2}
MAX_EPISODE_STEPS = {"CartPole-v1": 500, "Acrobot-v1": 500, "MountainCar-v0": 200}      # TimeLimit of the gym registrations


ENTRY_DTYPE = np.dtype([("step", "<u8"), ("env", "<u4"), ("ret", "<f4"), ("len", "<i4"), ("pad", "<u4")])     # drl_ep_entry_t
HDR_BYTES = 32        # u32 count | pad | f64 sum_ret | f64 sum_len | pad


class LogRead:
    """A device->host read of the episode log in flight (EpisodeLog.read_async): result() waits for it."""

    def __init__(self, log: "EpisodeLog", host: torch.Tensor, guess: int, event: torch.cuda.Event):
        self.log, self.host, self.guess, self.event = log, host, guess, event

    def result(self):
        """(count, sum_ret, sum_len, entries as a numpy structured array view [step, env, ret, len], entries lost).  Entries
        beyond the speculative prefix were already overwritten when this is called and count as lost (none in the steady
        state: the prefix is twice the previous count)."""
        self.event.synchronize()
        raw = self.host.numpy()
        n = int(raw[0:4].view(np.uint32)[0])
        sums = raw[8:24].view(np.float64)
        k = min(n, self.guess)
        entries = raw[HDR_BYTES:HDR_BYTES + 24 * k].view(ENTRY_DTYPE)
        self.log._guess = min(self.log.cap, 2 * n + 1024)
        return n, float(sums[0]), float(sums[1]), entries, n - k


class EpisodeLog:
    """Finished-episode log written by the kernels (drl_ep_log_t): one device buffer = 32-byte header (count, sums) followed by
    `cap` 24-byte records, so that header + a prefix of the records travel to the host in ONE copy."""

    def __init__(self, cap: int, device):
        self.cap = int(cap)
        self.buf = torch.zeros(HDR_BYTES + 24 * self.cap, dtype=torch.uint8, device=device)
        self._hdr = self.buf[:HDR_BYTES]
        self.count = self.buf[0:4].view(torch.int32)
        self.sums = self.buf[8:24].view(torch.float64)      # sum_ret, sum_len
        self._words = self.buf[HDR_BYTES:].view(torch.int32).view(self.cap, 6)      # step lo, step hi, env, ret bits, len, pad
        self.struct = _lib.EpLogT(self.buf.data_ptr(), self.buf.data_ptr() + 8, self.buf.data_ptr() + 16,
                                  self.buf.data_ptr() + HDR_BYTES, self.cap)
        self._host = [torch.zeros(HDR_BYTES + 24 * self.cap, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self._flip = 0
        self._guess = self.cap        # the first read fetches the whole log once; later reads twice the previous count
        self.last_d2h_bytes = 0

    # typed views of the record fields (tests, single-env gym API)
    @property
    def ret(self) -> torch.Tensor:
        return self._words[:, 3].contiguous().view(torch.float32)

    @property
    def len(self) -> torch.Tensor:
        return self._words[:, 4]

    @property
    def env(self) -> torch.Tensor:
        return self._words[:, 2]

    @property
    def step(self) -> torch.Tensor:
        return self._words[:, 0:2].contiguous().view(torch.int64).reshape(-1)

    def clear(self):
        self._hdr.zero_()

    def read_async(self) -> LogRead:
        """Enqueue (current stream) ONE copy of header + a speculative prefix of the records into pinned memory, then the
        clear of the header; returns the handle whose result() waits for the copy.  Two pinned buffers alternate, so the
        read of update k may be consumed while update k+1 is already running."""
        host = self._host[self._flip]
        self._flip ^= 1
        g = self._guess
        nbytes = HDR_BYTES + 24 * g
        host[:nbytes].copy_(self.buf[:nbytes], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.clear()
        self.last_d2h_bytes = nbytes
        return LogRead(self, host, g, ev)

    def drain_arrays(self):
        """Synchronous D2H read + clear: (count, sum_ret, sum_len, {"step","env","ret","len"} numpy views of a pinned host
        buffer, unsorted, valid until the next-but-one read).  One copy + one synchronisation in the steady state; only when
        more episodes finished than the speculative prefix holds does a second copy fetch the rest (before the clear)."""
        host = self._host[self._flip]
        self._flip ^= 1
        g = self._guess
        host[:HDR_BYTES + 24 * g].copy_(self.buf[:HDR_BYTES + 24 * g], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        raw = host.numpy()
        n = int(raw[0:4].view(np.uint32)[0])
        sums = raw[8:24].view(np.float64)
        sum_ret, sum_len = float(sums[0]), float(sums[1])
        k = min(n, self.cap)
        if k > g:
            host[HDR_BYTES + 24 * g:HDR_BYTES + 24 * k].copy_(self.buf[HDR_BYTES + 24 * g:HDR_BYTES + 24 * k], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        self.clear()
        self.last_d2h_bytes = HDR_BYTES + 24 * max(k, g)
        self._guess = min(self.cap, 2 * n + 1024)
        e = raw[HDR_BYTES:HDR_BYTES + 24 * k].view(ENTRY_DTYPE)
        return n, sum_ret, sum_len, {"step": e["step"], "env": e["env"], "ret": e["ret"], "len": e["len"]}

    def read_header(self):
        """(count, sum_ret, sum_len): one 32-byte D2H copy (synchronises the current stream)."""
        host = self._host[self._flip]
        host[:HDR_BYTES].copy_(self._hdr, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        raw = host.numpy()
        sums = raw[8:24].view(np.float64)
        return int(raw[0:4].view(np.uint32)[0]), float(sums[0]), float(sums[1])

    def drain(self, with_entries: bool = True):
        """D2H read + clear.  Returns (count, sum_ret, sum_len, entries [(step, env, ret, len)] sorted by (step, env))."""
        if not with_entries:
            n, sum_ret, sum_len = self.read_header()
            self.clear()
            return n, sum_ret, sum_len, []
        n, sum_ret, sum_len, a = self.drain_arrays()
        entries = sorted(zip(a["step"].tolist(), a["env"].tolist(), a["ret"].tolist(), a["len"].tolist()))
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
            # lasts >= 8 steps), 24 B per entry, bounded at 16 Mi entries (403 MB: the 1 Mi-env stress config finishes ~12 M
            # episodes per 256-step rollout); beyond it entries are dropped and COUNTED (metrics()["episodes_dropped"])
            log_capacity = min(max(1 << 16, N * 32), 1 << 24)
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
