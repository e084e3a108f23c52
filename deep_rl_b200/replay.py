"""Device-resident replay buffer: the storage, index samplers and batch gather of the reference's off-policy scripts.

    deep_rl/dqn.py:73-77,96-109   storage `observations[T+1, O]`, `actions[T+1]`, `rewards[T+1]`, `terminated[T+1]` with the one-slot
                                   shift (reward / terminated / next observation of transition i live at i + 1)
    deep_rl/dqn.py:116-122         batch_inds = np.random.randint(global_step, size=batch_size) and the five gathers
    deep_rl/per.py:78,104,127-146  priorities, torch.multinomial draw, probabilities, priority update

Everything stays in HBM as flat struct-of-arrays (observation rows padded to 4 / 8 floats = 16 / 32 bytes, so one transition's two
observation rows are 1-2 DRAM sectors); the kernels are behind the C ABI (drl_replay_*).  Index draws follow the build's Philox
contract, not numpy's / torch's MT19937 streams.  No CPU fallback.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _lib


class ReplayBuffer:
    def __init__(self, capacity: int, obs_dim: int, seed: int = 1, prioritized: bool = False, alpha: float = 0.6,
                 device: Optional[torch.device] = None):
        _lib.require_cuda()
        self.L = _lib.lib()
        self.capacity, self.obs_dim = int(capacity), int(obs_dim)
        if obs_dim > 8:
            raise ValueError(f"obs_dim={obs_dim}: observation rows of up to 8 floats are supported")
        self.obs_stride = 4 if obs_dim <= 4 else 8
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        n = self.capacity + 1
        d = self.device
        self.observations = torch.zeros((n, self.obs_stride), dtype=torch.float32, device=d)
        self.actions = torch.zeros(n, dtype=torch.int32, device=d)
        self.rewards = torch.zeros(n, dtype=torch.float32, device=d)
        self.terminated = torch.zeros(n, dtype=torch.uint8, device=d)
        self.prioritized, self.alpha = bool(prioritized), float(alpha)
        self.priorities = torch.zeros(n, dtype=torch.float32, device=d) if prioritized else None
        self.max_priority = torch.full((1,), 1e-2, dtype=torch.float32, device=d)        # per.py:81
        self.seed, self.draws = int(seed), 0
        self.size = 0                                                                     # = the script's global_step
        self._scratch = torch.zeros(int(self.L.drl_replay_scratch_bytes(n)), dtype=torch.uint8, device=d) if prioritized else None

    # -- storage (dqn.py:96-109, per.py:101-117) -----------------------------------------------------
    def store_initial(self, observation: torch.Tensor) -> None:
        self.observations[0, : self.obs_dim] = observation

    def store(self, action, next_observation: torch.Tensor, reward, terminated) -> None:
        """actions[t] = action; (per.py) priorities[t] = max_priority; then t += 1 and observations / rewards / terminated[t]."""
        t = self.size
        if t >= self.capacity:
            raise IndexError("replay buffer full (the reference allocates total_timesteps + 1 slots and never wraps)")
        self.actions[t] = int(action)
        if self.prioritized:
            self.priorities[t] = self.max_priority[0]
        t += 1
        self.observations[t, : self.obs_dim] = next_observation
        self.rewards[t] = float(reward)
        self.terminated[t] = int(bool(terminated))
        self.size = t

    # -- sampling ------------------------------------------------------------------------------------
    def sample_indices(self, batch_size: int) -> torch.Tensor:
        """dqn.py:116 (uniform over [0, size)) or per.py:129 (proportional to the raw priorities)."""
        idx = torch.empty(batch_size, dtype=torch.int32, device=self.device)
        self._prob = None
        if self.prioritized:
            self._prob = torch.empty(batch_size, dtype=torch.float32, device=self.device)
            _lib.check(self.L.drl_replay_sample_priority(self.priorities.data_ptr(), self.size, self.alpha, batch_size, self.seed, self.draws,
                                                         idx.data_ptr(), self._prob.data_ptr(), self._scratch.data_ptr(),
                                                         self._scratch.numel(), _lib.stream_ptr()))
        else:
            _lib.check(self.L.drl_replay_sample_uniform(idx.data_ptr(), batch_size, self.size, self.seed, self.draws, _lib.stream_ptr()))
        self.draws += 1
        return idx

    def gather(self, batch_inds: torch.Tensor) -> Dict[str, torch.Tensor]:
        """dqn.py:118-122: b_observations, b_actions, b_next_observations, b_rewards, b_terminated."""
        idx = batch_inds.to(self.device, torch.int32).contiguous()
        b = idx.numel()
        d = self.device
        out = {"observations": torch.empty((b, self.obs_stride), dtype=torch.float32, device=d),
               "next_observations": torch.empty((b, self.obs_stride), dtype=torch.float32, device=d),
               "actions": torch.empty(b, dtype=torch.int32, device=d), "rewards": torch.empty(b, dtype=torch.float32, device=d),
               "terminated": torch.empty(b, dtype=torch.uint8, device=d)}
        _lib.check(self.L.drl_replay_gather(self.observations.data_ptr(), self.actions.data_ptr(), self.rewards.data_ptr(),
                                            self.terminated.data_ptr(), idx.data_ptr(), b, self.obs_stride,
                                            out["observations"].data_ptr(), out["next_observations"].data_ptr(),
                                            out["actions"].data_ptr(), out["rewards"].data_ptr(), out["terminated"].data_ptr(),
                                            _lib.stream_ptr()))
        out["observations"] = out["observations"][:, : self.obs_dim]
        out["next_observations"] = out["next_observations"][:, : self.obs_dim]
        return out

    def sample(self, batch_size: int) -> Dict[str, torch.Tensor]:
        idx = self.sample_indices(batch_size)
        batch = self.gather(idx)
        batch["batch_inds"] = idx
        if self.prioritized:
            batch["probabilities"] = self._prob           # per.py:128,131
        return batch

    def importance_weights(self, probabilities: torch.Tensor, beta: float) -> torch.Tensor:
        """per.py:149-150: (size * p) ** -beta, normalised by the maximum."""
        w = (self.size * probabilities) ** -beta
        return w / torch.max(w)

    def update_priorities(self, batch_inds: torch.Tensor, td_errors: torch.Tensor) -> None:
        """per.py:144-146."""
        idx = batch_inds.to(self.device, torch.int32).contiguous()
        td = td_errors.detach().to(self.device, torch.float32).contiguous()
        _lib.check(self.L.drl_replay_update_priorities(self.priorities.data_ptr(), idx.data_ptr(), td.data_ptr(), idx.numel(),
                                                       self.max_priority.data_ptr(), _lib.stream_ptr()))
