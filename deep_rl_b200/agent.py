"""ActorCritic with the reference's surface (deep_rl/ppo.py:25-59) over a flat device parameter buffer.

`.actor` / `.critic` are ordinary `nn.Sequential(Linear, Tanh, Linear, Tanh, Linear)` modules whose
weights and biases are views into one contiguous fp32 CUDA tensor (`flat_params`, state_dict order), so
`state_dict()` / `load_state_dict()` keep working while the CUDA kernels read and update the flat
buffer (and its packed kernel layout) directly.  `get_value`, `get_action_distribution`, `get_action`
keep the reference signatures; `get_action_and_value` is the CleanRL-style alias.
All four run the forward pass through the C-ABI CUDA kernels, not through torch.nn.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch
from torch import Tensor, nn
from torch.distributions import Categorical, Distribution

from . import _lib

PARAM_NAMES = [
    "actor.0.weight", "actor.0.bias", "actor.2.weight", "actor.2.bias", "actor.4.weight", "actor.4.bias",
    "critic.0.weight", "critic.0.bias", "critic.2.weight", "critic.2.bias", "critic.4.weight", "critic.4.bias",
]


def layer_init(layer: nn.Linear, std: float = float(np.sqrt(2)), bias_const: float = 0.0) -> nn.Linear:
    """Orthogonal weight / constant bias, as ppo.py:25-28."""
    torch.nn.init.orthogonal_(layer.weight, std)
    torch.nn.init.constant_(layer.bias, bias_const)
    return layer


class ActorCritic(nn.Module):
    def __init__(self, env, hidden: int = 64, device: Optional[torch.device] = None, sample_seed: int = 1):
        super().__init__()
        _lib.require_cuda()
        self.L = _lib.lib()
        obs_dim = int(np.array(env.observation_space.shape).prod())
        n_act = int(env.action_space.n)
        obs_stride = int(getattr(env, "obs_stride", 4 if obs_dim <= 4 else 8))
        self.net = _lib.NetT(obs_dim, int(hidden), n_act, obs_stride)
        self.obs_dim, self.num_actions, self.obs_stride, self.hidden = obs_dim, n_act, obs_stride, int(hidden)
        P = self.L.drl_param_count(C.byref(self.net))
        if P < 0:
            _lib.check(int(P))
        self.device = torch.device(device if device is not None else getattr(env, "device", f"cuda:{torch.cuda.current_device()}"))

        # Same construction order as the reference => same draws from the torch CPU stream.
        self.actor = nn.Sequential(
            layer_init(nn.Linear(obs_dim, hidden)), nn.Tanh(),
            layer_init(nn.Linear(hidden, hidden)), nn.Tanh(),
            layer_init(nn.Linear(hidden, n_act), std=0.01),
        )
        self.critic = nn.Sequential(
            layer_init(nn.Linear(obs_dim, hidden)), nn.Tanh(),
            layer_init(nn.Linear(hidden, hidden)), nn.Tanh(),
            layer_init(nn.Linear(hidden, 1), std=1.0),
        )
        flat = torch.cat([p.detach().reshape(-1) for p in self.parameters()]).to(self.device, torch.float32).contiguous()
        assert flat.numel() == P, (flat.numel(), P)
        self.flat_params = flat
        off = 0
        for mod in list(self.actor) + list(self.critic):
            if isinstance(mod, nn.Linear):
                for name in ("weight", "bias"):
                    old = getattr(mod, name)
                    n = old.numel()
                    setattr(mod, name, nn.Parameter(flat[off:off + n].view(old.shape), requires_grad=True))
                    off += n
        self.packed = torch.zeros(int(self.L.drl_packed_count(C.byref(self.net))), dtype=torch.float32, device=self.device)
        self._packed_version = None
        self._sample_seed = int(sample_seed)
        self._sample_calls = 0
        self.sync()

    # -- packed kernel layout ---------------------------------------------------------------
    def sync(self) -> None:
        """Refresh the packed kernel layout from flat_params (after load_state_dict or any torch-side edit)."""
        _lib.check(self.L.drl_pack_params(C.byref(self.net), self.flat_params.data_ptr(), self.packed.data_ptr(),
                                          _lib.stream_ptr()))
        self._packed_version = self.flat_params._version

    def mark_packed_current(self) -> None:
        """Called by the trainer after a kernel refreshed `packed` together with `flat_params`."""
        self._packed_version = self.flat_params._version

    def _ensure_packed(self) -> None:
        if self._packed_version != self.flat_params._version:
            self.sync()

    # -- forward through the CUDA kernels ---------------------------------------------------
    def _forward(self, observation: Tensor) -> Tuple[Tensor, Tensor, tuple]:
        self._ensure_packed()
        obs = torch.as_tensor(observation, dtype=torch.float32, device=self.device)
        if obs.shape[-1] not in (self.obs_dim, self.obs_stride):
            raise ValueError(f"observation last dim {obs.shape[-1]} != {self.obs_dim}")
        lead = obs.shape[:-1]
        x = obs.reshape(-1, obs.shape[-1])
        if x.shape[-1] != self.obs_stride:
            x = torch.nn.functional.pad(x, (0, self.obs_stride - x.shape[-1]))
        x = x.contiguous()
        n = x.shape[0]
        logits = torch.empty((n, self.num_actions), dtype=torch.float32, device=self.device)
        value = torch.empty(n, dtype=torch.float32, device=self.device)
        _lib.check(self.L.drl_policy_forward(C.byref(self.net), self.packed.data_ptr(), x.data_ptr(), n,
                                             logits.data_ptr(), value.data_ptr(), _lib.stream_ptr()))
        return logits.reshape(*lead, self.num_actions), value.reshape(lead), lead

    @torch.no_grad()
    def get_value(self, observation: Tensor) -> Tensor:
        return self._forward(observation)[1]

    @torch.no_grad()
    def get_action_distribution(self, observation: Tensor) -> Distribution:
        return Categorical(logits=self._forward(observation)[0])

    def _sample(self, logits: Tensor, env_gid0: int = 0, step: Optional[int] = None) -> Tuple[Tensor, Tensor]:
        lead = logits.shape[:-1]
        lg = logits.reshape(-1, self.num_actions).contiguous()
        n = lg.shape[0]
        act = torch.empty(n, dtype=torch.int32, device=self.device)
        logp = torch.empty(n, dtype=torch.float32, device=self.device)
        if step is None:
            step = self._sample_calls
            self._sample_calls += 1
        _lib.check(self.L.drl_sample(lg.data_ptr(), n, self.num_actions, self._sample_seed, env_gid0, step,
                                     act.data_ptr(), logp.data_ptr(), _lib.stream_ptr()))
        return act.to(torch.int64).reshape(lead), logp.reshape(lead)

    @torch.no_grad()
    def get_action(self, observation: Tensor) -> Tuple[Tensor, Tensor]:
        """(action, log_prob); the draw uses the Philox stream (sample_seed; row index, call counter)."""
        logits, _, _ = self._forward(observation)
        return self._sample(logits)

    @torch.no_grad()
    def get_action_and_value(self, observation: Tensor, action: Optional[Tensor] = None):
        """(action, log_prob, entropy, value) composed from the three reference methods."""
        logits, value, _ = self._forward(observation)
        dist = Categorical(logits=logits)
        if action is None:
            action, log_prob = self._sample(logits)
        else:
            log_prob = dist.log_prob(action)
        return action, log_prob, dist.entropy(), value
