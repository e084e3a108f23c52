"""REINFORCE oracle -- CPU ORACLE, test infrastructure only (never imported by deep_rl_b200/).

Restates deep_rl/reinforce.py:38-77 with stock PyTorch CPU ops: the policy forward with an explicit dropout keep mask (the
reference's mask comes from torch's CPU stream inside nn.Dropout; the build draws it from Philox -- SURVEY.md D4 -- so parity is
checked with the mask given), the reward-to-go exactly as the script accumulates it (reinforce.py:67), the normalised loss
(reinforce.py:71-74) with autograd, and the Adam step (reinforce.py:45,77).  Pinned against tests/golden/ref_reinforce.npz, which is
generated from the unmodified script.
"""
from __future__ import annotations

import numpy as np
import torch

P = 4 * 128 + 128 + 2 * 128 + 2
LOG_STD_MIN = -5


def split(flat: torch.Tensor):
    w1, b1 = flat[:512].view(128, 4), flat[512:640]
    w2, b2 = flat[640:896].view(2, 128), flat[896:898]
    return w1, b1, w2, b2


def probs(flat: torch.Tensor, obs: torch.Tensor, keep: torch.Tensor) -> torch.Tensor:
    """agent(obs) with the dropout keep mask given: Linear, Dropout(p=0.6) [x * keep / 0.4], ReLU, Linear, Softmax."""
    w1, b1, w2, b2 = split(flat)
    z = torch.nn.functional.linear(obs, w1, b1)
    h = torch.relu(z * keep.to(z.dtype) / (1.0 - 0.6))
    return torch.softmax(torch.nn.functional.linear(h, w2, b2), -1)


def reward_to_go(rewards: np.ndarray, gamma: float) -> torch.Tensor:
    """reinforce.py:67, literally: after every step, returns[:step] += gamma ** flip(arange(step)) * reward."""
    n = len(rewards)
    ret = torch.zeros(n + 1)
    for step in range(1, n + 1):
        ret[:step] += gamma ** torch.flip(torch.arange(step), (0,)) * float(rewards[step - 1])
    return ret[:n]


def episode_loss_and_grad(flat_np, obs, act, keep, rewards, gamma: float):
    """(policy_loss, flat gradient, normalised returns, log_probs) of one episode, reinforce.py:71-76."""
    flat = torch.tensor(np.asarray(flat_np, dtype=np.float32), requires_grad=True)
    obs_t = torch.as_tensor(np.asarray(obs, dtype=np.float32))
    p = probs(flat, obs_t, torch.as_tensor(np.asarray(keep)))
    logp = torch.distributions.Categorical(p).log_prob(torch.as_tensor(np.asarray(act, dtype=np.int64)))
    ret = reward_to_go(rewards, gamma)
    b_returns = (ret - ret.mean()) / (ret.std() + np.exp(LOG_STD_MIN))
    loss = torch.sum(-logp * b_returns)
    loss.backward()
    return float(loss.detach()), flat.grad.numpy().copy(), b_returns.numpy().copy(), logp.detach().numpy().copy()


def adam(flat_np, grad_np, m_np, v_np, step: int, lr: float = 1e-2):
    p = torch.nn.Parameter(torch.tensor(np.asarray(flat_np, dtype=np.float32)))
    p.grad = torch.tensor(np.asarray(grad_np, dtype=np.float32))
    opt = torch.optim.Adam([p], lr=lr)
    st = opt.state[p]
    st["step"] = torch.tensor(float(step - 1))
    st["exp_avg"] = torch.tensor(np.asarray(m_np, dtype=np.float32))
    st["exp_avg_sq"] = torch.tensor(np.asarray(v_np, dtype=np.float32))
    opt.step()
    return p.detach().numpy().copy(), st["exp_avg"].numpy().copy(), st["exp_avg_sq"].numpy().copy()


def unpack_mask_bits(words: np.ndarray) -> np.ndarray:
    """[..., 4] uint32 words of drl_reinforce_episodes' mask plane -> [..., 128] bool keep flags (unit u = bit u % 32 of word u // 32)."""
    w = np.asarray(words).astype(np.uint32)
    bits = (w[..., :, None] >> np.arange(32, dtype=np.uint32)) & 1
    return bits.reshape(*w.shape[:-1], 128).astype(bool)


def pack_mask_bits(keep: np.ndarray) -> np.ndarray:
    k = np.asarray(keep).astype(np.uint32).reshape(*np.asarray(keep).shape[:-1], 4, 32)
    return (k << np.arange(32, dtype=np.uint32)).sum(-1).astype(np.uint32)
