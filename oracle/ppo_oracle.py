"""PPO math oracle -- CPU ORACLE, test infrastructure only (never imported by deep_rl_b200/).

Restates, with stock PyTorch CPU fp32 ops, the numerics of the reference hot path
(qgallouedec/deep_rl `deep_rl/ppo.py`): model forward (ppo.py:34-54), rollout order of
operations (ppo.py:113-141), loss (ppo.py:166-187), clip + Adam (ppo.py:189-192, 90, 107-108).
GAE, env dynamics, the Philox sampler and the permutation live in drl_oracle.c.

Pinned against the unmodified reference script through tests/golden/ref_ppo_seed1.npz
(tests/test_oracle_golden.py); env dynamics remain "parity unpinned" vs real gym.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np
import torch

from . import clib

# ----------------------------------------------------------------------------------------------
# Flat parameter layout: the twelve tensors of ActorCritic (ppo.py:34-47) in state_dict order,
# each stored [out, in] row-major like nn.Linear.
# ----------------------------------------------------------------------------------------------
PARAM_NAMES = [
    "actor.0.weight", "actor.0.bias", "actor.2.weight", "actor.2.bias", "actor.4.weight", "actor.4.bias",
    "critic.0.weight", "critic.0.bias", "critic.2.weight", "critic.2.bias", "critic.4.weight", "critic.4.bias",
]


def param_shapes(O: int, H: int, A: int) -> List[Tuple[int, ...]]:
    return [(H, O), (H,), (H, H), (H,), (A, H), (A,), (H, O), (H,), (H, H), (H,), (1, H), (1,)]


def param_count(O: int, H: int, A: int) -> int:
    return sum(int(np.prod(s)) for s in param_shapes(O, H, A))


def split_flat(flat: torch.Tensor, O: int, H: int, A: int) -> Dict[str, torch.Tensor]:
    out, off = {}, 0
    for name, shp in zip(PARAM_NAMES, param_shapes(O, H, A)):
        n = int(np.prod(shp))
        out[name] = flat[off:off + n].view(*shp)
        off += n
    assert off == flat.numel()
    return out


def init_params(O: int, H: int, A: int, seed: int) -> torch.Tensor:
    """layer_init of ppo.py:25-28 with the gains of ppo.py:35-46, consuming the torch CPU stream in
    the same order as `ActorCritic.__init__` (nn.Linear's own init draws first, then orthogonal_)."""
    torch.manual_seed(seed)
    gains = {"actor.0": 2 ** 0.5, "actor.2": 2 ** 0.5, "actor.4": 0.01,
             "critic.0": 2 ** 0.5, "critic.2": 2 ** 0.5, "critic.4": 1.0}
    dims = {"actor.0": (O, H), "actor.2": (H, H), "actor.4": (H, A),
            "critic.0": (O, H), "critic.2": (H, H), "critic.4": (H, 1)}
    chunks = []
    for key in ["actor.0", "actor.2", "actor.4", "critic.0", "critic.2", "critic.4"]:
        lin = torch.nn.Linear(*dims[key])
        torch.nn.init.orthogonal_(lin.weight, gains[key])
        torch.nn.init.constant_(lin.bias, 0.0)
        chunks += [lin.weight.detach().reshape(-1), lin.bias.detach().reshape(-1)]
    return torch.cat(chunks).clone()


def mlp_forward(flat: torch.Tensor, obs: torch.Tensor, O: int, H: int, A: int):
    """obs [..., O] -> (logits [..., A], value [...]).  ppo.py:34-54."""
    p = split_flat(flat, O, H, A)
    F = torch.nn.functional
    ha = torch.tanh(F.linear(obs, p["actor.0.weight"], p["actor.0.bias"]))
    ha = torch.tanh(F.linear(ha, p["actor.2.weight"], p["actor.2.bias"]))
    logits = F.linear(ha, p["actor.4.weight"], p["actor.4.bias"])
    hc = torch.tanh(F.linear(obs, p["critic.0.weight"], p["critic.0.bias"]))
    hc = torch.tanh(F.linear(hc, p["critic.2.weight"], p["critic.2.bias"]))
    value = F.linear(hc, p["critic.4.weight"], p["critic.4.bias"]).squeeze(-1)
    return logits, value


@dataclass
class Coeffs:
    clip_coef: float = 0.2
    ent_coef: float = 0.01
    vf_coef: float = 0.5


def minibatch_loss(flat, obs, act, logp_old, adv, ret, val_old, O, H, A, c: Coeffs = Coeffs(),
                   adv_mean=None, adv_std=None):
    """Loss of ppo.py:166-187 on one minibatch.  Returns (loss, pg_loss, v_loss, entropy).
    adv_mean/adv_std override the per-minibatch statistics (used for multi-rank checks)."""
    logits, new_values = mlp_forward(flat, obs, O, H, A)
    logp_all = logits - torch.logsumexp(logits, dim=-1, keepdim=True)
    mean = torch.mean(adv) if adv_mean is None else adv_mean
    std = torch.std(adv) if adv_std is None else adv_std
    nadv = (adv - mean) / (std + 1e-8)
    new_logp = logp_all.gather(-1, act.long().unsqueeze(-1)).squeeze(-1)
    ratio = torch.exp(new_logp - logp_old)
    pg1 = -nadv * ratio
    pg2 = -nadv * torch.clamp(ratio, 1 - c.clip_coef, 1 + c.clip_coef)
    pg_loss = torch.mean(torch.max(pg1, pg2))
    probs = torch.exp(logp_all)
    entropy = torch.mean(-(probs * logp_all).sum(-1))
    v_un = (new_values - ret) ** 2
    v_cl = val_old + torch.clamp(new_values - val_old, -c.clip_coef, c.clip_coef)
    v_loss = 0.5 * torch.mean(torch.max(v_un, (v_cl - ret) ** 2))
    loss = pg_loss - c.ent_coef * entropy + v_loss * c.vf_coef
    return loss, pg_loss, v_loss, entropy


def minibatch_loss_and_grad(flat_np, obs, act, logp_old, adv, ret, val_old, O, H, A, c: Coeffs = Coeffs(),
                            adv_mean=None, adv_std=None):
    flat = torch.tensor(np.asarray(flat_np, dtype=np.float32), requires_grad=True)
    t = lambda x, dt=torch.float32: torch.as_tensor(np.asarray(x), dtype=dt)
    terms = minibatch_loss(flat, t(obs), t(act, torch.int64), t(logp_old), t(adv), t(ret), t(val_old), O, H, A, c,
                           adv_mean, adv_std)
    terms[0].backward()
    return [float(x.detach()) for x in terms], flat.grad.detach().numpy().copy()


def clip_adam(flat_np, grad_np, m_np, v_np, step: int, lr: float, max_grad_norm: float = 0.5,
              beta1=0.9, beta2=0.999, eps=1e-5):
    """clip_grad_norm_ (ppo.py:191) then one torch.optim.Adam step (ppo.py:90,192) on the flat vector.
    `step` is the 1-based step count AFTER this call.  Returns (params, m, v, total_norm)."""
    p = torch.nn.Parameter(torch.tensor(np.asarray(flat_np, dtype=np.float32)))
    p.grad = torch.tensor(np.asarray(grad_np, dtype=np.float32))
    opt = torch.optim.Adam([p], lr=lr, betas=(beta1, beta2), eps=eps)
    st = opt.state[p]
    st["step"] = torch.tensor(float(step - 1))
    st["exp_avg"] = torch.tensor(np.asarray(m_np, dtype=np.float32))
    st["exp_avg_sq"] = torch.tensor(np.asarray(v_np, dtype=np.float32))
    norm = torch.nn.utils.clip_grad_norm_([p], max_grad_norm)
    opt.step()
    return (p.detach().numpy().copy(), st["exp_avg"].numpy().copy(), st["exp_avg_sq"].numpy().copy(), float(norm))


# ----------------------------------------------------------------------------------------------
# Rollout with the build's Philox streams (N envs), order of operations of ppo.py:110-141.
# Buffers are [T+1, N] with the reference's one-slot shift.
# ----------------------------------------------------------------------------------------------
def rollout(flat_np, env: clib.OracleVecEnv, obs0: np.ndarray, T: int, H: int, carry=None):
    O, A, N = env.obs_dim, env.num_actions, env.n
    flat = torch.tensor(np.asarray(flat_np, dtype=np.float32))
    obs = np.zeros((T + 1, N, O), np.float32)
    val = np.zeros((T + 1, N), np.float32)
    act = np.zeros((T + 1, N), np.int32)
    logp = np.zeros((T + 1, N), np.float32)
    rew = np.zeros((T + 1, N), np.float32)
    done = np.zeros((T + 1, N), np.float32)
    logits_all = np.zeros((T, N, A), np.float32)
    if carry is not None:  # slot 0 of rew/done is stale from the previous rollout (ppo.py:93-98)
        rew[0], done[0] = carry
    fin = []
    obs[0] = obs0
    with torch.no_grad():
        for t in range(T + 1):
            logits, v = mlp_forward(flat, torch.from_numpy(obs[t]), O, H, A)
            val[t] = v.numpy()
            if t == T:
                break
            logits_all[t] = logits.numpy()
            a, lp = clib.sample(logits_all[t], env.seed, env.gid0, env.step_count)
            act[t], logp[t] = a, lp
            o, r, d, info = env.step(a)
            obs[t + 1], rew[t + 1], done[t + 1] = o, r, d
            for i in np.nonzero(d)[0]:
                fin.append((env.step_count - 1, int(i), float(info["final_return"][i]), int(info["final_length"][i])))
    return dict(obs=obs, val=val, act=act, logp=logp, rew=rew, done=done, logits=logits_all, episodes=fin)


# ----------------------------------------------------------------------------------------------
# bf16-emulating restatement of the tensor-core kernels' numerics (deep_rl_b200/csrc/update_tc.cu, rollout_tc.cu).
# The GEMM operands are rounded to bf16 at exactly the points where the kernels write their shared-memory operand
# tiles (observation hi + lo halves, W1 | b1, h1, W2, h2 / dout for dW4, dz2, dz1); products of bf16 values are exact
# in fp32 and the accumulation is done in float64 here (fp32 in tensor memory on the GPU), everything else is fp32.
# `tanh_fn` is the activation: the kernels use the SFU instruction tanh.approx.f32, which has no CPU equivalent, so the
# GPU tests pass the device's own elementwise tanh.approx (drl_selftest_tanh); torch.tanh gives the idealised variant.
# With the same activation function the kernels must agree with this oracle to ~1e-4 (fp32 accumulation order),
# two orders of magnitude tighter than their distance to the fp32 reference maths (which is bf16 rounding noise).
# ----------------------------------------------------------------------------------------------
def _bf(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def _mm(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """fp32 <- exact products, float64 accumulation."""
    return (a.double() @ b.double()).float()


def _net_forward_bf16(p, pre: str, obs: torch.Tensor, tanh_fn):
    W1, b1, W2, b2 = p[pre + ".0.weight"], p[pre + ".0.bias"], p[pre + ".2.weight"], p[pre + ".2.bias"]
    W4, b4 = p[pre + ".4.weight"], p[pre + ".4.bias"]
    hi = _bf(obs)
    lo = _bf(obs - hi)
    W1b = _bf(W1)
    z1 = ((hi.double() @ W1b.double().T) + _bf(b1).double() + (lo.double() @ W1b.double().T)).float()
    h1 = _bf(tanh_fn(z1))
    z2 = _mm(h1, _bf(W2).T)
    h2 = tanh_fn(z2 + b2)
    out = _mm(h2, W4.T) + b4
    return dict(hi=hi, lo=lo, h1=h1, h2=h2, out=out, W2b=_bf(W2), W4=W4)


def mlp_forward_bf16_emulated(flat, obs, O: int, H: int, A: int, tanh_fn=torch.tanh):
    """(logits, value) as rollout_tc_kernel / ppo_grad_tc_kernel compute them."""
    p = split_flat(torch.as_tensor(np.asarray(flat, dtype=np.float32)), O, H, A)
    x = torch.as_tensor(np.asarray(obs, dtype=np.float32))
    with torch.no_grad():
        a = _net_forward_bf16(p, "actor", x, tanh_fn)
        c = _net_forward_bf16(p, "critic", x, tanh_fn)
    return a["out"], c["out"].squeeze(-1)


def minibatch_grad_bf16_emulated(flat_np, obs, act, logp_old, adv, val_old, O, H, A, adv_mean: float, adv_std: float,
                                 c: Coeffs = Coeffs(), tanh_fn=torch.tanh):
    """Loss terms [loss, pg, v, entropy, approx_kl, clipfrac] and the flat gradient (state_dict order) of one minibatch,
    with the closed-form backward and the bf16 rounding points of ppo_grad_tc_kernel."""
    t = lambda x, dt=torch.float32: torch.as_tensor(np.asarray(x), dtype=dt)
    p = split_flat(t(flat_np), O, H, A)
    x, act, logp_old, adv, val_old = t(obs), t(act, torch.int64), t(logp_old), t(adv), t(val_old)
    M = x.shape[0]
    inv_m = torch.tensor(1.0 / M, dtype=torch.float32)
    f32 = lambda v: torch.tensor(v, dtype=torch.float32)
    grads = {}
    with torch.no_grad():
        # ---- actor ----
        a = _net_forward_bf16(p, "actor", x, tanh_fn)
        out = a["out"]
        m = out.max(-1, keepdim=True).values
        lse = m + torch.log(torch.exp(out - m).sum(-1, keepdim=True))
        lp = out - lse
        pr = torch.exp(lp)
        ent = -(pr * lp).sum(-1)
        new_logp = lp.gather(-1, act.unsqueeze(-1)).squeeze(-1)
        nadv = (adv - f32(adv_mean)) * (f32(1.0) / (f32(adv_std) + f32(1e-8)))
        logratio = new_logp - logp_old
        ratio = torch.exp(logratio)
        pg1 = -nadv * ratio
        pg2 = -nadv * torch.clamp(ratio, 1.0 - c.clip_coef, 1.0 + c.clip_coef)
        dpg = torch.where(pg1 >= pg2, pg1, torch.zeros_like(pg1))
        onehot = torch.nn.functional.one_hot(act, A).float()
        d_act = inv_m * (dpg.unsqueeze(-1) * (onehot - pr) + f32(c.ent_coef) * pr * (lp + ent.unsqueeze(-1)))
        pg_loss = torch.maximum(pg1, pg2).sum() * inv_m
        entropy = ent.sum() * inv_m
        approx_kl = ((ratio - 1.0) - logratio).sum() * inv_m
        clipfrac = ((ratio - 1.0).abs() > c.clip_coef).float().sum() * inv_m
        # ---- critic ----
        cr = _net_forward_bf16(p, "critic", x, tanh_fn)
        v = cr["out"].squeeze(-1)
        ret = adv + val_old
        vd = v - ret
        vu = vd * vd
        vdiff = v - val_old
        vcd = (val_old + torch.clamp(vdiff, -c.clip_coef, c.clip_coef)) - ret
        vcl = vcd * vcd
        gcl = torch.where((vdiff >= -c.clip_coef) & (vdiff <= c.clip_coef), vcd, torch.zeros_like(vcd))
        gv = torch.where(vu > vcl, vd, torch.where(vcl > vu, gcl, 0.5 * (vd + gcl)))
        d_cri = (f32(c.vf_coef) * gv * inv_m).unsqueeze(-1)
        v_loss = 0.5 * torch.maximum(vu, vcl).sum() * inv_m
        # ---- backward, identical for both nets ----
        for pre, f, d in (("actor", a, d_act), ("critic", cr, d_cri)):
            h1, h2, W2b, W4 = f["h1"], f["h2"], f["W2b"], f["W4"]
            dz2 = _bf((d @ W4) * (1.0 - h2 * h2))
            dh1 = _mm(dz2, W2b)
            dz1 = _bf(dh1 * (1.0 - h1 * h1))
            grads[pre + ".0.weight"] = _mm(dz1.T, f["hi"]) + _mm(dz1.T, f["lo"])
            grads[pre + ".0.bias"] = dz1.double().sum(0).float()
            grads[pre + ".2.weight"] = _mm(dz2.T, h1)
            grads[pre + ".2.bias"] = dz2.double().sum(0).float()
            grads[pre + ".4.weight"] = _mm(_bf(d).T, _bf(h2))
            grads[pre + ".4.bias"] = d.double().sum(0).float()
    loss = pg_loss - c.ent_coef * entropy + v_loss * c.vf_coef
    flat_grad = torch.cat([grads[n].reshape(-1) for n in PARAM_NAMES]).numpy().copy()
    terms = [float(loss), float(pg_loss), float(v_loss), float(entropy), float(approx_kl), float(clipfrac)]
    return terms, flat_grad
