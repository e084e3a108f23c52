"""ctypes bindings for oracle/libdrl_oracle.so -- CPU ORACLE, test infrastructure only.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (deep_rl_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libdrl_oracle.so")

ENV_CARTPOLE, ENV_ACROBOT, ENV_MOUNTAINCAR = 0, 1, 2
ENV_KINDS = {"CartPole-v1": ENV_CARTPOLE, "Acrobot-v1": ENV_ACROBOT, "MountainCar-v0": ENV_MOUNTAINCAR}
MAX_EPISODE_STEPS = {"CartPole-v1": 500, "Acrobot-v1": 500, "MountainCar-v0": 200}
UINT64_MAX = (1 << 64) - 1


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "drl_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libdrl_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u32p, i32p, f32p, f64p, u8p = (C.POINTER(t) for t in (C.c_uint32, C.c_int32, C.c_float, C.c_double, C.c_uint8))
        L.drl_or_philox4x32_10.argtypes = [u32p, u32p, u32p]
        L.drl_or_exp_det.argtypes = [C.c_float]
        L.drl_or_exp_det.restype = C.c_float
        L.drl_or_sample_from_uniform.argtypes = [f32p, C.c_int32, C.c_float, f32p]
        L.drl_or_sample_from_uniform.restype = C.c_int32
        L.drl_or_action_uniform.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64]
        L.drl_or_action_uniform.restype = C.c_float
        L.drl_or_sample.argtypes = [f32p, C.c_int32, C.c_int32, C.c_uint64, C.c_uint32, C.c_uint64, i32p, f32p]
        L.drl_or_perm_index.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32]
        L.drl_or_perm_index.restype = C.c_uint32
        L.drl_or_perm_position.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32]
        L.drl_or_perm_position.restype = C.c_uint32
        L.drl_or_permutation.argtypes = [u32p, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32]
        L.drl_or_cartpole_step.argtypes = [f64p, C.c_int32]
        L.drl_or_cartpole_step.restype = C.c_int32
        L.drl_or_acrobot_step.argtypes = [f64p, C.c_int32, f64p]
        L.drl_or_acrobot_step.restype = C.c_int32
        L.drl_or_mountaincar_step.argtypes = [f64p, C.c_int32]
        L.drl_or_mountaincar_step.restype = C.c_int32
        L.drl_or_obs_dim.argtypes = [C.c_int32]
        L.drl_or_obs_dim.restype = C.c_int32
        L.drl_or_num_actions.argtypes = [C.c_int32]
        L.drl_or_num_actions.restype = C.c_int32
        L.drl_or_reset_state.argtypes = [C.c_int32, C.c_uint64, C.c_uint32, C.c_uint64, f64p]
        L.drl_or_vec_step.argtypes = [C.c_int32, C.c_int32, C.c_uint64, C.c_uint32, C.c_uint64, f64p, i32p, f32p,
                                      i32p, i32p, f32p, f32p, u8p, f32p, i32p, C.c_int32]
        L.drl_or_vec_reset.argtypes = [C.c_int32, C.c_int32, C.c_uint64, C.c_uint32, f64p, i32p, f32p, i32p, f32p]
        L.drl_or_vec_obs.argtypes = [C.c_int32, C.c_int32, f64p, f32p]
        L.drl_or_gae.argtypes = [f32p, f32p, f32p, f32p, f32p, C.c_int32, C.c_int32, C.c_float, C.c_float]
        L.drl_or_replay_uniform.argtypes = [u32p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64]
        L.drl_or_replay_priority.argtypes = [f32p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, u32p, f64p]
        _lib = L
    return _lib


def _p(a: np.ndarray, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def philox4x32_10(ctr, key) -> np.ndarray:
    c = np.asarray(ctr, dtype=np.uint32).copy()
    k = np.asarray(key, dtype=np.uint32).copy()
    o = np.zeros(4, dtype=np.uint32)
    lib().drl_or_philox4x32_10(_p(c, C.c_uint32), _p(k, C.c_uint32), _p(o, C.c_uint32))
    return o


def exp_det(x: float) -> float:
    return float(lib().drl_or_exp_det(C.c_float(x)))


def sample(logits: np.ndarray, seed: int, env_gid0: int, step: int):
    """logits [n, A] float32 -> (actions int32 [n], logp float32 [n])."""
    lg = np.ascontiguousarray(logits, dtype=np.float32)
    n, A = lg.shape
    act = np.zeros(n, dtype=np.int32)
    lp = np.zeros(n, dtype=np.float32)
    lib().drl_or_sample(_p(lg, C.c_float), n, A, seed, env_gid0, step, _p(act, C.c_int32), _p(lp, C.c_float))
    return act, lp


def action_uniform(seed: int, env_gid: int, step: int) -> float:
    return float(lib().drl_or_action_uniform(seed, env_gid, step))


def permutation(B: int, seed: int, epoch_ctr: int, rank: int = 0) -> np.ndarray:
    out = np.zeros(B, dtype=np.uint32)
    lib().drl_or_permutation(_p(out, C.c_uint32), B, seed, epoch_ctr, rank)
    return out


def cartpole_step(state: np.ndarray, action: int):
    s = np.array(state, dtype=np.float64)
    term = lib().drl_or_cartpole_step(_p(s, C.c_double), int(action))
    return s, bool(term)


def acrobot_step(state: np.ndarray, action: int):
    s = np.array(state, dtype=np.float64)
    r = C.c_double(0.0)
    term = lib().drl_or_acrobot_step(_p(s, C.c_double), int(action), C.byref(r))
    return s, float(r.value), bool(term)


def mountaincar_step(state: np.ndarray, action: int):
    s = np.zeros(4, dtype=np.float64)
    s[:2] = np.asarray(state, dtype=np.float64)[:2]
    term = lib().drl_or_mountaincar_step(_p(s, C.c_double), int(action))
    return s[:2].copy(), bool(term)


def gae(rew: np.ndarray, done: np.ndarray, val: np.ndarray, gamma: float, lam: float):
    """rew/done/val [T+1, N] float32 (one-slot shift, ppo.py:93-98) -> adv, ret [T+1, N]."""
    rew = np.ascontiguousarray(rew, dtype=np.float32)
    done = np.ascontiguousarray(done, dtype=np.float32)
    val = np.ascontiguousarray(val, dtype=np.float32)
    T1, N = rew.shape
    adv = np.zeros_like(rew)
    ret = np.zeros_like(rew)
    lib().drl_or_gae(_p(rew, C.c_float), _p(done, C.c_float), _p(val, C.c_float), _p(adv, C.c_float),
                     _p(ret, C.c_float), T1 - 1, N, gamma, lam)
    return adv, ret


class OracleVecEnv:
    """N independent envs with the wrapper chain of ppo.py:79 and auto-reset of ppo.py:127-129."""

    def __init__(self, env_id: str, num_envs: int, seed: int, env_gid0: int = 0, max_episode_steps: int = 0):
        self.kind = ENV_KINDS[env_id]
        max_episode_steps = max_episode_steps or MAX_EPISODE_STEPS[env_id]      # TimeLimit of the gym registration
        self.n = int(num_envs)
        self.seed = int(seed)
        self.gid0 = int(env_gid0)
        self.max_steps = int(max_episode_steps)
        self.obs_dim = lib().drl_or_obs_dim(self.kind)
        self.num_actions = lib().drl_or_num_actions(self.kind)
        self.state = np.zeros((self.n, 4), dtype=np.float64)
        self.elapsed = np.zeros(self.n, dtype=np.int32)
        self.ep_ret = np.zeros(self.n, dtype=np.float32)
        self.ep_len = np.zeros(self.n, dtype=np.int32)
        self.step_count = 0  # global step index fed to the Philox counter

    def reset(self) -> np.ndarray:
        obs = np.zeros((self.n, self.obs_dim), dtype=np.float32)
        lib().drl_or_vec_reset(self.kind, self.n, self.seed, self.gid0, _p(self.state, C.c_double),
                               _p(self.elapsed, C.c_int32), _p(self.ep_ret, C.c_float),
                               _p(self.ep_len, C.c_int32), _p(obs, C.c_float))
        return obs

    def set_state(self, state: np.ndarray) -> np.ndarray:
        self.state[...] = np.asarray(state, dtype=np.float64).reshape(self.n, 4)
        return self.observe()

    def observe(self) -> np.ndarray:
        obs = np.zeros((self.n, self.obs_dim), dtype=np.float32)
        lib().drl_or_vec_obs(self.kind, self.n, _p(self.state, C.c_double), _p(obs, C.c_float))
        return obs

    def step(self, actions: np.ndarray):
        a = np.ascontiguousarray(actions, dtype=np.int32).reshape(self.n)
        obs = np.zeros((self.n, self.obs_dim), dtype=np.float32)
        rew = np.zeros(self.n, dtype=np.float32)
        done = np.zeros(self.n, dtype=np.uint8)
        fin_ret = np.full(self.n, np.nan, dtype=np.float32)
        fin_len = np.full(self.n, -1, dtype=np.int32)
        lib().drl_or_vec_step(self.kind, self.n, self.seed, self.gid0, self.step_count, _p(self.state, C.c_double),
                              _p(self.elapsed, C.c_int32), _p(self.ep_ret, C.c_float), _p(self.ep_len, C.c_int32),
                              _p(a, C.c_int32), _p(obs, C.c_float), _p(rew, C.c_float), _p(done, C.c_uint8),
                              _p(fin_ret, C.c_float), _p(fin_len, C.c_int32), self.max_steps)
        self.step_count += 1
        return obs, rew, done, {"final_return": fin_ret, "final_length": fin_len}


# ---- replay-buffer samplers and gather (deep_rl/dqn.py:116-122, deep_rl/per.py:127-146) ----
def replay_uniform(batch: int, size: int, seed: int, draw_ctr: int) -> np.ndarray:
    out = np.zeros(batch, dtype=np.uint32)
    lib().drl_or_replay_uniform(_p(out, C.c_uint32), batch, size, seed, draw_ctr)
    return out


def replay_priority(priorities: np.ndarray, batch: int, seed: int, draw_ctr: int) -> np.ndarray:
    pri = np.ascontiguousarray(priorities, dtype=np.float32)
    out = np.zeros(batch, dtype=np.uint32)
    scratch = np.zeros((len(pri) + 1023) // 1024 + 1, dtype=np.float64)
    lib().drl_or_replay_priority(_p(pri, C.c_float), len(pri), batch, seed, draw_ctr, _p(out, C.c_uint32), _p(scratch, C.c_double))
    return out


def replay_gather(obs, act, rew, term, idx):
    """The five gathers of dqn.py:118-122 with numpy fancy indexing (one-slot shift: reward / terminated / next obs at idx + 1)."""
    i = np.asarray(idx, dtype=np.int64)
    return obs[i], act[i], obs[i + 1], rew[i + 1], term[i + 1]


def replay_probabilities(priorities: np.ndarray, idx, alpha: float) -> np.ndarray:
    """per.py:128,131: (priorities ** alpha / sum(priorities ** alpha))[batch_inds], float32 like the reference."""
    import torch
    p = torch.as_tensor(np.asarray(priorities, dtype=np.float32))
    prob = p ** alpha / torch.sum(p ** alpha)
    return prob[torch.as_tensor(np.asarray(idx, dtype=np.int64))].numpy()
