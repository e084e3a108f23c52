"""Minimal stand-in for `gym==0.21` -- CPU ORACLE, test infrastructure only.

gym 0.21 cannot be installed in this environment (no network; incompatible with numpy 2.x), so
this package restates just enough of its 0.21 semantics [gym-recall, SURVEY.md App. A] for the
UNMODIFIED reference script deep_rl/ppo.py to run: `gym.Env`, `gym.Wrapper`, `gym.make`,
`gym.wrappers.RecordEpisodeStatistics`, spaces with `.shape` / `.n`, `env.seed`, `env.close`,
and the 4-tuple step API.  Physics is delegated to the C oracle (oracle/drl_oracle.c).

Deviation that cannot be avoided: the env's reset RNG is `np.random.RandomState(seed)`, not
gym's SHA-512 `hash_seed` stream, so episodes are not bit-comparable to a real gym run.
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import clib as _clib  # noqa: E402

__version__ = "0.21.0-shim"


class _Box:
    def __init__(self, shape):
        self.shape = tuple(shape)
        self.dtype = np.float32


class _Discrete:
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.int64
        self._rng = np.random.RandomState()

    def seed(self, seed=None):            # gym.spaces.Space.seed / .sample (dqn.py:67,94)
        self._rng = np.random.RandomState(seed)
        return [seed]

    def sample(self):
        return int(self._rng.randint(self.n))


class Env:
    observation_space = None
    action_space = None

    def step(self, action):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    def seed(self, seed=None):
        return [seed]

    def close(self):
        pass

    @property
    def unwrapped(self):
        return self


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self.observation_space = env.observation_space
        self.action_space = env.action_space

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    def step(self, action):
        return self.env.step(action)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def seed(self, seed=None):
        return self.env.seed(seed)

    def close(self):
        return self.env.close()

    @property
    def unwrapped(self):
        return self.env.unwrapped


class _ClassicControl(Env):
    """One CartPole-v1 / Acrobot-v1 instance; float64 state, float32 observation."""

    def __init__(self, env_id):
        self._kind = _clib.ENV_KINDS[env_id]
        self._half = 0.05 if self._kind == _clib.ENV_CARTPOLE else 0.1
        self._rng = np.random.RandomState()
        self._state = np.zeros(4, dtype=np.float64)
        self.observation_space = _Box((_clib.lib().drl_or_obs_dim(self._kind),))
        self.action_space = _Discrete(_clib.lib().drl_or_num_actions(self._kind))

    def seed(self, seed=None):
        self._rng = np.random.RandomState(seed)
        return [seed]

    def _obs(self):
        vec = _clib.OracleVecEnv.__new__(_clib.OracleVecEnv)
        vec.kind, vec.n, vec.obs_dim, vec.state = self._kind, 1, self.observation_space.shape[0], self._state.reshape(1, 4)
        return vec.observe()[0]

    def reset(self):
        self._state = self._rng.uniform(low=-self._half, high=self._half, size=(4,)).astype(np.float64)
        return self._obs()

    def step(self, action):
        a = int(np.asarray(action).item())
        if self._kind == _clib.ENV_CARTPOLE:
            self._state, term = _clib.cartpole_step(self._state, a)
            reward = 1.0
        else:
            self._state, reward, term = _clib.acrobot_step(self._state, a)
        return self._obs(), reward, term, {}


class _TimeLimit(Wrapper):
    def __init__(self, env, max_episode_steps):
        super().__init__(env)
        self._max_episode_steps = max_episode_steps
        self._elapsed_steps = None

    def step(self, action):
        observation, reward, done, info = self.env.step(action)
        self._elapsed_steps += 1
        if self._elapsed_steps >= self._max_episode_steps:
            info["TimeLimit.truncated"] = not done
            done = True
        return observation, reward, done, info

    def reset(self, **kwargs):
        self._elapsed_steps = 0
        return self.env.reset(**kwargs)


_REGISTRY = {"CartPole-v1": 500, "Acrobot-v1": 500}


class _Spec:
    def __init__(self, env_id, max_episode_steps):
        self.id, self.max_episode_steps = env_id, max_episode_steps


def make(env_id):
    env = _TimeLimit(_ClassicControl(env_id), _REGISTRY[env_id])
    env.spec = _Spec(env_id, _REGISTRY[env_id])      # env.spec.max_episode_steps (reinforce.py:53-54)
    return env


from . import wrappers  # noqa: E402,F401
