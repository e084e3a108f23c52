"""gym.wrappers subset of the 0.21 shim -- CPU ORACLE, test infrastructure only."""
import time
from collections import deque

import numpy as np

from . import Wrapper


class RecordEpisodeStatistics(Wrapper):
    """[gym-recall, SURVEY.md App. A.4] float32 return / int32 length accumulators; on done the
    info dict gains {"episode": {"r", "l", "t"}} and the accumulators are zeroed."""

    def __init__(self, env, deque_size=100):
        super().__init__(env)
        self.num_envs = getattr(env, "num_envs", 1)
        self.t0 = time.perf_counter()
        self.episode_count = 0
        self.episode_returns = None
        self.episode_lengths = None
        self.return_queue = deque(maxlen=deque_size)
        self.length_queue = deque(maxlen=deque_size)

    def reset(self, **kwargs):
        observations = self.env.reset(**kwargs)
        self.episode_returns = np.zeros(self.num_envs, dtype=np.float32)
        self.episode_lengths = np.zeros(self.num_envs, dtype=np.int32)
        return observations

    def step(self, action):
        observation, reward, done, info = self.env.step(action)
        self.episode_returns += reward
        self.episode_lengths += 1
        if done:
            info = dict(info)
            info["episode"] = {
                "r": self.episode_returns[0],
                "l": self.episode_lengths[0],
                "t": round(time.perf_counter() - self.t0, 6),
            }
            self.return_queue.append(self.episode_returns[0])
            self.length_queue.append(self.episode_lengths[0])
            self.episode_count += 1
            self.episode_returns[0] = 0
            self.episode_lengths[0] = 0
        return observation, reward, done, info
