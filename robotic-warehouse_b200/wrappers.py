"""Host/device-side reshapes on top of the step outputs, mirroring
rware/utils/wrappers.py: `FlattenAgents` (9-48), `DictAgents` (51-80) and `FlattenSAObservation` (83-108) for the single-world facade, plus
`VecFlattenAgents`, the batched form (a view of the observation tensor and a row sum of the rewards -- no copy, no
kernel).  They work without gymnasium; they only need an object with the Warehouse / VecWarehouse surface."""
import math

import numpy as np

from . import spaces
from .warehouse import Action


class _Wrapper:
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return getattr(self.env, "unwrapped", self.env)


class FlattenAgents(_Wrapper):
    """One joint agent: concatenated observations, summed rewards, concatenated actions (wrappers.py:9-48)."""

    def __init__(self, env):
        super().__init__(env)
        n, mb = env.n_agents, env.msg_bits
        sa = [len(Action), *mb * (2,)]
        self.action_space = spaces.Discrete(sa[0]) if (len(sa) == 1 and n == 1) else spaces.MultiDiscrete(n * sa)
        self.observation_space = spaces.Tuple(tuple(s for s in env.observation_space))

    def _flat(self, obs):
        # every agent's observation goes through its space's flatten rule first (wrappers.py:26-32, 41-46), so DICT /
        # IMAGE / IMAGE_DICT observations become vectors too
        return np.concatenate([spaces.flatten(s, o) for s, o in zip(self.observation_space, obs)])

    def reset(self, **kwargs):
        obs, info = self.env.reset(**kwargs)
        return self._flat(obs), info

    def step(self, action):
        try:
            action = np.split(np.asarray(action), self.env.n_agents)
            action = [a if self.env.msg_bits else int(a[0]) for a in action]
        except (AttributeError, IndexError, ValueError):
            action = [action]
        obs, reward, done, truncated, info = self.env.step(action)
        return self._flat(obs), np.sum(reward), done, truncated, info


class DictAgents(_Wrapper):
    """Per-agent dictionaries keyed `agent_<i>` (wrappers.py:51-80; `reset` returns (obs, info) here -- the
    reference's version iterates the (obs, info) tuple by mistake)."""

    def _keys(self):
        digits = int(math.log10(self.env.n_agents)) + 1
        return [f"agent_{i:{digits}}" for i in range(self.env.n_agents)]

    def reset(self, **kwargs):
        obs, info = self.env.reset(**kwargs)
        return dict(zip(self._keys(), obs)), info

    def step(self, action):
        keys = self._keys()
        assert keys == sorted(action.keys())
        obs, reward, done, truncated, info = self.env.step([action[k] for k in keys])
        return (dict(zip(keys, obs)), dict(zip(keys, reward)), {k: done for k in keys}, {k: truncated for k in keys}, info)


class FlattenSAObservation(_Wrapper):
    """Flattens every agent's observation with the space's own flatten rule (wrappers.py:83-108): DICT and IMAGE_DICT
    observations become one vector per agent (Discrete members one-hot, Box / MultiBinary raveled, in the key order of the
    space)."""

    def __init__(self, env):
        super().__init__(env)
        inf = float("inf")
        self.observation_space = spaces.Tuple(tuple(
            spaces.Box(low=-inf, high=inf, shape=(spaces.flatdim(sa),), dtype=np.float32) for sa in env.observation_space))

    def observation(self, observation):
        return [spaces.flatten(sa, obs) for sa, obs in zip(self.env.observation_space, observation)]

    def reset(self, **kwargs):
        obs, info = self.env.reset(**kwargs)
        return self.observation(obs), info

    def step(self, action):
        obs, reward, done, truncated, info = self.env.step(action)
        return self.observation(obs), reward, done, truncated, info


class VecFlattenAgents(_Wrapper):
    """Batched FlattenAgents over a VecWarehouse: obs [E, n*L] (a view), reward [E] (sum over agents), done [E]."""

    def reset(self, **kwargs):
        return self.env.reset(**kwargs).view(self.env.num_envs, -1)

    def step(self, actions):
        obs, reward, done = self.env.step(actions)
        return obs.view(self.env.num_envs, -1), reward.sum(dim=1), done
