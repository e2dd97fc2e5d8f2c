"""Minimal stand-in for the `gymnasium` package (TEST INFRASTRUCTURE ONLY).

gymnasium is not installed in the build container and cannot be fetched (no network).
This stand-in provides exactly the surface the *unmodified* reference
(`/root/reference/rware`, never copied into this repo) touches, so that the reference can be
imported here to (a) pin the C oracle in `oracle/` and (b) generate the golden fixtures under
`tests/golden/`.  Nothing in the product package imports this module.

Surface: Env, Wrapper, ObservationWrapper, register/make, spaces.*,
utils.seeding.np_random, error.Error.
"""
import importlib

import numpy as np

from . import error, spaces  # noqa: F401
from .utils import seeding

__version__ = "0.0-standin"


class Env:
    metadata = {"render_modes": []}
    render_mode = None
    spec = None
    _np_random = None
    _np_random_seed = None

    @property
    def unwrapped(self):
        return self

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random, self._np_random_seed = seeding.np_random()
        return self._np_random

    @np_random.setter
    def np_random(self, value):
        self._np_random = value
        self._np_random_seed = -1

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self._np_random, self._np_random_seed = seeding.np_random(seed)

    def step(self, action):
        raise NotImplementedError

    def render(self):
        raise NotImplementedError

    def close(self):
        pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self._action_space = None
        self._observation_space = None

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    @property
    def action_space(self):
        return self.env.action_space if self._action_space is None else self._action_space

    @action_space.setter
    def action_space(self, space):
        self._action_space = space

    @property
    def observation_space(self):
        if self._observation_space is None:
            return self.env.observation_space
        return self._observation_space

    @observation_space.setter
    def observation_space(self, space):
        self._observation_space = space

    def reset(self, *, seed=None, options=None):
        return self.env.reset(seed=seed, options=options)

    def step(self, action):
        return self.env.step(action)

    def render(self):
        return self.env.render()

    def close(self):
        return self.env.close()


class ObservationWrapper(Wrapper):
    def reset(self, *, seed=None, options=None):
        obs, info = self.env.reset(seed=seed, options=options)
        return self.observation(obs), info

    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        return self.observation(obs), reward, terminated, truncated, info

    def observation(self, observation):
        raise NotImplementedError


registry = {}


def register(id, entry_point=None, kwargs=None, **_ignored):
    registry[id] = (entry_point, dict(kwargs or {}))


def make(id, **overrides):
    if ":" in id and id not in registry:
        module, id = id.split(":", 1)
        importlib.import_module(module)
    entry_point, kwargs = registry[id]
    kwargs = {**kwargs, **overrides}
    if isinstance(entry_point, str):
        mod, attr = entry_point.split(":")
        entry_point = getattr(importlib.import_module(mod), attr)
    return entry_point(**kwargs)
