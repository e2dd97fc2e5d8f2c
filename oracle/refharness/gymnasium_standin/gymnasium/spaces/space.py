import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None):
        self._shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)
        self._rng = None

    @property
    def shape(self):
        return self._shape

    @property
    def np_random(self):
        if self._rng is None:
            self._rng = np.random.default_rng()
        return self._rng

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)

    def __contains__(self, x):
        return self.contains(x)
