"""Spaces subset: Discrete, MultiDiscrete, MultiBinary, Box, Tuple, Dict, flatdim, flatten."""
from collections import OrderedDict
from collections.abc import Mapping

import numpy as np

from . import space  # noqa: F401
from .space import Space


class Discrete(Space):
    def __init__(self, n, start=0):
        super().__init__((), np.int64)
        self.n = int(n)
        self.start = int(start)

    def sample(self):
        return int(self.start + self.np_random.integers(self.n))

    def contains(self, x):
        if isinstance(x, (np.generic, np.ndarray)):
            if not (np.issubdtype(np.asarray(x).dtype, np.integer) and np.asarray(x).shape == ()):
                return False
            x = int(x)
        elif not isinstance(x, int):
            return False
        return self.start <= x < self.start + self.n

    def __eq__(self, other):
        return isinstance(other, Discrete) and self.n == other.n and self.start == other.start

    def __repr__(self):
        return f"Discrete({self.n})"


class MultiDiscrete(Space):
    def __init__(self, nvec, dtype=np.int64):
        self.nvec = np.array(nvec, dtype=dtype, copy=True)
        super().__init__(self.nvec.shape, dtype)

    def sample(self):
        return (self.np_random.random(self.nvec.shape) * self.nvec).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return bool(x.shape == self.shape and np.issubdtype(x.dtype, np.integer)
                    and np.all(0 <= x) and np.all(x < self.nvec))

    def __eq__(self, other):
        return isinstance(other, MultiDiscrete) and self.nvec.shape == other.nvec.shape \
            and bool(np.all(self.nvec == other.nvec))

    def __repr__(self):
        return f"MultiDiscrete({self.nvec})"


class MultiBinary(Space):
    def __init__(self, n):
        self.n = n
        shape = (n,) if isinstance(n, (int, np.integer)) else tuple(n)
        super().__init__(shape, np.int8)

    def sample(self):
        return self.np_random.integers(0, 2, size=self.shape, dtype=self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return bool(x.shape == self.shape and np.all((x == 0) | (x == 1)))

    def __eq__(self, other):
        return isinstance(other, MultiBinary) and self.shape == other.shape

    def __repr__(self):
        return f"MultiBinary({self.n})"


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        dtype = np.dtype(dtype)
        if shape is None:
            shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
        shape = tuple(shape)
        super().__init__(shape, dtype)
        with np.errstate(invalid="ignore", over="ignore"):
            self.low = np.broadcast_to(np.asarray(low, dtype=np.float64), shape).astype(dtype)
            self.high = np.broadcast_to(np.asarray(high, dtype=np.float64), shape).astype(dtype)
        self._low64 = np.broadcast_to(np.asarray(low, dtype=np.float64), shape)
        self._high64 = np.broadcast_to(np.asarray(high, dtype=np.float64), shape)

    def sample(self):
        lo = np.where(np.isfinite(self._low64), self._low64, -1e6)
        hi = np.where(np.isfinite(self._high64), self._high64, 1e6)
        return self.np_random.uniform(lo, hi).astype(self.dtype)

    def contains(self, x):
        if not isinstance(x, np.ndarray):
            try:
                x = np.asarray(x, dtype=self.dtype)
            except (ValueError, TypeError):
                return False
        return bool(np.can_cast(x.dtype, self.dtype) and x.shape == self.shape
                    and np.all(x >= self.low) and np.all(x <= self.high))

    def __eq__(self, other):
        return isinstance(other, Box) and self.shape == other.shape and self.dtype == other.dtype \
            and np.allclose(self.low, other.low) and np.allclose(self.high, other.high)

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


class Tuple(Space):
    def __init__(self, spaces):
        self.spaces = tuple(spaces)
        super().__init__(None, None)

    def sample(self):
        return tuple(s.sample() for s in self.spaces)

    def contains(self, x):
        if isinstance(x, (list, np.ndarray)):
            x = tuple(x)
        return isinstance(x, tuple) and len(x) == len(self.spaces) and all(
            s.contains(p) for s, p in zip(self.spaces, x))

    def __getitem__(self, i):
        return self.spaces[i]

    def __len__(self):
        return len(self.spaces)

    def __iter__(self):
        return iter(self.spaces)

    def __eq__(self, other):
        return isinstance(other, Tuple) and self.spaces == other.spaces

    def __repr__(self):
        return "Tuple(" + ", ".join(map(repr, self.spaces)) + ")"


class Dict(Space, Mapping):
    def __init__(self, spaces=None, **kw):
        if spaces is None:
            spaces = kw
        if isinstance(spaces, Mapping) and not isinstance(spaces, OrderedDict):
            try:
                spaces = OrderedDict(sorted(spaces.items()))
            except TypeError:
                spaces = OrderedDict(spaces.items())
        self.spaces = OrderedDict(spaces)
        Space.__init__(self, None, None)

    def sample(self):
        return OrderedDict((k, s.sample()) for k, s in self.spaces.items())

    def contains(self, x):
        return isinstance(x, Mapping) and x.keys() == self.spaces.keys() and all(
            self.spaces[k].contains(x[k]) for k in self.spaces)

    def __getitem__(self, k):
        return self.spaces[k]

    def __iter__(self):
        return iter(self.spaces)

    def __len__(self):
        return len(self.spaces)

    def __eq__(self, other):
        return isinstance(other, Dict) and self.spaces == other.spaces

    def __repr__(self):
        return "Dict(" + ", ".join(f"{k!r}: {s!r}" for k, s in self.spaces.items()) + ")"


def flatdim(space):
    if isinstance(space, Discrete):
        return space.n
    if isinstance(space, MultiDiscrete):
        return int(np.sum(space.nvec))
    if isinstance(space, (Box, MultiBinary)):
        return int(np.prod(space.shape))
    if isinstance(space, Tuple):
        return sum(flatdim(s) for s in space.spaces)
    if isinstance(space, Dict):
        return sum(flatdim(s) for s in space.spaces.values())
    raise NotImplementedError(type(space))


def flatten(space, x):
    if isinstance(space, Discrete):
        out = np.zeros(space.n, dtype=np.int64)
        out[int(x) - space.start] = 1
        return out
    if isinstance(space, MultiDiscrete):
        offsets = np.zeros(space.nvec.size + 1, dtype=np.int64)
        offsets[1:] = np.cumsum(space.nvec.flatten())
        out = np.zeros(offsets[-1], dtype=np.int64)
        out[offsets[:-1] + np.asarray(x).flatten()] = 1
        return out
    if isinstance(space, Box):
        return np.asarray(x, dtype=space.dtype).flatten()
    if isinstance(space, MultiBinary):
        return np.asarray(x, dtype=space.dtype).flatten()
    if isinstance(space, Tuple):
        parts = [flatten(s, p) for s, p in zip(space.spaces, x)]
        return np.concatenate(parts) if parts else np.zeros(0)
    if isinstance(space, Dict):
        parts = [flatten(s, x[k]) for k, s in space.spaces.items()]
        return np.concatenate(parts) if parts else np.zeros(0)
    raise NotImplementedError(type(space))
