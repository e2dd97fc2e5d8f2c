"""Observation/action spaces.  gymnasium's are used when gymnasium is importable; otherwise a minimal, API-compatible
subset (Discrete, MultiDiscrete, Box, Tuple with sample/contains/==) so that the package does not hard-depend on it
(the reference does: rware/warehouse.py:5-6, 255-260, 505-522)."""
import numpy as np

try:  # pragma: no cover - depends on the environment
    import gymnasium as _gym
    from gymnasium.spaces import Box, Dict, Discrete, MultiBinary, MultiDiscrete, Tuple  # noqa: F401
    HAVE_GYMNASIUM = True
except Exception:  # gymnasium absent
    _gym = None
    HAVE_GYMNASIUM = False

    class _Space:
        _rng = None

        @property
        def np_random(self):
            if self._rng is None:
                self._rng = np.random.default_rng()
            return self._rng

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

        def __contains__(self, x):
            return self.contains(x)

    class Discrete(_Space):
        def __init__(self, n):
            self.n = int(n)
            self.shape = ()
            self.dtype = np.dtype(np.int64)

        def sample(self):
            return int(self.np_random.integers(self.n))

        def contains(self, x):
            try:
                return 0 <= int(x) < self.n and float(x) == int(x)
            except (TypeError, ValueError):
                return False

        def __eq__(self, other):
            return isinstance(other, Discrete) and other.n == self.n

        def __repr__(self):
            return f"Discrete({self.n})"

    class MultiDiscrete(_Space):
        def __init__(self, nvec):
            self.nvec = np.asarray(nvec, dtype=np.int64)
            self.shape = self.nvec.shape
            self.dtype = np.dtype(np.int64)

        def sample(self):
            return (self.np_random.random(self.nvec.shape) * self.nvec).astype(np.int64)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= 0) and np.all(x < self.nvec))

        def __eq__(self, other):
            return isinstance(other, MultiDiscrete) and self.shape == other.shape and bool(np.all(self.nvec == other.nvec))

        def __repr__(self):
            return f"MultiDiscrete({self.nvec.tolist()})"

    class Box(_Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.dtype = np.dtype(dtype)
            self.shape = tuple(shape) if shape is not None else np.broadcast(np.asarray(low), np.asarray(high)).shape
            self.low = np.broadcast_to(np.asarray(low, dtype=np.float64), self.shape).astype(self.dtype)
            self.high = np.broadcast_to(np.asarray(high, dtype=np.float64), self.shape).astype(self.dtype)

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1e6)
            hi = np.where(np.isfinite(self.high), self.high, 1e6)
            return self.np_random.uniform(lo, hi).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __eq__(self, other):
            return isinstance(other, Box) and self.shape == other.shape and self.dtype == other.dtype

        def __repr__(self):
            return f"Box({self.shape}, {self.dtype})"

    class MultiBinary(_Space):
        def __init__(self, n):
            self.n = n
            self.shape = (n,) if isinstance(n, (int, np.integer)) else tuple(n)
            self.dtype = np.dtype(np.int8)

        def sample(self):
            return self.np_random.integers(0, 2, size=self.shape, dtype=np.int8)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all((x == 0) | (x == 1)))

        def __eq__(self, other):
            return isinstance(other, MultiBinary) and self.shape == other.shape

        def __repr__(self):
            return f"MultiBinary({self.n})"

    class Dict(_Space):
        def __init__(self, spaces):
            from collections import OrderedDict
            self.spaces = OrderedDict(spaces if isinstance(spaces, OrderedDict) else sorted(dict(spaces).items()))

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}

        def contains(self, x):
            return set(x.keys()) == set(self.spaces.keys()) and all(s.contains(x[k]) for k, s in self.spaces.items())

        def __getitem__(self, k):
            return self.spaces[k]

        def keys(self):
            return self.spaces.keys()

        def __iter__(self):
            return iter(self.spaces)

        def __len__(self):
            return len(self.spaces)

        def __eq__(self, other):
            return isinstance(other, Dict) and self.spaces == other.spaces

        def __repr__(self):
            return "Dict(" + ", ".join(f"{k!r}: {v!r}" for k, v in self.spaces.items()) + ")"

    class Tuple(_Space):
        def __init__(self, spaces):
            self.spaces = tuple(spaces)

        def sample(self):
            return tuple(s.sample() for s in self.spaces)

        def contains(self, x):
            return len(x) == len(self.spaces) and all(s.contains(v) for s, v in zip(self.spaces, x))

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


def flatdim(space):
    """gymnasium.spaces.flatdim for the spaces this package uses."""
    if HAVE_GYMNASIUM:
        return _gym.spaces.flatdim(space)
    if isinstance(space, Discrete):
        return int(space.n)
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
    """gymnasium.spaces.flatten: Discrete / MultiDiscrete become one-hot vectors, Box and MultiBinary are raveled, Tuple and
    Dict concatenate their members (Dict in the key order of the space)."""
    if HAVE_GYMNASIUM:
        return _gym.spaces.flatten(space, x)
    if isinstance(space, Discrete):
        out = np.zeros(space.n, dtype=np.int64)
        out[int(x)] = 1
        return out
    if isinstance(space, MultiDiscrete):
        nvec = np.asarray(space.nvec).ravel()
        offsets = np.concatenate([[0], np.cumsum(nvec)])
        out = np.zeros(int(offsets[-1]), dtype=np.int64)
        out[offsets[:-1] + np.asarray(x).ravel()] = 1
        return out
    if isinstance(space, (Box, MultiBinary)):
        return np.asarray(x, dtype=space.dtype).ravel()
    if isinstance(space, Tuple):
        parts = [flatten(s, p) for s, p in zip(space.spaces, x)]
        return np.concatenate(parts) if parts else np.zeros(0)
    if isinstance(space, Dict):
        parts = [flatten(s, x[k]) for k, s in space.spaces.items()]
        return np.concatenate(parts) if parts else np.zeros(0)
    raise NotImplementedError(type(space))


class MultiAgentObservationSpace(list):
    """rware/utils/spaces.py:4-21: a list of per-agent spaces with sample() / contains()."""

    def __init__(self, ma_space):
        for x in ma_space:
            assert hasattr(x, "sample") and hasattr(x, "contains")
        super().__init__(ma_space)

    def sample(self):
        return [sa.sample() for sa in self]

    def contains(self, obs):
        return all(space.contains(ob) for space, ob in zip(self, obs))


class MultiAgentActionSpace(list):
    """rware/utils/spaces.py:24-32."""

    def __init__(self, ma_space):
        for x in ma_space:
            assert hasattr(x, "sample") and hasattr(x, "contains")
        super().__init__(ma_space)

    def sample(self):
        return [sa.sample() for sa in self]

