"""Minimal stand-ins for the Gymnasium types the engine touches, used ONLY when ``gymnasium`` is not importable
(e.g. a GPU box without the host framework installed).  With Gymnasium present the real classes are used and the
engine is a genuine ``gymnasium.vector.VectorEnv`` subclass (see _api.py).

Written from the API contract (gymnasium/vector/vector_env.py:34-351, gymnasium/spaces/{box,discrete,multi_discrete}.py),
not copied: only the attributes/methods the engine and its tests need.
"""
from __future__ import annotations

from enum import Enum

import numpy as np


class AutoresetMode(Enum):
    NEXT_STEP = "NextStep"
    SAME_STEP = "SameStep"
    DISABLED = "Disabled"


class Space:
    def __init__(self, shape, dtype, seed=None):
        self.shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)
        self._rng = None
        if seed is not None:
            self.seed(seed)

    @property
    def np_random(self):
        if self._rng is None:
            self.seed()
        return self._rng

    def seed(self, seed=None):
        self._rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        return seed


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        low, high = np.asarray(low, dtype=dtype), np.asarray(high, dtype=dtype)
        if shape is None:
            shape = np.broadcast(low, high).shape
        self.low = np.broadcast_to(low, shape).astype(dtype)
        self.high = np.broadcast_to(high, shape).astype(dtype)
        super().__init__(shape, dtype, seed)

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return self.np_random.uniform(lo, hi, size=self.shape).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


class Discrete(Space):
    def __init__(self, n, seed=None, start=0):
        self.n, self.start = int(n), int(start)
        super().__init__((), np.int64, seed)

    def sample(self):
        return np.int64(self.start + self.np_random.integers(self.n))

    def contains(self, x):
        return np.ndim(x) == 0 and self.start <= int(x) < self.start + self.n

    def __repr__(self):
        return f"Discrete({self.n})"


class MultiDiscrete(Space):
    def __init__(self, nvec, dtype=np.int64, seed=None, start=None):
        self.nvec = np.asarray(nvec, dtype=dtype)
        self.start = np.zeros_like(self.nvec) if start is None else np.asarray(start, dtype=dtype)
        super().__init__(self.nvec.shape, dtype, seed)

    def sample(self):
        return (self.np_random.random(self.nvec.shape) * self.nvec).astype(self.dtype) + self.start

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.start) and np.all(x < self.start + self.nvec))

    def __repr__(self):
        return f"MultiDiscrete({self.nvec})"


class Tuple(Space):
    def __init__(self, spaces, seed=None):
        self.spaces = tuple(spaces)
        super().__init__(None, None, seed)

    def sample(self):
        return tuple(s.sample() for s in self.spaces)

    def contains(self, x):
        return isinstance(x, (tuple, list)) and len(x) == len(self.spaces) and all(s.contains(v) for s, v in zip(self.spaces, x))

    def __len__(self):
        return len(self.spaces)

    def __getitem__(self, i):
        return self.spaces[i]

    def __repr__(self):
        return "Tuple(" + ", ".join(map(repr, self.spaces)) + ")"


def batch_space(space, n=1):
    if isinstance(space, Tuple):
        return Tuple(tuple(batch_space(s, n) for s in space.spaces))
    if isinstance(space, Box):
        reps = (n,) + (1,) * space.low.ndim
        return Box(np.tile(space.low, reps), np.tile(space.high, reps), dtype=space.dtype)
    if isinstance(space, Discrete):
        return MultiDiscrete(np.full((n,), space.n, dtype=np.int64), start=np.full((n,), space.start, dtype=np.int64))
    raise TypeError(f"cannot batch {space!r}")


class VectorEnv:
    metadata: dict = {}
    spec = None
    render_mode = None
    closed = False
    num_envs: int

    def reset(self, *, seed=None, options=None):
        raise NotImplementedError

    def step(self, actions):
        raise NotImplementedError

    def render(self):
        raise NotImplementedError(f"{self} render function is not implemented.")

    def close(self, **kwargs):
        if self.closed:
            return
        self.close_extras(**kwargs)
        self.closed = True

    def close_extras(self, **kwargs):
        pass

    @property
    def unwrapped(self):
        return self

    def __del__(self):
        if not getattr(self, "closed", True):
            self.close()

    def __repr__(self):
        if self.spec is None:
            return f"{self.__class__.__name__}(num_envs={self.num_envs})"
        return f"{self.__class__.__name__}({self.spec.id}, num_envs={self.num_envs})"
