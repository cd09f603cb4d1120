"""Minimal observation/action space objects (the reference uses gym.spaces,
which is not a dependency here).  Same attributes the reference code touches:
``low``, ``high``, ``shape``, ``dtype``, ``sample()``; ``Dict`` is a mapping."""
import numpy as np


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float64):
        low = np.asarray(low, dtype=dtype)
        high = np.asarray(high, dtype=dtype)
        if shape is None:
            shape = np.broadcast(low, high).shape
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.low = np.broadcast_to(low, self.shape).copy()
        self.high = np.broadcast_to(high, self.shape).copy()

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"


class Dict(dict):
    """Mapping of sub-spaces (gym.spaces.Dict stand-in)."""

    @property
    def spaces(self):
        return self

    def sample(self):
        return {k: s.sample() for k, s in self.items()}
