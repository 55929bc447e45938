"""Tiny stand-ins for gym.spaces (gym is optional): only what the reference's callers read --
`.n`, `.sample()`, `.shape`, `.dtype`, `.low`, `.high` (test_env.py:20, train_dqn.py:107,199)."""
import numpy as np


class Discrete:
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)

    def sample(self):
        return int(np.random.randint(self.n))

    def contains(self, x):
        return 0 <= int(x) < self.n

    def __repr__(self):
        return f'Discrete({self.n})'


class Box:
    def __init__(self, low, high, shape, dtype=np.uint8):
        self.low, self.high = low, high
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)

    def sample(self):
        return np.random.randint(self.low, self.high + 1, size=self.shape).astype(self.dtype)

    def __repr__(self):
        return f'Box({self.low}, {self.high}, {self.shape}, {self.dtype})'
