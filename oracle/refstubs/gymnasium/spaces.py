"""Stub spaces (see gymnasium/__init__.py in this directory)."""
import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None):
        self.shape = shape
        self.dtype = np.dtype(dtype) if dtype is not None else None


class Discrete(Space):
    def __init__(self, n, start=0):
        super().__init__((), np.int64)
        self.n = int(n)
        self.start = int(start)

    def contains(self, x):
        if isinstance(x, (int, np.integer)):
            v = int(x)
        elif isinstance(x, np.ndarray) and x.shape == () and np.issubdtype(x.dtype, np.integer):
            v = int(x)
        else:
            return False
        return self.start <= v < self.start + self.n

    def sample(self):
        return int(np.random.randint(self.start, self.start + self.n))


class MultiDiscrete(Space):
    def __init__(self, nvec, dtype=np.int64):
        self.nvec = np.asarray(nvec, dtype=dtype)
        super().__init__(self.nvec.shape, dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.nvec.shape and bool(((x >= 0) & (x < self.nvec)).all())

    def sample(self):
        return (np.random.random(self.nvec.shape) * self.nvec).astype(self.dtype)


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.asarray(low).shape
        super().__init__(tuple(shape), dtype)
        self.low = np.broadcast_to(np.asarray(low, dtype=np.float64), self.shape)
        self.high = np.broadcast_to(np.asarray(high, dtype=np.float64), self.shape)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(((x >= self.low) & (x <= self.high)).all())

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return (lo + (hi - lo) * np.random.random(self.shape)).astype(self.dtype)


class Dict(Space):
    def __init__(self, spaces=None, **kw):
        super().__init__(None, None)
        self.spaces = dict(spaces or {}, **kw)


class MultiBinary(Space):
    def __init__(self, n):
        super().__init__((int(n),), np.int8)
        self.n = int(n)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(((x == 0) | (x == 1)).all())

    def sample(self):
        return (np.random.random(self.shape) < 0.5).astype(np.int8)
