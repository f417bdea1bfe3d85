"""Observation / action spaces.

gymnasium is not installed in this image, so the package must not
hard-depend on it: when `gymnasium` imports, its real space classes are used; otherwise these
minimal classes provide the attributes callers of the reference touch
(`.n`, `.shape`, `.dtype`, `.low`, `.high`, `.nvec`, `.contains()`, `.sample()`).
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the environment
    from gymnasium.spaces import Box, Dict, Discrete, MultiBinary, MultiDiscrete  # type: ignore

    HAVE_GYMNASIUM = True
except Exception:  # gymnasium absent
    HAVE_GYMNASIUM = False

    class _Space:
        def __init__(self, shape, dtype):
            self.shape = tuple(shape)
            self.dtype = np.dtype(dtype)
            self._rng = np.random.default_rng()

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

        def __repr__(self):
            return f"{type(self).__name__}(shape={self.shape}, dtype={self.dtype})"

    class Discrete(_Space):
        def __init__(self, n, start=0):
            super().__init__((), np.int64)
            self.n, self.start = int(n), int(start)

        def contains(self, x):
            if isinstance(x, (int, np.integer)) and not isinstance(x, bool):
                v = int(x)
            elif isinstance(x, np.ndarray) and x.shape == () and np.issubdtype(x.dtype, np.integer):
                v = int(x)
            else:
                return False
            return self.start <= v < self.start + self.n

        def sample(self):
            return int(self._rng.integers(self.start, self.start + self.n))

    class MultiDiscrete(_Space):
        def __init__(self, nvec, dtype=np.int64):
            self.nvec = np.asarray(nvec, dtype=dtype)
            super().__init__(self.nvec.shape, dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.nvec.shape and bool(((x >= 0) & (x < self.nvec)).all())

        def sample(self):
            return (self._rng.random(self.nvec.shape) * self.nvec).astype(self.dtype)

    class Box(_Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            if shape is None:
                shape = np.asarray(low).shape
            super().__init__(shape, dtype)
            self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape)
            self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(((x >= self.low) & (x <= self.high)).all())

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1.0)
            hi = np.where(np.isfinite(self.high), self.high, 1.0)
            return (lo + (hi - lo) * self._rng.random(self.shape)).astype(self.dtype)


if not HAVE_GYMNASIUM:

    class MultiBinary(_Space):
        def __init__(self, n):
            super().__init__((int(n),), np.int8)
            self.n = int(n)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(((x == 0) | (x == 1)).all())

        def sample(self):
            return (self._rng.random(self.shape) < 0.5).astype(np.int8)

    class Dict:
        """Minimal dictionary space (name -> space)."""

        def __init__(self, spaces=None, **kw):
            self.spaces = dict(spaces or {}, **kw)

        def __getitem__(self, key):
            return self.spaces[key]

        def keys(self):
            return self.spaces.keys()

        def contains(self, x):
            return isinstance(x, dict) and set(x) == set(self.spaces) and all(
                sp.contains(x[k]) for k, sp in self.spaces.items())

        def sample(self):
            return {k: sp.sample() for k, sp in self.spaces.items()}

        def __repr__(self):
            return f"Dict({self.spaces!r})"


def batch_space(space, n: int):
    """The batched counterpart of a single-env space (gymnasium.vector.utils.batch_space)."""
    if isinstance(space, Discrete):
        return MultiDiscrete(np.full((n,), space.n, dtype=np.int64))
    if isinstance(space, MultiDiscrete):
        return MultiDiscrete(np.broadcast_to(space.nvec, (n,) + space.nvec.shape).copy())
    if isinstance(space, Box):
        return Box(np.broadcast_to(space.low, (n,) + space.shape).copy(),
                   np.broadcast_to(space.high, (n,) + space.shape).copy(), dtype=space.dtype)
    if isinstance(space, MultiBinary):
        return Box(0, 1, (n,) + tuple(space.shape), np.int8)
    if isinstance(space, Dict):
        return Dict({k: batch_space(sp, n) for k, sp in space.spaces.items()})
    raise TypeError(f"cannot batch {space!r}")
