"""ReplayRandom -- replays the engine's counter-based stream into the reference.

TEST INFRASTRUCTURE.  The reference's hot-path envs call the module-level `random`
(snake_env_classic/snake_env.py:4,125-126; crypto_trading_env/crypto_trading_env.py:13,135,...;
traffic_management_env/environment.py:12 and utils.py:6) and `np.random.normal`
(crypto_trading_env.py:148).  Rebinding those names to a ReplayRandom makes the reference
consume exactly the u32 draws the device kernel consumes for env `env_id` (SURVEY.md
section 0 fact 2), so trajectories can be compared bit for bit.
"""
from __future__ import annotations

import math

import numpy as np

from . import philox


class ReplayRandom:
    """Duck-types the subset of `random` the reference uses, backed by draws_u32()."""

    CHUNK = 4096

    def __init__(self, seed: int, env_id: int, stream: int = philox.STREAM_ENV, start: int = 0):
        self._seed = int(seed)
        self._env = int(env_id)
        self._stream = stream
        self.counter = int(start)  # index of the next u32 draw == the kernel's per-env `ctr`
        self._base = -1
        self._buf = None

    # -- raw stream -------------------------------------------------------
    def _u32(self) -> int:
        j = self.counter
        base = j - (j % self.CHUNK)
        if base != self._base:
            self._buf = philox.draws_u32(self._seed, [self._env], base, self.CHUNK, self._stream)[0]
            self._base = base
        self.counter = j + 1
        return int(self._buf[j - base])

    # -- `random` module surface -------------------------------------------
    def seed(self, *_a, **_k):
        """No-op: the stream is positioned by (seed, env_id, counter), never reseeded."""

    def randint(self, a: int, b: int) -> int:
        return a + ((self._u32() * (b - a + 1)) >> 32)

    def random(self) -> float:
        a = self._u32() >> 5
        b = self._u32() >> 6
        return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0)

    def uniform(self, a: float, b: float) -> float:
        return a + (b - a) * self.random()

    def choice(self, seq):
        return seq[self.randint(0, len(seq) - 1)]

    # -- `np.random.normal` replacement -------------------------------------
    def normal(self, loc: float = 0.0, scale: float = 1.0) -> float:
        u1 = self.random()
        u2 = self.random()
        z = math.sqrt(-2.0 * math.log(1.0 - u1)) * math.cos(2.0 * math.pi * u2)
        return loc + scale * z


class NumpyWithReplayNormal:
    """Stands in for the `np` name inside the reference crypto module: everything forwards to numpy except
    `np.random.normal` (crypto_trading_env.py:148) and `np.random.seed` (:306), which go to the replay stream."""

    class _Random:
        def __init__(self, rr):
            self._rr = rr

        def normal(self, loc=0.0, scale=1.0):
            return self._rr.normal(loc, scale)

        def seed(self, *_a, **_k):
            pass

    def __init__(self, rr):
        self.random = self._Random(rr)

    def __getattr__(self, name):
        return getattr(np, name)


class ReplayGenerator:
    """Duck-types the subset of `numpy.random.Generator` that smartclimate uses (`self.rng` in
    smartclimate_rl-main/smartclimate/env.py:31,49-52 and utils.py:5-24), backed by the engine's stream:

      uniform(a, b)      = a + (b - a) * random53()                                  (2 u32)
      integers(lo, hi)   = lo + mulhi(u32, hi - lo)            (hi exclusive)        (1 u32)
      normal(mu, sd)     = Box-Muller on two random53()                              (4 u32)
      choice(a, p=p)     = a[number of cdf entries <= random53()],  cdf = cumsum(p) / cumsum(p)[-1]   (2 u32)
                           -- the inverse-CDF rule numpy's Generator.choice itself uses
    """

    def __init__(self, seed: int, env_id: int, start: int = 0):
        self._rr = ReplayRandom(seed, env_id, start=start)

    @property
    def counter(self) -> int:
        return self._rr.counter

    def uniform(self, low=0.0, high=1.0):
        return low + (high - low) * self._rr.random()

    def integers(self, low, high=None):
        if high is None:
            low, high = 0, low
        return int(low) + ((self._rr._u32() * (int(high) - int(low))) >> 32)

    def normal(self, loc=0.0, scale=1.0):
        return self._rr.normal(loc, scale)

    def random(self):
        return self._rr.random()

    def choice(self, a, p=None):
        if p is None:
            return a[self.integers(0, len(a))]
        cdf = np.asarray(p, dtype=np.float64).cumsum()
        cdf /= cdf[-1]
        return a[int(np.searchsorted(cdf, self._rr.random(), side="right"))]


class NumpyWithReplayRandint:
    """Stands in for the `np` name inside the reference's game_logic module: everything forwards to numpy except
    `np.random.randint(n)` (world_builder_env/src/environment/game_logic.py:137), which becomes one draw of the
    engine's stream: randint(n) = mulhi(u32, n)."""

    class _Random:
        def __init__(self, rr):
            self._rr = rr

        def randint(self, low, high=None):
            if high is None:
                low, high = 0, low
            return self._rr.randint(int(low), int(high) - 1)

        def seed(self, *_a, **_k):
            pass

    def __init__(self, rr):
        self.random = self._Random(rr)

    def __getattr__(self, name):
        return getattr(np, name)
