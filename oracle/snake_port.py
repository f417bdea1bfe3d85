"""Pure-Python restatement of the reference SnakeEnvClassic -- the CPU-baseline "port".

TEST / BASELINE INFRASTRUCTURE.  `/root/reference` cannot travel to the GPU box, so the
"reference's pure-Python per-env step loop" that BASELINE.json's north_star asks to be timed
beside the GPU number is this port: same per-step work in the same interpreter (a Python list
body searched with `in`, a fresh numpy observation every step, the global-`random`-style
rejection loop), written from the behaviour in snake_env_classic/snake_env.py:49-143.
It is validated against the real reference in tests/test_snake_oracle.py (build container) and
against tests/golden/snake_*.npz (everywhere).
"""
from __future__ import annotations

import random as _global_random

import numpy as np

_DELTA = ((-1, 0), (0, 1), (1, 0), (0, -1))  # up, right, down, left as (drow, dcol); snake_env.py:77-85


class SnakePort:
    def __init__(self, grid_size: int = 20, rng=None, max_steps: int = 1000):
        self.G = grid_size
        self.max_steps = max_steps  # snake_env.py:47
        self.rng = rng if rng is not None else _global_random
        self.body = None
        self.food = None
        self.heading = None
        self.score = 0
        self.steps = 0

    # snake_env.py:121-129
    def _new_food(self):
        hi = self.G - 1
        draw = self.rng.randint
        while True:
            cand = (draw(0, hi), draw(0, hi))
            if cand not in self.body:
                self.food = cand
                return

    # snake_env.py:131-143
    def _grid(self):
        g = np.zeros((self.G, self.G), dtype=np.int8)
        for r, c in self.body:
            g[r, c] = 1
        if self.food:
            g[self.food[0], self.food[1]] = 2
        return g

    # snake_env.py:49-65
    def reset(self, seed=None, options=None):
        mid = self.G // 2
        self.body = [(mid, mid)]
        self.heading = 1
        self.score = 0
        self.steps = 0
        self._new_food()
        return self._grid(), {"score": 0, "snake_length": 1}

    # snake_env.py:67-119
    def step(self, action):
        if not (isinstance(action, (int, np.integer)) and 0 <= action < 4):
            raise ValueError(f"Invalid action: {action}")
        if abs(action - self.heading) != 2:
            self.heading = int(action)
        dr, dc = _DELTA[self.heading]
        nxt = (self.body[0][0] + dr, self.body[0][1] + dc)
        if not (0 <= nxt[0] < self.G and 0 <= nxt[1] < self.G) or nxt in self.body:
            return self._grid(), -10.0, True, False, {"score": self.score}
        self.body.insert(0, nxt)
        reward = 0
        if nxt == self.food:
            self.score += 1
            reward = 10.0
            self._new_food()
        else:
            self.body.pop()
        self.steps += 1
        return (self._grid(), reward, self.steps >= self.max_steps, False,
                {"score": self.score, "snake_length": len(self.body)})
