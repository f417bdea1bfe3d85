"""Pure-Python/NumPy restatement of the reference SmartClimateEnv -- the CPU-baseline "port".

TEST / BASELINE INFRASTRUCTURE.  Same per-step work in the same interpreter as smartclimate_rl-main/smartclimate/
env.py:84-117 and utils.py:5-50: a numpy Generator (`normal`, `choice` with probabilities), np.clip on scalars, a fresh
float32 observation array and an info dict every step.  Validated EXACTLY against tests/golden/climate_golden.npz.
"""
from __future__ import annotations

import math

import numpy as np


class ClimatePort:
    def __init__(self, max_occupancy=8, episode_minutes=1440, rng=None, seed=None):
        self.cap, self.minutes = max_occupancy, episode_minutes
        self.rng = rng if rng is not None else np.random.default_rng(seed)
        self.t = 0
        self._fresh()

    def _outside(self, hour):  # utils.py:5-13
        base = 25 if 0 <= hour < 8 else (45 if 8 <= hour < 16 else 35)
        return float(self.rng.normal(base, 5))

    def _fresh(self):  # env.py:48-60
        self.room = self.rng.uniform(22.0, 26.0)
        self.people = self.rng.integers(0, self.cap + 1)
        self.hour = 0.0
        self.outside = self._outside(self.hour)
        self.ac = 24.0
        self.lamps = np.zeros(4, dtype=np.int8)
        self.ret, self.comfy, self.energy = 0.0, 0, 0.0

    def reset(self, *, seed=None, options=None):  # env.py:62-70
        if seed is not None:
            self.rng = np.random.default_rng(seed)
        self.t = 0
        self._fresh()
        return self._obs(), {}

    def _obs(self):  # env.py:72-82
        return np.array([self.room, self.people, self.hour, self.outside, self.ac, *self.lamps], dtype=np.float32)

    def step(self, action):  # env.py:84-117
        self.ac = float(np.clip(action["ac_temp"][0], 16.0, 32.0))
        self.lamps = np.array(action["lights"], dtype=np.int8)
        self.t += 1
        self.hour = (self.t % 1440) / 60.0
        self.outside = self._outside(self.hour)
        if 9 <= self.hour < 18:  # utils.py:15-22
            delta = self.rng.choice([-1, 0, 1, 2], p=[0.1, 0.3, 0.4, 0.2])
        else:
            delta = self.rng.choice([-2, -1, 0, 1], p=[0.2, 0.4, 0.3, 0.1])
        self.people = int(np.clip(self.people + delta, 0, self.cap))
        nxt = self.room + 0.1 * (self.outside - self.room) + 0.2 * (self.ac - self.room) + self.people * 1.0
        self.room = float(np.clip(nxt, 10, 50))  # utils.py:24-28
        obs = self._obs()
        if 20 <= self.room <= 24:  # utils.py:30-50
            comfort = 10
        elif 18 <= self.room <= 26:
            comfort = 5
        elif 16 <= self.room <= 28:
            comfort = 0
        else:
            comfort = -15 * abs(self.room - 22)
        ac_pen = -0.5 * abs(self.ac - self.outside)
        on = int(np.sum(self.lamps))
        lamp_pen = -1 * max(0, on - min(4, math.ceil(self.people / 2)))
        reward = comfort + ac_pen + lamp_pen
        self.ret += reward
        if 20 <= self.room <= 24:
            self.comfy += 1
        self.energy += abs(self.ac - self.outside) + np.sum(self.lamps)
        info = {"comfort": comfort, "ac_penalty": ac_pen, "light_penalty": lamp_pen, "comfort_time": self.comfy,
                "energy_usage": self.energy, "step": self.t}
        return obs, reward, self.t >= self.minutes, False, info
