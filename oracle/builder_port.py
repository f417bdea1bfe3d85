"""Pure-Python/NumPy restatement of the reference WorldBuilderEnv -- the CPU-baseline "port".

TEST / BASELINE INFRASTRUCTURE.  Same per-step work in the same interpreter as world_builder_env/src/environment/
world_builder_env.py:125-231 and game_logic.py:59-203: resource dicts, an int8 numpy grid searched with np.where on every
successful build, a Dict observation of fresh numpy arrays and an info dict with copies every step.  Validated EXACTLY
against tests/golden/builder_golden.npz.
"""
from __future__ import annotations

import numpy as np

KINDS = ("pass", "farm", "lumberyard", "quarry", "house")
COST = {"farm": {"wood": 5}, "lumberyard": {"stone": 3}, "quarry": {"wood": 5}, "house": {"wood": 10, "stone": 5}}
YIELD = {"farm": ("food", 2), "lumberyard": ("wood", 3), "quarry": ("stone", 2)}
BONUS = {"farm": 3, "lumberyard": 2, "quarry": 2, "house": 4}


class BuilderPort:
    def __init__(self, grid_size=10, randint=None):
        self.G = grid_size
        self.randint = randint if randint is not None else np.random.randint
        self._clear()

    def _clear(self):  # game_logic.py:33-57
        self.board = np.zeros((self.G, self.G), dtype=np.int8)
        self.stock = {"food": 25, "wood": 20, "stone": 10}
        self.people, self.room = 3, 10
        self.built = {"farm": 0, "lumberyard": 0, "quarry": 0, "house": 0}
        self.t = self.held = 0
        self.peaked = False

    def reset(self, seed=None, options=None):  # world_builder_env.py:99-123
        self._clear()
        return self._obs(), self._info()

    def _place(self, kind):  # game_logic.py:125-156
        if any(self.stock[r] < c for r, c in COST[kind].items()):
            return False
        free = np.where(self.board == 0)
        if len(free[0]) == 0:
            return False
        k = self.randint(len(free[0]))
        for r, c in COST[kind].items():
            self.stock[r] -= c
        self.board[free[0][k], free[1][k]] = KINDS.index(kind)
        self.built[kind] += 1
        if kind == "house":
            self.room += 5
        return True

    def step(self, action):  # world_builder_env.py:125-166 + game_logic.py:59-123
        if not (isinstance(action, (int, np.integer)) and 0 <= action < 5):
            raise ValueError(f"Invalid action {action}")
        self.t += 1
        before, room_before, reward = self.people, self.room, 0
        if action:
            kind = KINDS[action]
            if self._place(kind):
                reward += BONUS[kind]
                if kind == "house" and before >= room_before - 1:
                    reward += 10
            else:
                reward -= 3
        for kind, count in self.built.items():  # production
            if count > 0 and kind in YIELD:
                res, amount = YIELD[kind]
                self.stock[res] += amount * count
        if self.stock["food"] < self.people:  # consumption
            self.people = 0
        else:
            self.stock["food"] -= self.people
        if self.people > 0 and self.stock["food"] > 2 and self.people < self.room:  # growth
            self.people += 1
            self.stock["food"] -= 1
        food = self.stock["food"]
        reward += 5 * (self.people > before) - 50 * (self.people < before) + (food > self.people * 2)
        reward -= 2 * (food < self.people) + 5 * (food < max(2, self.people))
        reward += abs(self.stock["wood"] - self.stock["stone"]) < 5
        reward -= action == 1 and food > self.people * 3
        if self.people >= 20 and not self.peaked:
            self.peaked = True
        if self.peaked:
            self.held += 1
        done = self.people <= 0 or (self.peaked and self.held >= 50)
        if done:
            reward = -100 if self.people <= 0 else (100 if self.held >= 50 else -50)
        return self._obs(), int(reward), done, False, self._info()

    def _obs(self):  # :186-217
        return {"grid": self.board.copy(),
                "resources": np.array([self.stock["food"], self.stock["wood"], self.stock["stone"], self.people],
                                      dtype=np.float32),
                "population_capacity": np.array([self.room], dtype=np.float32),
                "win_steps": np.array([self.held], dtype=np.int32)}

    def _info(self):  # :219-231
        return {"steps": self.t, "win_steps": self.held, "reached_win_population": self.peaked,
                "resources": self.stock.copy(), "population": self.people, "population_capacity": self.room,
                "building_counts": self.built.copy()}
