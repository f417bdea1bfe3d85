"""Pure-Python restatement of the reference TrafficManagementEnv -- the CPU-baseline "port".

TEST / BASELINE INFRASTRUCTURE.  `/root/reference` cannot travel to the GPU box, so the "reference's pure-Python
per-env step loop" timed beside the GPU number is this port.  It keeps the reference's per-step WORK, including the
part that is a tautology (every vehicle is tested against every intersection position with a sqrt distance,
environment.py:251-269 / utils.py:63-68 -- 34 % of the reference's step time) and the repeated queue-length helpers
in reward / observation / metrics, so that its speed is representative.  Validated EXACTLY against
tests/golden/traffic_golden.npz (tests/test_traffic_oracle.py).
"""
from __future__ import annotations

import random as _py_random

import numpy as np

PHASES = ("NS_GREEN", "NS_YELLOW", "EW_GREEN", "EW_YELLOW")  # config.py:17
N, E, S, W = 0, 1, 2, 3                                       # utils.py:16-21


class _Car:
    __slots__ = ("pos", "heading", "goal", "waited")

    def __init__(self, pos, heading, goal):
        self.pos, self.heading, self.goal, self.waited = pos, heading, goal, 0

    def near(self, p, tol=5.0):  # utils.py:63-68 (np.sqrt on Python floats, like the reference)
        return np.sqrt((self.pos[0] - p[0]) ** 2 + (self.pos[1] - p[1]) ** 2) <= tol


class _Node:
    def __init__(self, idx, pos):
        self.idx, self.pos = idx, pos
        self.clear()

    def clear(self):  # fresh TrafficLight (utils.py:75-77) + empty queues
        self.phase, self.timer = "NS_GREEN", 0
        self.lanes = {N: [], E: [], S: [], W: []}
        self.passed = 0
        self.waited = 0

    def lane_sizes(self):
        return {d: len(q) for d, q in self.lanes.items()}

    def load(self):
        return sum(len(q) for q in self.lanes.values())


class TrafficPort:
    def __init__(self, grid_size=(5, 5), num_intersections=9, max_vehicles=50, spawn_rate=0.3, rng=None,
                 max_timesteps=1000):
        self.rows, self.cols = grid_size
        self.ni = min(num_intersections, self.rows * self.cols)
        self.cap, self.rate, self.limit = max_vehicles, spawn_rate, max_timesteps
        self.rng = rng if rng is not None else _py_random
        self.nodes = [_Node(i, ((i % self.cols) * 100.0, (i // self.cols) * 100.0)) for i in range(self.ni)]
        self.cars, self.t, self.total = [], 0, 0.0

    def reset(self, seed=None, options=None):  # environment.py:141-166
        self.cars, self.t, self.total = [], 0, 0.0
        for n in self.nodes:
            n.clear()
        return self._observe(), self._info()

    def _around(self, i):  # utils.py:196-214, order N, S, W, E
        r, c = divmod(i, self.cols)
        out = []
        for dr, dc in ((-1, 0), (1, 0), (0, -1), (0, 1)):
            rr, cc = r + dr, c + dc
            if 0 <= rr < self.rows and 0 <= cc < self.cols:
                out.append(rr * self.cols + cc)
        return out

    def _heading(self, a, b):  # utils.py:230-248
        ar, ac = divmod(a, self.cols)
        br, bc = divmod(b, self.cols)
        if br < ar:
            return N
        if br > ar:
            return S
        return W if bc < ac else E

    def step(self, action):  # environment.py:168-203
        self.t += 1
        for i, a in enumerate(action):  # _apply_actions
            if i >= self.ni:
                break
            n = self.nodes[i]
            if a == 1 and n.phase != "NS_GREEN":
                n.phase, n.timer = "NS_GREEN", 5
            elif a == 2 and n.phase != "EW_GREEN":
                n.phase, n.timer = "EW_GREEN", 5
        for n in self.nodes:  # TrafficLight.update
            n.timer -= 1
            if n.timer <= 0:
                n.phase = PHASES[(PHASES.index(n.phase) + 1) % 4]
                n.timer = 3 if "YELLOW" in n.phase else self.rng.randint(5, 30)
        if len(self.cars) < self.cap and self.rng.random() < self.rate:  # _spawn_vehicles
            start = self.rng.randint(0, self.ni - 1)
            path = [start]
            for _ in range(self.rng.randint(2, min(5, self.ni)) - 1):
                path.append(self.rng.choice(self._around(path[-1])))
            if len(path) > 1:
                car = _Car(self.nodes[start].pos, self._heading(path[0], path[1]), path[-1])
                self.nodes[start].lanes[car.heading].append(car)
                self.cars.append(car)
        for car in self.cars[:]:  # _update_vehicles: the reference's (tautological) position checks
            if not any(car.near(n.pos) for n in self.nodes):
                raise AssertionError("unreachable: vehicles never leave their start intersection")
        for n in self.nodes:  # _process_intersections
            for d in (N, E, S, W):
                lane = n.lanes[d]
                if not lane:
                    continue
                go = (n.phase == "NS_GREEN" and d in (N, S)) or (n.phase == "EW_GREEN" and d in (E, W))
                if go:
                    while lane:
                        car = lane.pop(0)
                        n.passed += 1
                        if car.goal is not None and car.goal == n.idx:
                            car.goal = None
                else:
                    for car in lane:
                        car.waited += 1
                        n.waited += 1
        self.cars = [c for c in self.cars if c.goal is not None]  # _remove_completed_vehicles
        reward = 0.0  # _calculate_reward
        reward += sum(n.passed for n in self.nodes) * 1.0
        reward += sum(n.waited for n in self.nodes) * -0.1
        reward += sum(n.load() for n in self.nodes) * -0.05
        loads = [n.load() for n in self.nodes]
        if len(loads) > 1:
            reward += 0.5 / (1 + np.var(loads))
        self.total += reward
        return self._observe(), reward, self.t >= self.limit, False, self._info()

    def _metrics(self):  # utils.py:251-267
        passed = sum(n.passed for n in self.nodes)
        waited = sum(n.waited for n in self.nodes)
        queued = sum(n.load() for n in self.nodes)
        return {"total_vehicles_passed": passed, "total_waiting_time": waited,
                "average_waiting_time": waited / max(passed, 1), "total_queue_length": queued,
                "average_queue_length": queued / len(self.nodes), "throughput": passed / len(self.nodes)}

    def _info(self):  # environment.py:365-384 (built on every step and reset, like the reference)
        return {"timestep": self.t, "num_vehicles": len(self.cars), "total_reward": self.total,
                "metrics": self._metrics(),
                "intersection_states": [{"id": n.idx, "light_phase": n.phase, "queue_lengths": n.lane_sizes(),
                                         "vehicles_passed": n.passed, "total_waiting_time": n.waited}
                                        for n in self.nodes]}

    def _observe(self):  # environment.py:313-363
        out = []
        for n in self.nodes:
            out.extend(1 if n.phase == p else 0 for p in PHASES)
        for n in self.nodes:
            sizes = n.lane_sizes()
            out.extend(min(sizes[d], 20) for d in (N, E, S, W))
        for n in self.nodes:
            for d in (N, E, S, W):
                lane = n.lanes[d]
                out.append(min(sum(c.waited for c in lane) / len(lane), 100) if lane else 0)
        for n in self.nodes:
            out.append(n.passed)
            out.append(min(n.waited, 1000))
        m = self._metrics()
        out.extend([len(self.cars), min(m["average_waiting_time"], 100), min(m["average_queue_length"], 50),
                    m["throughput"]])
        return np.array(out, dtype=np.float32)
