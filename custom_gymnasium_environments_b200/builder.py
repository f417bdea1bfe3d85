"""world_builder_env on the B200 engine.

  BatchedWorldBuilderEnv   gymnasium.vector.VectorEnv-compatible; N envs stepped by ONE CUDA kernel
                           (csrc/builder.cu) through the C ABI (include/beng.h).
  WorldBuilderEnv          the reference's single-instance gym.Env surface
                           (world_builder_env/src/environment/world_builder_env.py:10-247), a 1-env view of the engine.

The reference's observation is a Dict {'grid' int8 (G,G), 'resources' float32 (4,), 'population_capacity' float32 (1,),
'win_steps' int32 (1,)} or, with flatten_obs=True, one float32 vector of G*G + 6 values (:71-86, :186-217); both forms are
kept.  There is no time limit: an episode ends by starvation (-100) or after 50 steps at population >= 20 (+100);
`truncated` is always False.  Integer dynamics, bit-exact against the reference.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .spaces import Box, Dict, Discrete, batch_space
from .vector import _EnvBase, AUTORESET_MODES, _VectorEnvBase, _mode_name, as_device_actions, host_source, require_cuda, stream_ptr

BUILDING_NAMES = ("farm", "lumberyard", "quarry", "house")
BUILDER_STAT_NAMES = ("n_episodes", "sum_return", "sum_length", "wins")
MAX_POPULATION, WIN_STEPS = 20, 50


class BatchedWorldBuilderEnv(_VectorEnvBase):
    """N independent WorldBuilderEnv instances; scalar state as [word][env] int32, the grid doubles as observation."""

    metadata = {"render_modes": [], "render_fps": 4, "autoreset_mode": "same_step"}

    def __init__(self, num_envs: int, grid_size: int = 10, render_mode=None, flatten_obs: bool = False, *,
                 device="cuda", seed: int = 0, env_id_base: int = 0, autoreset_mode="same_step",
                 debug_checks: bool = False):
        self.lib = _lib.load()
        self.device = require_cuda(device)
        self.num_envs = n = int(num_envs)
        self.grid_size = G = int(grid_size)
        self.flatten_obs = bool(flatten_obs)
        self.render_mode = render_mode
        self.debug_checks = bool(debug_checks)
        self.autoreset_mode = _mode_name(autoreset_mode)
        self.metadata = dict(type(self).metadata, autoreset_mode=self.autoreset_mode)
        self.closed = False

        self.single_action_space = Discrete(5)                                            # :60
        if self.flatten_obs:
            self.single_observation_space = Box(0.0, 1000.0, (G * G + 6,), np.float32)    # :71-78
        else:
            self.single_observation_space = Dict({                                        # :80-85
                "grid": Box(0, 4, (G, G), np.int8), "resources": Box(0.0, 1000.0, (4,), np.float32),
                "population_capacity": Box(0.0, 100.0, (1,), np.float32), "win_steps": Box(0, WIN_STEPS, (1,), np.int32)})
        self.action_space = batch_space(self.single_action_space, n)
        self.observation_space = batch_space(self.single_observation_space, n)

        self.params = _lib.BuilderParams(G, AUTORESET_MODES[self.autoreset_mode], int(seed), int(env_id_base))
        dev = self.device
        with torch.cuda.device(dev):
            z = lambda *shape, dt: torch.zeros(shape, dtype=dt, device=dev)  # noqa: E731
            self._words = z(10, n, dt=torch.int32)
            self.grid = z(n, G, G, dt=torch.int8)
            self.resources = z(n, 4, dt=torch.float32)
            self.population_capacity = z(n, 1, dt=torch.float32)
            self.win_steps = z(n, 1, dt=torch.int32)
            self.flat_obs = z(n, G * G + 6, dt=torch.float32) if self.flatten_obs else None
            self.reward = z(n, dt=torch.float32)
            self.terminated = z(n, dt=torch.bool)
            self.truncated = z(n, dt=torch.bool)
            self.ep_return = z(n, dt=torch.int32)
            self.ep_length = z(n, dt=torch.int32)
            self.stats = z(4, dt=torch.int64)
            self.invalid_count = z(1, dt=torch.int32)
            self._actions = z(n, dt=torch.int64)
        self._state = _lib.BuilderState(self._words.data_ptr())
        self._io = _lib.BuilderIO(self.grid.data_ptr(), self.resources.data_ptr(), self.population_capacity.data_ptr(),
                                  self.win_steps.data_ptr(), self.flat_obs.data_ptr() if self.flatten_obs else None,
                                  self.reward.data_ptr(), self.terminated.data_ptr(), self.truncated.data_ptr(),
                                  self.ep_return.data_ptr(), self.ep_length.data_ptr(), self.stats.data_ptr(),
                                  self.invalid_count.data_ptr())
        self._host = None
        self._needs_first_reset = True

    # ------------------------------------------------------------------ state views
    food = property(lambda self: self._words[0])
    wood = property(lambda self: self._words[1])
    stone = property(lambda self: self._words[2])
    population = property(lambda self: self._words[3])
    steps = property(lambda self: self._words[6])
    reached_win_population = property(lambda self: ((self._words[7] >> 16) & 1).bool())
    rng_counter = property(lambda self: self._words[8].to(torch.int64) & 0xFFFFFFFF)

    @property
    def building_counts(self):
        """(n, 4) counts of farm, lumberyard, quarry, house."""
        c = self._words[5]
        return torch.stack([(c >> (8 * b)) & 0xFF for b in range(4)], dim=1)

    def _obs(self):
        if self.flatten_obs:
            return self.flat_obs
        return {"grid": self.grid, "resources": self.resources, "population_capacity": self.population_capacity,
                "win_steps": self.win_steps}

    def _infos(self):
        return {"steps": self._words[6], "win_steps": self.win_steps[:, 0], "population": self._words[3],
                "population_capacity": self._words[4],
                "episode": {"r": self.ep_return, "l": self.ep_length}, "_episode": self.terminated}

    # ------------------------------------------------------------------ VectorEnv API
    def reset(self, *, seed=None, options=None):
        """WorldBuilderEnv.reset for every env (:99-123) -> (obs, infos); `seed` re-keys and rewinds the stream."""
        first = self._needs_first_reset
        if seed is not None:
            self.params.seed = int(seed)
            first = True
        mask = None if not options else options.get("reset_mask")
        mask_ptr = None
        if mask is not None:
            if self._needs_first_reset:
                raise RuntimeError("the first reset() must reset every env")
            mask = torch.as_tensor(mask).to(device=self.device, dtype=torch.uint8).contiguous()
            if mask.shape != (self.num_envs,):
                raise ValueError("reset_mask must have shape (num_envs,)")
            mask_ptr = mask.data_ptr()
        with torch.cuda.device(self.device):
            rc = self.lib.beng_builder_reset(C.byref(self.params), C.byref(self._state), C.byref(self._io), mask_ptr,
                                             self.num_envs, int(first), stream_ptr(self.device))
        _lib.check(rc, "beng_builder_reset")
        self._needs_first_reset = False
        return self._obs(), self._infos()

    def step(self, actions):
        """One step of every env (:125-166).  actions: int64 (n,), 0 pass / 1 farm / 2 lumberyard / 3 quarry / 4 house."""
        if self._needs_first_reset:
            raise RuntimeError("call reset() before step()")
        act = as_device_actions(actions, self._actions)
        with torch.cuda.device(self.device):
            rc = self.lib.beng_builder_step(C.byref(self.params), C.byref(self._state), act.data_ptr(),
                                            C.byref(self._io), self.num_envs, stream_ptr(self.device))
        _lib.check(rc, "beng_builder_step")
        if self.debug_checks:
            n_bad = int(self.invalid_count.item())
            if n_bad:
                self.invalid_count.zero_()
                raise ValueError(f"Invalid action: {n_bad} action(s) outside Discrete(5)")  # :135-136
        return self._obs(), self.reward, self.terminated, self.truncated, self._infos()

    def step_host(self, actions, *, copy_obs: bool = True, sync: bool = True):
        """step() for callers holding HOST arrays (numpy in, numpy out) through `beng_builder_step_host`."""
        if self._needs_first_reset:
            raise RuntimeError("call reset() before step()")
        if self._host is None:
            n, G = self.num_envs, self.grid_size
            pin = dict(pin_memory=True)
            self._host = {"actions": torch.zeros(n, dtype=torch.int64, **pin),
                          "grid": torch.zeros((n, G, G), dtype=torch.int8, **pin),
                          "resources": torch.zeros((n, 4), dtype=torch.float32, **pin),
                          "capacity": torch.zeros((n, 1), dtype=torch.float32, **pin),
                          "win_steps": torch.zeros((n, 1), dtype=torch.int32, **pin),
                          "reward": torch.zeros(n, dtype=torch.float32, **pin),
                          "terminated": torch.zeros(n, dtype=torch.bool, **pin)}
        h = self._host
        src = actions if isinstance(actions, torch.Tensor) else torch.as_tensor(np.asarray(actions))
        src = self._host_src = host_source(src, h["actions"])
        with torch.cuda.device(self.device):
            rc = self.lib.beng_builder_step_host(
                C.byref(self.params), C.byref(self._state), self._actions.data_ptr(), C.byref(self._io),
                self.num_envs, src.data_ptr(), h["grid"].data_ptr() if copy_obs else None,
                h["resources"].data_ptr(), h["capacity"].data_ptr(), h["win_steps"].data_ptr(),
                h["reward"].data_ptr(), h["terminated"].data_ptr(), stream_ptr(self.device))
            _lib.check(rc, "beng_builder_step_host")
            if sync:
                torch.cuda.current_stream(self.device).synchronize()
        obs = {"grid": h["grid"].numpy() if copy_obs else self.grid, "resources": h["resources"].numpy(),
               "population_capacity": h["capacity"].numpy(), "win_steps": h["win_steps"].numpy()}
        return obs, h["reward"].numpy(), h["terminated"].numpy(), np.zeros(self.num_envs, dtype=bool), {}

    def episode_stats(self) -> dict:
        return dict(zip(BUILDER_STAT_NAMES, self.stats.tolist()))

    def state_dict(self) -> dict:
        return {"words": self._words.clone(), "grid": self.grid.clone(), "stats": self.stats.clone(),
                "seed": int(self.params.seed), "env_id_base": int(self.params.env_id_base)}

    def load_state_dict(self, sd: dict):
        self._words.copy_(sd["words"])
        self.grid.copy_(sd["grid"])
        self.stats.copy_(sd["stats"])
        self.params.seed, self.params.env_id_base = int(sd["seed"]), int(sd["env_id_base"])
        self._needs_first_reset = False

    def render(self):
        return None  # pygame rendering is out of scope

    def close(self, **kwargs):
        self.closed = True


class WorldBuilderEnv(_EnvBase):
    """Single-instance gym.Env surface of the reference (world_builder_env.py:10-247) on the CUDA engine: a 1-env
    BatchedWorldBuilderEnv with auto-reset disabled; numpy observations, Python numbers, the reference's info keys."""

    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 4}

    def __init__(self, grid_size: int = 10, render_mode=None, flatten_obs: bool = False, *, device="cuda",
                 seed: int = 0, env_id: int = 0):
        self.grid_size, self.render_mode, self.flatten_obs = grid_size, render_mode, flatten_obs
        self.MAX_POPULATION, self.WIN_STEPS = MAX_POPULATION, WIN_STEPS
        self._vec = BatchedWorldBuilderEnv(1, grid_size, flatten_obs=flatten_obs, device=device, seed=seed,
                                           env_id_base=env_id, autoreset_mode="disabled")
        self.action_space = self._vec.single_action_space
        self.observation_space = self._vec.single_observation_space

    def _obs(self):
        v = self._vec
        if self.flatten_obs:
            return v.flat_obs[0].cpu().numpy().copy()
        return {"grid": v.grid[0].cpu().numpy().copy(), "resources": v.resources[0].cpu().numpy().copy(),
                "population_capacity": v.population_capacity[0].cpu().numpy().copy(),
                "win_steps": v.win_steps[0].cpu().numpy().copy()}

    def _get_info(self):  # :219-231
        v = self._vec
        w = v._words[:, 0].tolist()
        return {"steps": w[6], "win_steps": w[7] & 0xFFFF, "reached_win_population": bool((w[7] >> 16) & 1),
                "resources": {"food": w[0], "wood": w[1], "stone": w[2]}, "population": w[3],
                "population_capacity": w[4],
                "building_counts": {name: (w[5] >> (8 * b)) & 0xFF for b, name in enumerate(BUILDING_NAMES)}}

    def reset(self, seed=None, options=None):
        self._vec.reset(seed=seed)
        return self._obs(), self._get_info()

    def step(self, action):
        if not self.action_space.contains(action):
            raise ValueError(f"Invalid action {action}. Action space is {self.action_space}")  # :135-136
        _, rew, term, _, _ = self._vec.step(np.array([action], dtype=np.int64))
        return self._obs(), int(rew.item()), bool(term.item()), False, self._get_info()

    def render(self):
        return None

    def close(self):
        self._vec.close()
