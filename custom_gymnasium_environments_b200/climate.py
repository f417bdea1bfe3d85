"""smartclimate on the B200 engine.

  BatchedSmartClimateEnv   gymnasium.vector.VectorEnv-compatible; N envs stepped by ONE CUDA kernel
                           (csrc/climate.cu) through the C ABI (include/beng.h).
  SmartClimateEnv          the reference's single-instance gym.Env surface
                           (smartclimate_rl-main/smartclimate/env.py:10-128), a 1-env view of the same engine.

The reference's Dict action {'ac_temp': Box(16, 32, (1,)), 'lights': MultiBinary(4)} (env.py:37-41) is kept: batched
actions are {'ac_temp': float32 (n, 1) or (n,), 'lights': int8 (n, 4)}.  The 1440-minute limit is reported as
`terminated` (env.py:107); arithmetic is float64 on the device like the reference's Python floats.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .spaces import Box, Dict, MultiBinary, batch_space
from .vector import _EnvBase, AUTORESET_MODES, _VectorEnvBase, _mode_name, host_source, require_cuda, stream_ptr

OBS_DIM = 9
CLIMATE_STAT_NAMES = ("n_episodes", "sum_return", "sum_length")


class BatchedSmartClimateEnv(_VectorEnvBase):
    """N independent SmartClimateEnv instances; state resident in HBM as [field][env] arrays."""

    metadata = {"render_modes": [], "render_fps": 10, "autoreset_mode": "same_step"}

    def __init__(self, num_envs: int, max_occupancy: int = 8, comfort_temp_range=(20.0, 24.0),
                 episode_minutes: int = 1440, *, device="cuda", seed: int = 0, env_id_base: int = 0,
                 autoreset_mode="same_step", time_limit_truncation: bool = False, render_mode=None):
        self.lib = _lib.load()
        self.device = require_cuda(device)
        self.num_envs = n = int(num_envs)
        self.max_occupancy = int(max_occupancy)
        self.comfort_temp_range = tuple(comfort_temp_range)  # kept for API parity; unused by the reference's dynamics
        self.episode_minutes = int(episode_minutes)
        self.render_mode = render_mode
        self.autoreset_mode = _mode_name(autoreset_mode)
        self.metadata = dict(type(self).metadata, autoreset_mode=self.autoreset_mode)
        self.closed = False

        self.single_action_space = Dict({"ac_temp": Box(16.0, 32.0, (1,), np.float32), "lights": MultiBinary(4)})
        self.single_observation_space = Box(
            np.array([0.0, 0, 0.0, 10.0, 16.0, 0, 0, 0, 0]),
            np.array([50.0, self.max_occupancy, 23.99, 50.0, 32.0, 1, 1, 1, 1]), dtype=np.float32)   # env.py:42-46
        self.action_space = batch_space(self.single_action_space, n)
        self.observation_space = batch_space(self.single_observation_space, n)

        self.params = _lib.ClimateParams(self.max_occupancy, self.episode_minutes,
                                         AUTORESET_MODES[self.autoreset_mode], int(bool(time_limit_truncation)),
                                         int(seed), int(env_id_base))
        dev = self.device
        with torch.cuda.device(dev):
            z = lambda *shape, dt: torch.zeros(shape, dtype=dt, device=dev)  # noqa: E731
            self._f64 = z(5, n, dt=torch.float64)   # room_temp, outside_temp, ac_setting, total_reward, energy_usage
            self._i32 = z(4, n, dt=torch.int32)     # people|lights<<8|flags<<16 ; step ; comfort_time ; rng counter
            self.obs = z(n, OBS_DIM, dt=torch.float32)
            self.reward = z(n, dt=torch.float32)
            self.terminated = z(n, dt=torch.bool)
            self.truncated = z(n, dt=torch.bool)
            self.reward64 = z(n, dt=torch.float64)
            self.reward_terms = z(3, n, dt=torch.float64)
            self.ep_return = z(n, dt=torch.float64)
            self.ep_length = z(n, dt=torch.int32)
            self.stats = z(3, dt=torch.float64)
            self._ac = z(n, dt=torch.float32)
            self._lights = z(n, 4, dt=torch.int8)
        self._state = _lib.ClimateState(self._f64.data_ptr(), self._i32.data_ptr())
        self._io = _lib.ClimateIO(self.obs.data_ptr(), self.reward.data_ptr(), self.terminated.data_ptr(),
                                  self.truncated.data_ptr(), self.reward64.data_ptr(), self.reward_terms.data_ptr(),
                                  self.ep_return.data_ptr(), self.ep_length.data_ptr(), self.stats.data_ptr())
        self._host = None
        self._needs_first_reset = True

    # ------------------------------------------------------------------ state views
    room_temp = property(lambda self: self._f64[0])
    outside_temp = property(lambda self: self._f64[1])
    ac_setting = property(lambda self: self._f64[2])
    total_reward = property(lambda self: self._f64[3])
    energy_usage = property(lambda self: self._f64[4])
    num_people = property(lambda self: self._i32[0] & 0xFF)
    current_step = property(lambda self: self._i32[1])
    comfort_time = property(lambda self: self._i32[2])
    rng_counter = property(lambda self: self._i32[3].to(torch.int64) & 0xFFFFFFFF)

    def _infos(self):
        return {"comfort": self.reward_terms[0], "ac_penalty": self.reward_terms[1],
                "light_penalty": self.reward_terms[2], "comfort_time": self._i32[2], "energy_usage": self._f64[4],
                "step": self._i32[1], "reward64": self.reward64,
                "episode": {"r": self.ep_return, "l": self.ep_length}, "_episode": self.terminated}

    # ------------------------------------------------------------------ VectorEnv API
    def reset(self, *, seed=None, options=None):
        """SmartClimateEnv.reset for every env (env.py:62-70) -> (obs, infos); `seed` re-keys and rewinds the stream."""
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
            rc = self.lib.beng_climate_reset(C.byref(self.params), C.byref(self._state), C.byref(self._io), mask_ptr,
                                             self.num_envs, int(first), stream_ptr(self.device))
        _lib.check(rc, "beng_climate_reset")
        self._needs_first_reset = False
        return self.obs, {}

    def _split_action(self, actions):
        if isinstance(actions, dict):
            ac, lights = actions["ac_temp"], actions["lights"]
        else:
            ac, lights = actions
        return ac, lights

    def _to_device(self, value, buf):
        if isinstance(value, torch.Tensor):
            if value.device == buf.device and value.dtype == buf.dtype and value.is_contiguous() \
                    and value.numel() == buf.numel() and value.data_ptr() % 16 == 0:  # (a misaligned view goes through buf)
                return value
            buf.copy_(value.reshape(buf.shape), non_blocking=True)
            return buf
        arr = np.asarray(value)
        if arr.size != buf.numel():
            raise ValueError(f"action component must have {buf.numel()} elements, got shape {arr.shape}")
        buf.copy_(torch.from_numpy(np.ascontiguousarray(arr).reshape(tuple(buf.shape))).to(buf.dtype))
        return buf

    def step(self, actions):
        """One step of every env (env.py:84-117).  actions = {'ac_temp': float32 (n,1)|(n,), 'lights': int8 (n,4)}."""
        if self._needs_first_reset:
            raise RuntimeError("call reset() before step()")
        ac, lights = self._split_action(actions)
        ac_t, li_t = self._to_device(ac, self._ac), self._to_device(lights, self._lights)
        with torch.cuda.device(self.device):
            rc = self.lib.beng_climate_step(C.byref(self.params), C.byref(self._state), ac_t.data_ptr(),
                                            li_t.data_ptr(), C.byref(self._io), self.num_envs,
                                            stream_ptr(self.device))
        _lib.check(rc, "beng_climate_step")
        return self.obs, self.reward, self.terminated, self.truncated, self._infos()

    def step_host(self, actions, *, copy_obs: bool = True, sync: bool = True):
        """step() for callers holding HOST arrays (numpy in, numpy out) through `beng_climate_step_host`."""
        if self._needs_first_reset:
            raise RuntimeError("call reset() before step()")
        if self._host is None:
            n = self.num_envs
            pin = dict(pin_memory=True)
            self._host = {"ac": torch.zeros(n, dtype=torch.float32, **pin),
                          "lights": torch.zeros((n, 4), dtype=torch.int8, **pin),
                          "obs": torch.zeros((n, OBS_DIM), dtype=torch.float32, **pin),
                          "reward": torch.zeros(n, dtype=torch.float32, **pin),
                          "terminated": torch.zeros(n, dtype=torch.bool, **pin),
                          "truncated": torch.zeros(n, dtype=torch.bool, **pin)}
        h = self._host
        ac, lights = self._split_action(actions)
        ac_src = ac if isinstance(ac, torch.Tensor) else torch.as_tensor(np.asarray(ac, dtype=np.float32))
        li_src = lights if isinstance(lights, torch.Tensor) else torch.as_tensor(np.asarray(lights, dtype=np.int8))
        ac_src = host_source(ac_src, h["ac"])
        li_src = host_source(li_src, h["lights"])
        self._host_src = (ac_src, li_src)
        with torch.cuda.device(self.device):
            rc = self.lib.beng_climate_step_host(
                C.byref(self.params), C.byref(self._state), self._ac.data_ptr(), self._lights.data_ptr(),
                C.byref(self._io), self.num_envs, ac_src.data_ptr(), li_src.data_ptr(),
                h["obs"].data_ptr() if copy_obs else None, h["reward"].data_ptr(), h["terminated"].data_ptr(),
                h["truncated"].data_ptr(), stream_ptr(self.device))
            _lib.check(rc, "beng_climate_step_host")
            if sync:
                torch.cuda.current_stream(self.device).synchronize()
        obs = h["obs"].numpy() if copy_obs else self.obs
        return obs, h["reward"].numpy(), h["terminated"].numpy(), h["truncated"].numpy(), {}

    def episode_stats(self) -> dict:
        return dict(zip(CLIMATE_STAT_NAMES, self.stats.tolist()))

    def state_dict(self) -> dict:
        return {"f64": self._f64.clone(), "i32": self._i32.clone(), "stats": self.stats.clone(),
                "seed": int(self.params.seed), "env_id_base": int(self.params.env_id_base)}

    def load_state_dict(self, sd: dict):
        self._f64.copy_(sd["f64"])
        self._i32.copy_(sd["i32"])
        self.stats.copy_(sd["stats"])
        self.params.seed, self.params.env_id_base = int(sd["seed"]), int(sd["env_id_base"])
        self._needs_first_reset = False

    def render(self):
        return None  # matplotlib visualiser is out of scope

    def close(self, **kwargs):
        self.closed = True


class SmartClimateEnv(_EnvBase):
    """Single-instance gym.Env surface of the reference (env.py:10-128) on the CUDA engine: a 1-env
    BatchedSmartClimateEnv with auto-reset disabled; numpy observations, Python floats, the reference's info keys."""

    metadata = {"render_modes": ["human"], "render_fps": 10}

    def __init__(self, max_occupancy: int = 8, comfort_temp_range=(20.0, 24.0), episode_minutes: int = 1440,
                 seed=None, log_level=None, *, device="cuda", env_id: int = 0, **kwargs):
        self.max_occupancy, self.comfort_temp_range, self.episode_minutes = max_occupancy, comfort_temp_range, episode_minutes
        self._vec = BatchedSmartClimateEnv(1, max_occupancy, comfort_temp_range, episode_minutes, device=device,
                                           seed=0 if seed is None else int(seed), env_id_base=env_id,
                                           autoreset_mode="disabled")
        self.action_space = self._vec.single_action_space
        self.observation_space = self._vec.single_observation_space

    room_temp = property(lambda self: float(self._vec.room_temp.item()))
    num_people = property(lambda self: int(self._vec.num_people.item()))
    current_step = property(lambda self: int(self._vec.current_step.item()))

    def reset(self, *, seed=None, options=None):
        obs, _ = self._vec.reset(seed=seed)
        return obs[0].cpu().numpy().copy(), {}

    def step(self, action):
        act = {"ac_temp": np.asarray(action["ac_temp"], dtype=np.float32).reshape(1),
               "lights": np.asarray(action["lights"], dtype=np.int8).reshape(1, 4)}
        obs, rew, term, trunc, info = self._vec.step(act)
        out = {"comfort": float(info["comfort"].item()), "ac_penalty": float(info["ac_penalty"].item()),
               "light_penalty": float(info["light_penalty"].item()), "comfort_time": int(info["comfort_time"].item()),
               "energy_usage": float(info["energy_usage"].item()), "step": int(info["step"].item())}
        return obs[0].cpu().numpy().copy(), float(info["reward64"].item()), bool(term.item()), False, out

    def render(self, mode: str = "human"):
        return None

    def close(self):
        self._vec.close()
