"""ctypes binding of the plain-C oracle (oracle/c/*.c -> oracle/_c/liboracle.so).

TEST INFRASTRUCTURE.  `build()` compiles it with gcc; __graft_entry__.build() calls that so the
prebuilt .so travels to the GPU box.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_c", "liboracle.so")
_lib = None

AUTORESET = {"disabled": 0, "next_step": 1, "same_step": 2}


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, "c", f) for f in os.listdir(os.path.join(_HERE, "c")) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.run(["make", "-C", os.path.join(_HERE, "c")] + (["-B"] if force else []), check=True,
                       capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.snake_oracle_create.restype = C.c_void_p
        _lib.snake_oracle_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_int]
        _lib.snake_oracle_destroy.argtypes = [C.c_void_p]
        _lib.snake_oracle_reset.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        _lib.snake_oracle_step.restype = C.c_int
        _lib.snake_oracle_step.argtypes = [C.c_void_p] + [C.c_void_p] * 10
        _lib.snake_oracle_get_state.argtypes = [C.c_void_p] + [C.c_void_p] * 8
        _lib.snake_oracle_get_body.restype = C.c_int
        _lib.snake_oracle_get_body.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib.snake_oracle_get_stats.argtypes = [C.c_void_p, C.c_void_p]
        _lib.crypto_oracle_create.restype = C.c_void_p
        _lib.crypto_oracle_create.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_int]
        _lib.crypto_oracle_destroy.argtypes = [C.c_void_p]
        _lib.crypto_oracle_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.crypto_oracle_step.argtypes = [C.c_void_p] + [C.c_void_p] * 11
        _lib.crypto_oracle_get_state.argtypes = [C.c_void_p] + [C.c_void_p] * 8
        _lib.crypto_oracle_set_state.argtypes = [C.c_void_p] + [C.c_void_p] * 8
        _lib.crypto_oracle_get_stats.argtypes = [C.c_void_p, C.c_void_p]
        _lib.traffic_oracle_create.restype = C.c_void_p
        _lib.traffic_oracle_create.argtypes = [C.c_int] * 5 + [C.c_double, C.c_int, C.c_uint64, C.c_uint64, C.c_int]
        _lib.traffic_oracle_destroy.argtypes = [C.c_void_p]
        _lib.traffic_oracle_obs_dim.restype = C.c_int
        _lib.traffic_oracle_obs_dim.argtypes = [C.c_void_p]
        _lib.traffic_oracle_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.traffic_oracle_step.argtypes = [C.c_void_p] + [C.c_void_p] * 8
        _lib.traffic_oracle_get_state.argtypes = [C.c_void_p] + [C.c_void_p] * 10
        _lib.traffic_oracle_get_stats.argtypes = [C.c_void_p, C.c_void_p]
        _lib.climate_oracle_create.restype = C.c_void_p
        _lib.climate_oracle_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_int]
        _lib.climate_oracle_destroy.argtypes = [C.c_void_p]
        _lib.climate_oracle_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.climate_oracle_step.argtypes = [C.c_void_p] + [C.c_void_p] * 12
        _lib.climate_oracle_get_state.argtypes = [C.c_void_p] + [C.c_void_p] * 8
        _lib.climate_oracle_get_stats.argtypes = [C.c_void_p, C.c_void_p]
        _lib.builder_oracle_create.restype = C.c_void_p
        _lib.builder_oracle_create.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_int]
        _lib.builder_oracle_destroy.argtypes = [C.c_void_p]
        _lib.builder_oracle_reset.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        _lib.builder_oracle_step.restype = C.c_int
        _lib.builder_oracle_step.argtypes = [C.c_void_p] + [C.c_void_p] * 10
        _lib.builder_oracle_get_state.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        _lib.builder_oracle_get_stats.argtypes = [C.c_void_p, C.c_void_p]
        _lib.beng_oracle_draws_u32.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        _lib.beng_oracle_action_tape.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_uint32, C.c_int, C.c_int,
                                                 C.c_void_p]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def draws_u32(seed, env, stream, first, count):
    out = np.empty(count, np.uint32)
    lib().beng_oracle_draws_u32(seed, env, stream, first, count, _p(out))
    return out


def action_tape(seed, env_id_base, n_envs, step, n_choices, n_cols=1):
    out = np.empty((n_envs, n_cols), np.int64)
    lib().beng_oracle_action_tape(seed, env_id_base, n_envs, step, n_choices, n_cols, _p(out))
    return out[:, 0] if n_cols == 1 else out


class SnakeOracle:
    """Batched CPU oracle with the same outputs as the device engine's snake step."""

    def __init__(self, n_envs, grid_size=20, max_steps=1000, seed=0, env_id_base=0, autoreset="same_step"):
        self.n, self.G = int(n_envs), int(grid_size)
        self._h = C.c_void_p(lib().snake_oracle_create(self.n, self.G, max_steps, seed, env_id_base,
                                                        AUTORESET[autoreset]))
        n, G = self.n, self.G
        self.obs = np.zeros((n, G, G), np.int8)
        self.reward = np.zeros(n, np.float32)
        self.terminated = np.zeros(n, np.uint8)
        self.truncated = np.zeros(n, np.uint8)
        self.score = np.zeros(n, np.int32)
        self.length = np.zeros(n, np.int32)
        self.ep_return = np.zeros(n, np.float32)
        self.ep_length = np.zeros(n, np.int32)
        self.ep_score = np.zeros(n, np.int32)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().snake_oracle_destroy(self._h)
            self._h = None

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().snake_oracle_reset(self._h, _p(m), _p(self.obs), _p(self.score), _p(self.length))
        return self.obs

    def step(self, actions, want_obs=True):
        a = np.ascontiguousarray(actions, np.int64)
        assert a.shape == (self.n,)
        self.invalid = lib().snake_oracle_step(
            self._h, _p(a), _p(self.obs) if want_obs else None, _p(self.reward), _p(self.terminated),
            _p(self.truncated), _p(self.score), _p(self.length), _p(self.ep_return), _p(self.ep_length),
            _p(self.ep_score))
        return self.obs, self.reward, self.terminated, self.truncated

    def state(self):
        names = ["head_r", "head_c", "food_r", "food_c", "direction", "steps", "length"]
        arrs = [np.zeros(self.n, np.int32) for _ in names]
        ctr = np.zeros(self.n, np.uint32)
        lib().snake_oracle_get_state(self._h, *[_p(a) for a in arrs], _p(ctr))
        d = dict(zip(names, arrs))
        d["rng_counter"] = ctr
        return d

    def body(self, env):
        out = np.zeros(self.G * self.G + 1, np.int32)
        n = lib().snake_oracle_get_body(self._h, env, _p(out))
        return out[:n].copy()

    def stats(self):
        out = np.zeros(5, np.int64)
        lib().snake_oracle_get_stats(self._h, _p(out))
        return dict(zip(["n_episodes", "sum_return", "sum_length", "sum_score", "max_score"], out.tolist()))


CRYPTO_DEFAULT_CFG = (10000.0, 0.001, 0.0005, 100.0, 100000.0, 0.02, 0.1)  # TradingConfig, crypto_trading_env.py:28-38
CRYPTO_OBS_DIM = 261
CRYPTO_HIST = 50


class CryptoOracle:
    """Batched CPU oracle (float64) with the same outputs as the device engine's crypto step."""

    def __init__(self, n_envs, seed=0, env_id_base=0, autoreset="same_step", action_type="discrete",
                 cfg=CRYPTO_DEFAULT_CFG, max_steps=1000):
        self.n = int(n_envs)
        self.continuous = action_type == "continuous"
        cfg_arr = np.asarray(cfg, np.float64)
        self._h = C.c_void_p(lib().crypto_oracle_create(self.n, seed, env_id_base, AUTORESET[autoreset],
                                                         int(self.continuous), _p(cfg_arr), max_steps))
        n = self.n
        self.obs = np.zeros((n, CRYPTO_OBS_DIM), np.float32)
        self.reward = np.zeros(n, np.float32)
        self.reward64 = np.zeros(n, np.float64)
        self.terminated = np.zeros(n, np.uint8)
        self.truncated = np.zeros(n, np.uint8)
        self.portfolio_value = np.zeros(n, np.float64)
        self.current_price = np.zeros(n, np.float64)
        self.trade_kind = np.zeros(n, np.uint8)
        self.ep_return = np.zeros(n, np.float64)
        self.ep_length = np.zeros(n, np.int32)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().crypto_oracle_destroy(self._h)
            self._h = None

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().crypto_oracle_reset(self._h, _p(m), _p(self.obs))
        return self.obs

    def step(self, actions, want_obs=True):
        if self.continuous:
            a = np.ascontiguousarray(actions, np.float32)
            assert a.shape == (self.n, 2)
        else:
            a = np.ascontiguousarray(actions, np.int64)
            assert a.shape == (self.n,)
        lib().crypto_oracle_step(self._h, _p(a), _p(self.obs) if want_obs else None, _p(self.reward),
                                 _p(self.terminated), _p(self.truncated), _p(self.reward64),
                                 _p(self.portfolio_value), _p(self.current_price), _p(self.trade_kind),
                                 _p(self.ep_return), _p(self.ep_length))
        return self.obs, self.reward, self.terminated, self.truncated

    def state(self, with_candles=False):
        n = self.n
        d = {"cash": np.zeros(n), "holdings": np.zeros(n), "trend_strength": np.zeros(n), "psychology": np.zeros(n),
             "regime": np.zeros(n, np.int32), "step": np.zeros(n, np.int32), "rng_counter": np.zeros(n, np.uint32)}
        candles = np.zeros((n, CRYPTO_HIST, 5)) if with_candles else None
        lib().crypto_oracle_get_state(self._h, *[_p(d[k]) for k in ("cash", "holdings", "trend_strength",
                                      "psychology", "regime", "step", "rng_counter")], _p(candles))
        if with_candles:
            d["candles"] = candles
        return d

    def set_state(self, d):
        arrs = [np.ascontiguousarray(d[k], dt) for k, dt in (("cash", np.float64), ("holdings", np.float64),
                ("trend_strength", np.float64), ("psychology", np.float64), ("regime", np.int32),
                ("step", np.int32), ("rng_counter", np.uint32))]
        candles = np.ascontiguousarray(d["candles"], np.float64) if "candles" in d else None
        lib().crypto_oracle_set_state(self._h, *[_p(a) for a in arrs], _p(candles))

    def stats(self):
        out = np.zeros(4, np.float64)
        lib().crypto_oracle_get_stats(self._h, _p(out))
        return dict(zip(["n_episodes", "sum_return", "sum_length", "sum_final_value"], out.tolist()))


class TrafficOracle:
    """Batched CPU oracle with the same outputs as the device engine's traffic step."""

    def __init__(self, n_envs, grid_size=(5, 5), num_intersections=9, max_vehicles=50, spawn_rate=0.3, seed=0,
                 env_id_base=0, autoreset="same_step", max_timesteps=1000):
        self.n = int(n_envs)
        self._h = C.c_void_p(lib().traffic_oracle_create(self.n, grid_size[0], grid_size[1], num_intersections,
                                                          max_vehicles, spawn_rate, max_timesteps, seed, env_id_base,
                                                          AUTORESET[autoreset]))
        self.ni = min(num_intersections, grid_size[0] * grid_size[1])
        self.obs_dim = lib().traffic_oracle_obs_dim(self._h)
        n = self.n
        self.obs = np.zeros((n, self.obs_dim), np.float32)
        self.reward = np.zeros(n, np.float32)
        self.reward64 = np.zeros(n, np.float64)
        self.terminated = np.zeros(n, np.uint8)
        self.truncated = np.zeros(n, np.uint8)
        self.ep_return = np.zeros(n, np.float64)
        self.ep_length = np.zeros(n, np.int32)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().traffic_oracle_destroy(self._h)
            self._h = None

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().traffic_oracle_reset(self._h, _p(m), _p(self.obs))
        return self.obs

    def step(self, actions, want_obs=True):
        a = np.ascontiguousarray(actions, np.int64)
        assert a.shape == (self.n, self.ni)
        lib().traffic_oracle_step(self._h, _p(a), _p(self.obs) if want_obs else None, _p(self.reward),
                                  _p(self.terminated), _p(self.truncated), _p(self.reward64), _p(self.ep_return),
                                  _p(self.ep_length))
        return self.obs, self.reward, self.terminated, self.truncated

    def state(self):
        n, ni = self.n, self.ni
        d = {"timestep": np.zeros(n, np.int32), "num_vehicles": np.zeros(n, np.int32),
             "rng_counter": np.zeros(n, np.uint32), "total_reward": np.zeros(n, np.float64),
             "phase": np.zeros((n, ni), np.int32), "timer": np.zeros((n, ni), np.int32),
             "passed": np.zeros((n, ni), np.int32), "waiting": np.zeros((n, ni), np.int32),
             "qlen": np.zeros((n, ni, 4), np.int32), "qwait": np.zeros((n, ni, 4), np.int32)}
        lib().traffic_oracle_get_state(self._h, *[_p(d[k]) for k in ("timestep", "num_vehicles", "rng_counter",
                                       "total_reward", "phase", "timer", "passed", "waiting", "qlen", "qwait")])
        return d

    def stats(self):
        out = np.zeros(3, np.float64)
        lib().traffic_oracle_get_stats(self._h, _p(out))
        return dict(zip(["n_episodes", "sum_return", "sum_length"], out.tolist()))


class ClimateOracle:
    """Batched CPU oracle (float64) with the same outputs as the device engine's smartclimate step."""

    def __init__(self, n_envs, max_occupancy=8, episode_minutes=1440, seed=0, env_id_base=0, autoreset="same_step"):
        self.n = int(n_envs)
        self._h = C.c_void_p(lib().climate_oracle_create(self.n, max_occupancy, episode_minutes, seed, env_id_base,
                                                          AUTORESET[autoreset]))
        n = self.n
        self.obs = np.zeros((n, 9), np.float32)
        self.reward = np.zeros(n, np.float32)
        self.reward64 = np.zeros(n, np.float64)
        self.terminated = np.zeros(n, np.uint8)
        self.truncated = np.zeros(n, np.uint8)
        self.comfort = np.zeros(n, np.float64)
        self.ac_penalty = np.zeros(n, np.float64)
        self.light_penalty = np.zeros(n, np.float64)
        self.ep_return = np.zeros(n, np.float64)
        self.ep_length = np.zeros(n, np.int32)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().climate_oracle_destroy(self._h)
            self._h = None

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().climate_oracle_reset(self._h, _p(m), _p(self.obs))
        return self.obs

    def step(self, ac_temp, lights):
        a = np.ascontiguousarray(ac_temp, np.float32).reshape(self.n)
        l = np.ascontiguousarray(lights, np.int8).reshape(self.n, 4)
        lib().climate_oracle_step(self._h, _p(a), _p(l), _p(self.obs), _p(self.reward), _p(self.terminated),
                                  _p(self.truncated), _p(self.reward64), _p(self.comfort), _p(self.ac_penalty),
                                  _p(self.light_penalty), _p(self.ep_return), _p(self.ep_length))
        return self.obs, self.reward, self.terminated, self.truncated

    def state(self):
        n = self.n
        d = {"room_temp": np.zeros(n), "outside_temp": np.zeros(n), "total_reward": np.zeros(n),
             "energy_usage": np.zeros(n), "num_people": np.zeros(n, np.int32), "current_step": np.zeros(n, np.int32),
             "comfort_time": np.zeros(n, np.int32), "rng_counter": np.zeros(n, np.uint32)}
        lib().climate_oracle_get_state(self._h, *[_p(d[k]) for k in ("room_temp", "outside_temp", "total_reward",
                                       "energy_usage", "num_people", "current_step", "comfort_time", "rng_counter")])
        return d

    def stats(self):
        out = np.zeros(3, np.float64)
        lib().climate_oracle_get_stats(self._h, _p(out))
        return dict(zip(["n_episodes", "sum_return", "sum_length"], out.tolist()))


class BuilderOracle:
    """Batched CPU oracle with the same outputs as the device engine's world-builder step."""

    def __init__(self, n_envs, grid_size=10, seed=0, env_id_base=0, autoreset="same_step"):
        self.n, self.G = int(n_envs), int(grid_size)
        self._h = C.c_void_p(lib().builder_oracle_create(self.n, self.G, seed, env_id_base, AUTORESET[autoreset]))
        n, G = self.n, self.G
        self.grid = np.zeros((n, G, G), np.int8)
        self.resources = np.zeros((n, 4), np.float32)
        self.capacity = np.zeros((n, 1), np.float32)
        self.win_steps = np.zeros((n, 1), np.int32)
        self.reward = np.zeros(n, np.float32)
        self.terminated = np.zeros(n, np.uint8)
        self.truncated = np.zeros(n, np.uint8)
        self.ep_return = np.zeros(n, np.int32)
        self.ep_length = np.zeros(n, np.int32)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().builder_oracle_destroy(self._h)
            self._h = None

    def obs(self):
        return {"grid": self.grid, "resources": self.resources, "population_capacity": self.capacity,
                "win_steps": self.win_steps}

    def flat_obs(self):
        return np.concatenate([self.grid.reshape(self.n, -1).astype(np.float32), self.resources, self.capacity,
                               self.win_steps.astype(np.float32)], axis=1)

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().builder_oracle_reset(self._h, _p(m), _p(self.grid), _p(self.resources), _p(self.capacity),
                                   _p(self.win_steps))
        return self.obs()

    def step(self, actions):
        a = np.ascontiguousarray(actions, np.int64)
        assert a.shape == (self.n,)
        self.invalid = lib().builder_oracle_step(self._h, _p(a), _p(self.grid), _p(self.resources), _p(self.capacity),
                                                 _p(self.win_steps), _p(self.reward), _p(self.terminated),
                                                 _p(self.truncated), _p(self.ep_return), _p(self.ep_length))
        return self.obs(), self.reward, self.terminated, self.truncated

    def state(self):
        n = self.n
        d = {"steps": np.zeros(n, np.int32), "reached": np.zeros(n, np.int32), "rng_counter": np.zeros(n, np.uint32),
             "building_counts": np.zeros((n, 4), np.int32)}
        lib().builder_oracle_get_state(self._h, _p(d["steps"]), _p(d["reached"]), _p(d["rng_counter"]),
                                       _p(d["building_counts"]))
        return d

    def stats(self):
        out = np.zeros(4, np.int64)
        lib().builder_oracle_get_stats(self._h, _p(out))
        return dict(zip(["n_episodes", "sum_return", "sum_length", "wins"], out.tolist()))
