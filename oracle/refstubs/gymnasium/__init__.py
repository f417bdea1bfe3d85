"""Minimal stand-in for `gymnasium`, used ONLY to import the reference's files in the
build container (gymnasium is not installed in this image).

Test infrastructure: it is put on sys.path by oracle/ref_loader.py and nowhere else.
It implements just what the three hot-path reference modules touch at import time
and in __init__/reset: Env.reset(seed=...), spaces.{Discrete,Box,MultiDiscrete},
register(), envs.registration.register, core.{ObsType,ActType}.
"""
import numpy as _np

from . import spaces  # noqa: F401


class Env:
    metadata = {}
    render_mode = None
    np_random = None

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self.np_random = _np.random.default_rng(seed)

    def close(self):
        pass


_REGISTRY = {}


def register(id, entry_point=None, max_episode_steps=None, **kwargs):
    _REGISTRY[id] = dict(entry_point=entry_point, max_episode_steps=max_episode_steps, kwargs=kwargs)


def make(id, **kwargs):
    import importlib

    spec = _REGISTRY[id]
    mod, cls = spec["entry_point"].split(":")
    kw = dict(spec["kwargs"].get("kwargs", {}))
    kw.update(kwargs)
    return getattr(importlib.import_module(mod), cls)(**kw)
