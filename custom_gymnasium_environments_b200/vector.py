"""Shared host-side plumbing of the batched (VectorEnv-compatible) classes.

`gymnasium.vector.VectorEnv` is duck-typed (gymnasium is not installed in this image): the
batched classes expose num_envs, single_observation_space, single_action_space,
observation_space, action_space, metadata["autoreset_mode"], reset(seed=, options=),
step(actions), close().  When gymnasium IS importable they also subclass the real VectorEnv.
"""
from __future__ import annotations

import numpy as np
import torch

try:  # pragma: no cover
    from gymnasium.vector import VectorEnv as _VectorEnvBase  # type: ignore
except Exception:
    class _VectorEnvBase:  # minimal stand-in
        metadata = {}
        render_mode = None
        closed = False

        def close(self, **kwargs):
            self.closed = True

try:  # pragma: no cover
    from gymnasium import Env as _EnvBase  # type: ignore  (gym.make refuses entry points that are not gymnasium.Env)
except Exception:
    class _EnvBase:  # minimal stand-in for the single-instance facades
        metadata = {}
        render_mode = None

AUTORESET_MODES = {"disabled": 0, "next_step": 1, "same_step": 2}


def _mode_name(mode) -> str:
    name = getattr(mode, "value", mode)
    name = str(name).lower()
    if name not in AUTORESET_MODES:
        raise ValueError(f"autoreset_mode must be one of {sorted(AUTORESET_MODES)}, got {mode!r}")
    return name


def require_cuda(device) -> torch.device:
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("the batched engine runs on CUDA devices only (there is no CPU fallback)")
    if not torch.cuda.is_available():
        raise RuntimeError("CUDA is not available: the batched engine has no CPU fallback")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def as_device_actions(actions, buf: torch.Tensor) -> torch.Tensor:
    """Return a contiguous int64 CUDA tensor shaped like `buf` holding `actions` (zero-copy when possible)."""
    if isinstance(actions, torch.Tensor):
        if actions.device == buf.device and actions.dtype == torch.int64 and actions.is_contiguous() \
                and actions.shape == buf.shape and actions.data_ptr() % 16 == 0:  # (a misaligned view goes through buf)
            return actions
        buf.copy_(actions.reshape(buf.shape), non_blocking=True)
        return buf
    arr = np.asarray(actions)
    if arr.shape != tuple(buf.shape):
        raise ValueError(f"actions must have shape {tuple(buf.shape)}, got {arr.shape}")
    buf.copy_(torch.from_numpy(np.ascontiguousarray(arr, dtype=np.int64)), non_blocking=False)
    return buf


def host_source(src: "torch.Tensor", staging: "torch.Tensor") -> "torch.Tensor":
    """The pinned host tensor a step_host() call uploads its actions from: the caller's own tensor when it already is
    pinned, contiguous and of the staging buffer's dtype and size (no host-to-host copy inside the call), otherwise the
    env's pinned staging buffer after copying into it.  The upload is asynchronous: with sync=False the caller must keep
    the tensor alive and unchanged until the stream has consumed it."""
    if src.device.type == "cpu" and src.dtype == staging.dtype and src.numel() == staging.numel() \
            and src.is_contiguous() and src.is_pinned():
        return src
    if src.data_ptr() != staging.data_ptr():
        staging.copy_(src.reshape(staging.shape))
    return staging


class LazyInfos(dict):
    """infos dict whose derived entries are computed on access (no per-step kernels or HBM writes for values
    nobody reads).  `lazy[key]` is a zero-argument callable returning the tensor."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.lazy = {}

    def __missing__(self, key):
        if key in self.lazy:
            return self.lazy[key]()
        raise KeyError(key)

    def __contains__(self, key):
        return super().__contains__(key) or key in self.lazy

    def get(self, key, default=None):
        return self[key] if key in self else default

    def keys(self):
        return list(super().keys()) + list(self.lazy)

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return super().__len__() + len(self.lazy)

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def values(self):
        return [self[k] for k in self.keys()]
