"""Philox4x32-10 and the engine's draw contract, restated in numpy.

TEST INFRASTRUCTURE (oracle side).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package.

The reference draws from Python's global `random` (Mersenne Twister;
snake_env_classic/snake_env.py:4,125-126).  A batched device engine cannot share one
sequential generator across a million envs, so the engine defines its OWN counter-based
stream per env and parity is established by replaying that stream into the reference
(oracle/replay.py rebinding `snake_env.random`), as BASELINE.json's north_star prescribes.

Contract (identical in csrc/beng_rng.cuh, oracle/c/beng_oracle_rng.h and here):

  u32 draw number j of env e on stream s under seed S  =
      philox4x32_10(counter = (j >> 2, e & 0xffffffff, e >> 32, s),
                    key     = (S & 0xffffffff, S >> 32))[j & 3]

  randint(a, b)   = a + ((u32 * (b - a + 1)) >> 32)            (one draw)
  random()        = (u32a >> 5) * 2**26 + (u32b >> 6)) / 2**53   (two draws, 53-bit like CPython)
  uniform(a, b)   = a + (b - a) * random()                       (two draws; same expression as CPython)
  choice(seq)     = seq[randint(0, len(seq) - 1)]                (one draw)
  normal(mu, sd)  = mu + sd * sqrt(-2 ln(1 - u1)) * cos(2 pi u2), u1, u2 = random(), random()  (four draws)

Streams: 0 = env dynamics, 1 = synthetic action tape (bench / tests).

Known-answer vectors are the Random123 distribution's kat_vectors for philox4x32 10
(checked in tests/test_rng.py).
"""
from __future__ import annotations

import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)

STREAM_ENV = 0
STREAM_ACTION = 1


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All arguments broadcastable integer arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK32 for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def draws_u32(seed: int, env_id, first: int, count: int, stream: int = STREAM_ENV) -> np.ndarray:
    """u32 draws [first, first+count) of each env in `env_id` -> array (len(env_id), count)."""
    env_id = np.atleast_1d(np.asarray(env_id, dtype=np.uint64))
    j = np.arange(first, first + count, dtype=np.uint64)
    blk = (j >> np.uint64(2))[None, :]
    e_lo = (env_id & MASK32)[:, None]
    e_hi = (env_id >> np.uint64(32))[:, None]
    r = philox4x32_10(blk, e_lo, e_hi, np.uint64(stream), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    lanes = np.stack(r, axis=-1)  # (E, count, 4)
    sel = (j & np.uint64(3)).astype(np.int64)
    out = np.take_along_axis(lanes, np.broadcast_to(sel[None, :, None], lanes.shape[:2] + (1,)), axis=-1)
    return out[..., 0]


def randint_from_u32(u, a: int, b: int):
    """randint(a, b) inclusive from one u32 draw (multiply-high range map)."""
    span = np.uint64(b - a + 1)
    return (a + ((np.asarray(u, dtype=np.uint64) * span) >> np.uint64(32))).astype(np.int64)


def random53_from_u32(ua, ub):
    """random() in [0,1) with 53 bits from two u32 draws (same construction CPython uses)."""
    a = np.asarray(ua, dtype=np.uint64) >> np.uint64(5)
    b = np.asarray(ub, dtype=np.uint64) >> np.uint64(6)
    return (a.astype(np.float64) * 67108864.0 + b.astype(np.float64)) * (1.0 / 9007199254740992.0)


def action_tape(seed: int, env_id, step_first: int, n_steps: int, n_choices: int, n_cols: int = 1) -> np.ndarray:
    """Synthetic uniform action tape, identical to csrc `beng_fill_random_actions`.

    action[e, t, c] = randint(0, n_choices-1) from draw (t * n_cols + c) of stream 1.
    Returns int64 array (len(env_id), n_steps, n_cols) (n_cols axis dropped when 1).
    """
    u = draws_u32(seed, env_id, step_first * n_cols, n_steps * n_cols, stream=STREAM_ACTION)
    a = randint_from_u32(u, 0, n_choices - 1).reshape(len(np.atleast_1d(env_id)), n_steps, n_cols)
    return a[..., 0] if n_cols == 1 else a
