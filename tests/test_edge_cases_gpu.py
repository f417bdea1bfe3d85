"""GPU edge cases through the C ABI: tiny and ragged env counts (1, 31, 33: less than a warp, a warp plus one), the
largest traffic grid the engine accepts (25 intersections = 26-warp CTAs, 255 vehicles, every step spawns), empty
batches, and argument rejection.  Same bars as the per-env parity files: integer envs EXACT, float envs rtol 1e-5."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL, ATOL = 1e-5, 1e-6  # float32 observations of the float64 envs (crypto, climate), as in their own test files


@pytest.fixture(scope="module")
def pkg():
    import custom_gymnasium_environments_b200 as p

    assert torch.cuda.is_available()
    p._lib.load()
    return p


def np_(t):
    return t.cpu().numpy()


def fill(pkg, actions, n_choices, t, base, seed):
    n = actions.shape[0]
    cols = actions.numel() // max(n, 1)
    pkg._lib.load().beng_fill_random_actions(actions.data_ptr(), n, cols, n_choices, t, base, seed,
                                             torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("n", [1, 31, 33])
def test_tiny_batches_all_envs(pkg, n):
    from oracle.c_oracle import BuilderOracle, ClimateOracle, CryptoOracle, SnakeOracle, TrafficOracle

    seed, base = 3, 12345
    # snake: short episodes (max_steps 40) so the auto-reset path runs
    env = pkg.BatchedSnakeEnv(n, max_steps=40, device=DEV, seed=seed, env_id_base=base)
    orc = SnakeOracle(n, max_steps=40, seed=seed, env_id_base=base)
    assert np.array_equal(np_(env.reset()[0]), orc.reset())
    a = torch.zeros(n, dtype=torch.int64, device=DEV)
    for t in range(100):
        fill(pkg, a, 4, t, base, seed)
        obs, rew, term, trunc, _ = env.step(a)
        orc.step(np_(a))
        assert np.array_equal(np_(obs), orc.obs) and np.array_equal(np_(rew), orc.reward), t
        assert np.array_equal(np_(term).astype(np.uint8), orc.terminated), t

    # traffic
    env = pkg.BatchedTrafficManagementEnv(n, device=DEV, seed=seed, env_id_base=base, max_timesteps=30)
    orc = TrafficOracle(n, seed=seed, env_id_base=base, max_timesteps=30)
    assert np.array_equal(np_(env.reset()[0]), orc.reset())
    a = torch.zeros((n, 9), dtype=torch.int64, device=DEV)
    for t in range(70):
        fill(pkg, a, 3, t, base, seed)
        env.step(a)
        orc.step(np_(a))
        assert np.array_equal(np_(env.obs), orc.obs) and np.array_equal(np_(env.reward64), orc.reward64), t
        assert np.array_equal(np_(env.terminated).astype(np.uint8), orc.terminated), t

    # builder
    env = pkg.BatchedWorldBuilderEnv(n, device=DEV, seed=seed, env_id_base=base)
    orc = BuilderOracle(n, seed=seed, env_id_base=base)
    env.reset(), orc.reset()
    a = torch.zeros(n, dtype=torch.int64, device=DEV)
    for t in range(60):
        fill(pkg, a, 5, t, base, seed)
        _, rew, term, _, _ = env.step(a)
        orc.step(np_(a))
        assert np.array_equal(np_(env.grid), orc.grid) and np.array_equal(np_(rew), orc.reward), t
        assert np.array_equal(np_(env.resources), orc.resources), t

    # crypto (float64 dynamics, float32 observation)
    env = pkg.BatchedCryptoTradingEnv(n, None, "discrete", device=DEV, seed=seed, env_id_base=base)
    orc = CryptoOracle(n, seed=seed, env_id_base=base)
    np.testing.assert_allclose(np_(env.reset()[0]), orc.reset(), rtol=RTOL, atol=ATOL)
    a = torch.zeros(n, dtype=torch.int64, device=DEV)
    for t in range(40):
        fill(pkg, a, 5, t, base, seed)
        env.step(a)
        orc.step(np_(a))
        np.testing.assert_allclose(np_(env.obs), orc.obs, rtol=RTOL, atol=ATOL, err_msg=f"crypto obs, step {t}")
        assert np.array_equal(np_(env.terminated).astype(np.uint8), orc.terminated), t

    # climate
    env = pkg.BatchedSmartClimateEnv(n, episode_minutes=25, device=DEV, seed=seed, env_id_base=base)
    orc = ClimateOracle(n, episode_minutes=25, seed=seed, env_id_base=base)
    np.testing.assert_allclose(np_(env.reset()[0]), orc.reset(), rtol=RTOL, atol=ATOL)
    gen = torch.Generator(device=DEV).manual_seed(n)
    for t in range(60):
        ac = torch.rand(n, device=DEV, generator=gen) * 24 + 12
        li = torch.randint(0, 2, (n, 4), device=DEV, generator=gen).to(torch.int8)
        env.step({"ac_temp": ac, "lights": li})
        orc.step(np_(ac), np_(li))
        np.testing.assert_allclose(np_(env.obs), orc.obs, rtol=RTOL, atol=ATOL, err_msg=f"climate obs, step {t}")
        assert np.array_equal(np_(env.terminated).astype(np.uint8), orc.terminated), t


@pytest.mark.parametrize("mode", ["same_step", "next_step"])
def test_traffic_largest_grid(pkg, mode):
    """5x5 grid with all 25 intersections (BENG_TRAFFIC_MAX_INTERSECTIONS), 255 vehicles, a spawn every step: the
    26-warp CTA of the generic-NI kernel, the >= 24-element branch of the pairwise np.var sum, and light draws that
    run past the env warp's pre-computed Philox blocks."""
    from oracle.c_oracle import TrafficOracle

    n, seed, base, T = 1027, 17, 5, 260
    kw = dict(grid_size=(5, 5), num_intersections=25, max_vehicles=255, spawn_rate=1.0)
    env = pkg.BatchedTrafficManagementEnv(n, device=DEV, seed=seed, env_id_base=base, autoreset_mode=mode,
                                          max_timesteps=120, **kw)
    orc = TrafficOracle(n, seed=seed, env_id_base=base, autoreset=mode, max_timesteps=120, **kw)
    assert env.num_intersections == 25 and env.single_observation_space.shape == (25 * 14 + 4,)
    assert np.array_equal(np_(env.reset()[0]), orc.reset())
    a = torch.zeros((n, 25), dtype=torch.int64, device=DEV)
    for t in range(T):
        fill(pkg, a, 3, t, base, seed)
        env.step(a)
        orc.step(np_(a))
        assert np.array_equal(np_(env.reward64), orc.reward64), t
        assert np.array_equal(np_(env.terminated).astype(np.uint8), orc.terminated), t
        if t % 10 == 0 or 115 <= t <= 125:
            assert np.array_equal(np_(env.obs), orc.obs), t
            st = orc.state()
            assert np.array_equal(np_(env.rng_counter), st["rng_counter"].astype(np.int64)), t
            assert np.array_equal(np_(env.num_vehicles), st["num_vehicles"]), t
    assert env.episode_stats()["n_episodes"] == orc.stats()["n_episodes"] == 2 * n


def test_empty_batch_and_argument_rejection(pkg):
    lib, L = pkg._lib.load(), pkg._lib
    stream = torch.cuda.current_stream().cuda_stream
    launches = lib.beng_launch_count()

    # a valid snake call with n_envs = 0 is a no-op that launches nothing
    env = pkg.BatchedSnakeEnv(64, device=DEV)
    env.reset()
    p, st, io, act = C.byref(env.params), C.byref(env._state), C.byref(env._ios_full[0]), env._actions.data_ptr()
    assert lib.beng_snake_step(p, st, act, io, 0, stream) == 0
    assert lib.beng_snake_reset(p, st, io, None, 0, 1, stream) == 0
    assert lib.beng_launch_count() == launches + 1  # only the reset above

    # NULL / negative arguments are rejected without touching the device
    assert lib.beng_snake_step(None, st, act, io, 64, stream) != 0
    assert lib.beng_snake_step(p, st, None, io, 64, stream) != 0
    assert lib.beng_snake_step(p, st, act, io, -1, stream) != 0

    # parameters outside what the kernels were built for are refused with an error code, not clamped
    tenv = pkg.BatchedTrafficManagementEnv(32, device=DEV)
    tenv.reset()
    for field, value in (("num_intersections", 26), ("max_vehicles", 256), ("max_timesteps", 70000),
                         ("autoreset_mode", 7)):
        p = L.TrafficParams.from_buffer_copy(tenv.params)
        setattr(p, field, value)
        if field == "num_intersections":
            p.grid_rows, p.grid_cols = 6, 6
        rc = lib.beng_traffic_step(C.byref(p), C.byref(tenv._state), tenv._actions.data_ptr(), C.byref(tenv._io), 32,
                                   stream)
        assert rc != 0, field
    torch.cuda.synchronize()
    with pytest.raises(RuntimeError, match="BENG_ERR_UNSUPPORTED"):  # the host class surfaces the code, loudly
        pkg.BatchedTrafficManagementEnv(8, grid_size=(6, 6), num_intersections=30, device=DEV).reset()
