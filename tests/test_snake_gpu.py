"""GPU parity tests: the CUDA path (through the C ABI in libbeng.so) against the CPU oracle and the
golden vectors produced by the reference.  Bit-exact: everything here is integer/byte work."""
import zlib

import numpy as np
import pytest
import torch

from conftest import case_meta

pytestmark = pytest.mark.gpu

CASES = ["random_g20", "random_g20_hi", "greedy_g20", "greedy_g8", "greedy_g5", "random_g15", "circle_g20",
         "reverse_g20"]
DEV = "cuda:0"


@pytest.fixture(scope="module")
def pkg():
    import custom_gymnasium_environments_b200 as p

    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    p._lib.load()  # fails loudly when libbeng.so is missing: no fallback
    return p


def crc_rows(obs):
    return np.array([zlib.crc32(o.tobytes()) for o in obs], dtype=np.uint32)


def np_(t):
    return t.cpu().numpy()


def greedy_actions(obs, st, G):
    """Vectorised food-seeking policy (same rule as oracle/gen_golden.greedy_action)."""
    n = obs.shape[0]
    hr, hc, fr, fc, d = (st[k].astype(np.int64) for k in ("head_r", "head_c", "food_r", "food_c", "direction"))
    best = d.copy()
    best_key = np.full(n, np.iinfo(np.int64).max)
    for a, (dr, dc) in enumerate(((-1, 0), (0, 1), (1, 0), (0, -1))):
        nr, nc = hr + dr, hc + dc
        wall = (nr < 0) | (nr >= G) | (nc < 0) | (nc >= G)
        body = np.zeros(n, bool)
        ok = ~wall
        body[ok] = obs[np.nonzero(ok)[0], nr[ok], nc[ok]] == 1
        dead = wall | body
        key = dead * 1_000_000 + (np.abs(nr - fr) + np.abs(nc - fc)) * 10 + a
        key = np.where(np.abs(a - d) == 2, np.iinfo(np.int64).max, key)
        upd = key < best_key
        best, best_key = np.where(upd, a, best), np.where(upd, key, best_key)
    return best


def assert_same(env, orc, t, check_state=True, check_obs=True):
    if check_obs:
        assert np.array_equal(np_(env.obs), orc.obs), f"obs mismatch at step {t}"
    assert np.array_equal(np_(env.reward), orc.reward), f"reward mismatch at step {t}"
    assert np.array_equal(np_(env.terminated).astype(np.uint8), orc.terminated), f"terminated mismatch at {t}"
    assert not np_(env.truncated).any()
    assert np.array_equal(np_(env.score), orc.score) and np.array_equal(np_(env.snake_length), orc.length), t
    if check_state:
        dev, cpu = env.export_state(), orc.state()
        for k in cpu:
            assert np.array_equal(np_(dev[k]).astype(np.int64), cpu[k].astype(np.int64)), (k, t)


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", CASES)
def test_golden_replay(pkg, golden, name):
    """Replay the action tapes recorded from the REFERENCE and compare every recorded field."""
    m = case_meta(golden, name)
    g = {k: golden[f"{name}/{k}"] for k in ("action", "reward", "terminated", "score", "length", "head_r", "head_c",
                                             "food_r", "food_c", "direction", "steps", "rng_counter", "obs_crc",
                                             "snap_obs", "reset_obs", "final_score", "final_steps")}
    env = pkg.BatchedSnakeEnv(m["n_envs"], m["G"], device=DEV, seed=m["seed"], env_id_base=m["base"])
    obs, info = env.reset()
    assert np.array_equal(np_(obs), g["reset_obs"])
    assert np_(info["score"]).tolist() == [0] * m["n_envs"] and np_(info["snake_length"]).tolist() == [1] * m["n_envs"]
    acts = torch.from_numpy(g["action"].astype(np.int64)).to(DEV)
    n_ep = 0
    for t in range(m["n_steps"]):
        obs, rew, term, trunc, info = env.step(acts[:, t].contiguous())
        o = np_(obs)
        assert np.array_equal(np_(rew), g["reward"][:, t]), t
        assert np.array_equal(np_(term).astype(np.uint8), g["terminated"][:, t]), t
        assert not np_(trunc).any()
        assert np.array_equal(np_(info["score"]), g["score"][:, t])
        assert np.array_equal(np_(info["snake_length"]), g["length"][:, t])
        assert np.array_equal(crc_rows(o), g["obs_crc"][:, t]), t
        if t % 7 == 0 or t == m["n_steps"] - 1:
            st = env.export_state()
            for k in ("head_r", "head_c", "food_r", "food_c", "direction", "steps", "rng_counter"):
                assert np.array_equal(np_(st[k]).astype(np.int64), g[k][:, t].astype(np.int64)), (k, t)
        if t % m["snap_every"] == 0:
            assert np.array_equal(o, g["snap_obs"][:, t // m["snap_every"]])
        done = np_(term)
        if done.any():
            died = g["reward"][:, t][done] < 0
            assert np.array_equal(np_(info["episode"]["score"])[done], g["final_score"][:, t][done])
            assert np.array_equal(np_(info["episode"]["l"])[done], g["final_steps"][:, t][done] + died)
            assert np.array_equal(np_(info["episode"]["r"])[done], 10.0 * g["final_score"][:, t][done] - 10.0 * died)
            assert sorted(np_(env.finished_envs()).tolist()) == np.nonzero(done)[0].tolist()
            n_ep += int(done.sum())
        else:
            assert env.finished_envs().numel() == 0
    assert env.episode_stats()["n_episodes"] == n_ep


@pytest.mark.parametrize("mode", ["same_step", "next_step", "disabled"])
@pytest.mark.parametrize("n,G", [(10007, 20), (4096, 15), (333, 8), (130, 33)])
def test_random_rollout_vs_oracle(pkg, mode, n, G):
    """Random device-generated tapes (also checks beng_fill_random_actions against the oracle tape),
    ragged env counts, odd grids, all three auto-reset modes."""
    from oracle import c_oracle

    seed, base, T = 17, 1000, 160
    env = pkg.BatchedSnakeEnv(n, G, device=DEV, seed=seed, env_id_base=base, autoreset_mode=mode)
    orc = c_oracle.SnakeOracle(n, G, seed=seed, env_id_base=base, autoreset=mode)
    obs, _ = env.reset()
    assert np.array_equal(np_(obs), orc.reset())
    lib = pkg._lib.load()
    actions = torch.zeros(n, dtype=torch.int64, device=DEV)
    for t in range(T):
        assert lib.beng_fill_random_actions(actions.data_ptr(), n, 1, 4, t, base, seed,
                                            torch.cuda.current_stream().cuda_stream) == 0
        a = np_(actions)
        assert np.array_equal(a, c_oracle.action_tape(seed, base, n, t, 4))
        env.step(actions)
        orc.step(a)
        assert_same(env, orc, t, check_state=(t % 16 == 0))
    if mode != "disabled":
        assert env.episode_stats() == orc.stats()


@pytest.mark.parametrize("G,n,T", [(20, 2048, 1300), (8, 1024, 700), (4, 512, 300)])
def test_greedy_policy_long_snakes_vs_oracle(pkg, G, n, T):
    """A food-seeking policy grows long snakes: exercises self-collision, the food rejection loop, ring
    wrap-around, the time limit and the body un-draw on auto-reset."""
    from oracle.c_oracle import SnakeOracle

    seed = 5
    env = pkg.BatchedSnakeEnv(n, G, device=DEV, seed=seed)
    orc = SnakeOracle(n, G, seed=seed)
    env.reset(), orc.reset()
    max_len = 0
    for t in range(T):
        a = greedy_actions(orc.obs, orc.state(), G)
        if t % 50 == 49:  # sprinkle suicidal / invalid-free noise
            a[::7] = (a[::7] + 1) % 4
        env.step(torch.from_numpy(a).to(DEV))
        orc.step(a)
        assert_same(env, orc, t, check_state=(t % 25 == 0))
        max_len = max(max_len, int(orc.length.max()))
    assert max_len >= min(30, G * G - 2)
    body = np_(env.export_state(with_body=True)["body"])
    for e in range(0, n, 97):
        ref = orc.body(e)
        assert np.array_equal(body[e, : len(ref)], ref) and (body[e, len(ref):] == -1).all()
    assert env.episode_stats() == orc.stats()


def test_full_size_batch_vs_oracle_and_invariants(pkg):
    """BASELINE config: 1,048,576 envs.  50 steps against the C oracle bit for bit, then size-independent
    invariants of the observation encoding."""
    from oracle.c_oracle import SnakeOracle

    n, G, seed, T = 1 << 20, 20, 0, 50
    env = pkg.BatchedSnakeEnv(n, G, device=DEV, seed=seed)
    orc = SnakeOracle(n, G, seed=seed)
    assert np.array_equal(np_(env.reset()[0]), orc.reset())
    lib = pkg._lib.load()
    actions = torch.zeros(n, dtype=torch.int64, device=DEV)
    total_done = 0
    for t in range(T):
        lib.beng_fill_random_actions(actions.data_ptr(), n, 1, 4, t, 0, seed, torch.cuda.current_stream().cuda_stream)
        obs, rew, term, trunc, info = env.step(actions)
        orc.step(np_(actions), want_obs=(t % 10 == 9))
        assert_same(env, orc, t, check_state=(t == T - 1), check_obs=(t % 10 == 9))
        # invariants: exactly `length` ones and one food cell per env; rewards in {-10, 0, 10}
        ones = (obs == 1).sum(dim=(1, 2))
        twos = (obs == 2).sum(dim=(1, 2))
        assert torch.equal(ones.int(), info["snake_length"]) and bool((twos == 1).all())
        assert bool(((rew == 0) | (rew == 10) | (rew == -10)).all())
        total_done += int(term.sum())
        assert env.finished_envs().numel() == int(term.sum())
    st = env.episode_stats()
    assert st["n_episodes"] == total_done and st == orc.stats()


def test_trajectories_do_not_depend_on_sharding(pkg):
    """GPU-count invariance: env i behaves the same whether it lives in one batch of N or
    in one of several shards addressed by env_id_base."""
    n, T, seed = 6000, 120, 9
    whole = pkg.BatchedSnakeEnv(n, device=DEV, seed=seed)
    from custom_gymnasium_environments_b200.dist import shard_range

    shards = []
    for r in range(3):
        s, c = shard_range(n, r, 3)
        shards.append((s, c, pkg.BatchedSnakeEnv(c, device=DEV, seed=seed, env_id_base=s)))
    whole.reset()
    for _, _, e in shards:
        e.reset()
    lib = pkg._lib.load()
    actions = torch.zeros(n, dtype=torch.int64, device=DEV)
    for t in range(T):
        lib.beng_fill_random_actions(actions.data_ptr(), n, 1, 4, t, 0, seed, torch.cuda.current_stream().cuda_stream)
        whole.step(actions)
        for s, c, e in shards:
            e.step(actions[s:s + c].contiguous())
            assert torch.equal(e.obs, whole.obs[s:s + c]) and torch.equal(e.reward, whole.reward[s:s + c])
            assert torch.equal(e.terminated, whole.terminated[s:s + c])
    tot = sum(e.stats[:4] for _, _, e in shards)
    assert torch.equal(tot, whole.stats[:4])


def test_host_buffer_path_matches_device_path(pkg):
    n, seed, T = 5000, 3, 80
    a = pkg.BatchedSnakeEnv(n, device=DEV, seed=seed)
    b = pkg.BatchedSnakeEnv(n, device=DEV, seed=seed)
    a.reset(), b.reset()
    rng = np.random.default_rng(0)
    for t in range(T):
        act = rng.integers(0, 4, n)
        obs, rew, term, trunc, info = a.step_host(act)
        b.step(torch.from_numpy(act).to(DEV))
        assert isinstance(obs, np.ndarray) and obs.dtype == np.int8 and rew.dtype == np.float32
        assert np.array_equal(obs, np_(b.obs)) and np.array_equal(rew, np_(b.reward))
        assert np.array_equal(term, np_(b.terminated)) and not trunc.any()
        assert np.array_equal(info["score"], np_(b.score))
    obs2, *_ = a.step_host(rng.integers(0, 4, n), copy_obs=False)
    assert isinstance(obs2, torch.Tensor) and obs2.is_cuda


def test_invalid_actions(pkg):
    env = pkg.BatchedSnakeEnv(64, device=DEV, seed=1, debug_checks=True)
    env.reset()
    before = {k: v.clone() for k, v in env.export_state().items()}
    bad = torch.full((64,), 9, dtype=torch.int64, device=DEV)
    with pytest.raises(ValueError, match="Invalid action"):
        env.step(bad)
    after = env.export_state()
    for k in before:
        assert torch.equal(before[k], after[k]), k
    assert not env.terminated.any() and (env.reward == 0).all()
    env.step(torch.ones(64, dtype=torch.int64, device=DEV))  # recovers
    with pytest.raises(ValueError):
        env.step(np.zeros(3, dtype=np.int64))


def test_masked_reset_and_seed_rekey(pkg):
    n = 512
    env = pkg.BatchedSnakeEnv(n, device=DEV, seed=4)
    env.reset()
    for _ in range(5):
        env.step(torch.ones(n, dtype=torch.int64, device=DEV))
    st0 = env.export_state()
    mask = torch.zeros(n, dtype=torch.bool, device=DEV)
    mask[::2] = True
    obs, _ = env.reset(options={"reset_mask": mask})
    st1 = env.export_state()
    assert (st1["steps"][::2] == 0).all() and (st1["steps"][1::2] == 5).all()
    assert torch.equal(st1["head_c"][1::2], st0["head_c"][1::2])
    assert (st1["rng_counter"][::2] > st0["rng_counter"][::2]).all()       # masked reset keeps drawing forward
    assert bool(((obs == 1).sum(dim=(1, 2)) == st1["length"]).all()) and (st1["length"][::2] == 1).all()
    # re-keying with the same seed reproduces the constructor stream
    env.reset(seed=4)
    fresh = pkg.BatchedSnakeEnv(n, device=DEV, seed=4)
    fresh.reset()
    assert torch.equal(env.obs, fresh.obs)
    env.reset(seed=5)
    assert not torch.equal(env.obs, fresh.obs)


def test_rollout_buffer_output_and_state_dict(pkg):
    n, T = 3000, 20
    env = pkg.BatchedSnakeEnv(n, device=DEV, seed=8)
    twin = pkg.BatchedSnakeEnv(n, device=DEV, seed=8)
    env.reset(), twin.reset()
    rollout = torch.zeros((T, n, 20, 20), dtype=torch.int8, device=DEV)
    acts = torch.randint(0, 4, (T, n), device=DEV)
    for t in range(T):
        obs, *_ = env.step(acts[t], out_obs=rollout[t])
        assert obs.data_ptr() == rollout[t].data_ptr()
        twin.step(acts[t])
        assert torch.equal(rollout[t], twin.obs)
    sd = env.state_dict()
    clone = pkg.BatchedSnakeEnv(n, device=DEV, seed=999)
    clone.load_state_dict(sd)
    a = torch.randint(0, 4, (n,), device=DEV)
    env.step(a), clone.step(a)
    assert torch.equal(env.obs, clone.obs) and torch.equal(env.reward, clone.reward)


def test_single_env_facade_matches_reference_anchor(pkg, golden):
    """SnakeEnvClassic (gym.Env surface) against the frozen-state rows recorded from the reference."""
    rows = golden["frozen/rows"]
    env = pkg.SnakeEnvClassic(device=DEV, seed=0, env_id=0)
    obs, info = env.reset()
    assert obs.shape == (20, 20) and obs.dtype == np.int8 and info == {"score": 0, "snake_length": 1}
    assert env.action_space.n == 4 and env.observation_space.shape == (20, 20)
    for t, (r, term, steps, hr, hc, d, score, length, crc) in enumerate(rows):
        obs, rew, te, tr, info = env.step(1 if t < 12 else 0)
        assert (rew, te, tr) == (r, bool(term), False) and zlib.crc32(obs.tobytes()) == int(crc)
        assert info["score"] == score and info.get("snake_length", -1) == length
        assert env.steps == steps
        if r == 0:
            assert isinstance(rew, int)
    with pytest.raises(ValueError):
        env.step(4)
    with pytest.raises(ValueError):
        env.step(1.0)


def test_render_rgb_array_and_misaligned_action_views(pkg):
    """render(): the reference's 3-colour LUT (snake_env.py:175-188) over the current observation.  A contiguous action
    view that starts at an odd element (8-byte but not 16-byte aligned storage offset) must take the staging-buffer path
    and give the same step as an aligned copy; through the C ABI a misaligned pointer is an argument error, not a fault."""
    import ctypes as C

    n = 64
    env = pkg.BatchedSnakeEnv(n, device=DEV, seed=2)
    env.render_mode = "rgb_array"
    obs, _ = env.reset()
    img = env.render()
    assert img.shape == (n, 400, 400, 3) and img.dtype == np.uint8
    cell = img[:, ::20, ::20]  # one pixel per grid cell
    o = obs.cpu().numpy()
    assert np.array_equal(cell[..., 1] == 255, o == 1) and np.array_equal(cell[..., 0] == 255, o == 2)
    twin = pkg.BatchedSnakeEnv(n, device=DEV, seed=2)
    twin.reset()
    big = torch.randint(0, 4, (n + 1,), device=DEV)
    view = big[1:]  # data_ptr() % 16 == 8
    assert view.is_contiguous() and view.data_ptr() % 16 != 0
    a_obs, a_rew, *_ = env.step(view)
    b_obs, b_rew, *_ = twin.step(view.clone())
    assert torch.equal(a_obs, b_obs) and torch.equal(a_rew, b_rew)
    lib = pkg._lib.load()
    rc = lib.beng_snake_step(C.byref(env.params), C.byref(env._state), view.data_ptr() + 4, C.byref(env._ios_full[0]), n,
                             torch.cuda.current_stream().cuda_stream)
    assert rc == -1  # BENG_ERR_BAD_ARG
