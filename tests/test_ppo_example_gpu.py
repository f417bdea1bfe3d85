"""SURVEY.md 8(f) rank 1 / BASELINE.json configs[4]: the on-device PPO rollout consumer (examples/ppo_snake.py).

The env writes every step's observation straight into the (T + 1, N, 20, 20) rollout buffer (`step(..., out_obs=)`); the
test replays the recorded action tape through a SECOND env that uses its own observation buffer and checks that the
rollout buffer holds exactly that trajectory (so the zero-copy path does not perturb anything), that two runs with the
same seed agree, that the loss is finite and that the result carries the env-step / policy-forward / update time split.
"""
import importlib.util
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_example():
    spec = importlib.util.spec_from_file_location("ppo_snake_example", os.path.join(ROOT, "examples", "ppo_snake.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_ppo_rollout_buffer_is_the_env_trajectory():
    import custom_gymnasium_environments_b200 as pkg

    ex = load_example()
    argv = ["--envs", "4096", "--horizon", "8", "--iters", "2", "--minibatch", "8192", "--fwd-chunk", "4096"]
    runs = []
    for _ in range(2):
        rec = []
        out = ex.run(ex.parse_args(argv), record=rec)
        runs.append((out, rec))
    out, rec = runs[0]

    # the three-way split of BASELINE.json configs[4]
    assert set(out["per_iteration_ms"]) == {"env_step_ms", "policy_forward_ms", "update_ms"}
    assert all(v > 0 for v in out["per_iteration_ms"].values())
    assert abs(sum(out["share"].values()) - 1.0) < 1e-9
    assert math.isfinite(out["loss"]) and out["n_gpus"] == 1 and out["envs_per_gpu"] == 4096
    assert out["episodes"]["episodes"] > 0 and out["env_steps_per_s_env_only"] > 0

    # replay: a second env, same seed, its OWN observation buffer, the recorded actions
    env = pkg.BatchedSnakeEnv(4096, 20, device="cuda:0", seed=0, env_id_base=0)
    obs, _ = env.reset()
    assert torch.equal(obs, rec[0][1][0]), "reset observation"
    for it, (acts, obs_buf) in enumerate(rec):
        for t in range(acts.shape[0]):
            obs, *_ = env.step(acts[t])
            assert torch.equal(obs, obs_buf[t + 1]), f"iteration {it}, step {t}: rollout buffer != env trajectory"
    assert env.episode_stats()["n_episodes"] == out["episodes"]["episodes"]

    # the environment side is deterministic under a fixed seed and action tape: run 2's first iteration starts from
    # the same reset observation, and whenever its sampled actions coincide so do the observations
    out2, rec2 = runs[1]
    assert torch.equal(rec2[0][1][0], rec[0][1][0])
    assert math.isfinite(out2["loss"])
    if torch.equal(rec2[0][0], rec[0][0]):
        assert torch.equal(rec2[0][1], rec[0][1])
