"""The opt-in TimeLimit-equivalent truncation flag: what gym.make's TimeLimit wrapper
(max_episode_steps=1000 in the reference's register() calls) adds on top of the raw classes.  Default stays the raw
class: truncated is always False."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def pkg():
    import custom_gymnasium_environments_b200 as p

    assert torch.cuda.is_available()
    return p


def test_snake_truncation_flag(pkg):
    n, limit = 4096, 9
    on = pkg.BatchedSnakeEnv(n, device=DEV, seed=2, max_steps=limit, time_limit_truncation=True)
    off = pkg.BatchedSnakeEnv(n, device=DEV, seed=2, max_steps=limit)
    on.reset(), off.reset()
    seen = 0
    g = torch.Generator(device=DEV).manual_seed(0)
    for t in range(60):
        a = torch.randint(0, 4, (n,), device=DEV, generator=g)
        _, r1, term1, trunc1, info = on.step(a)
        _, r0, term0, trunc0, _ = off.step(a)
        assert torch.equal(term1, term0) and torch.equal(r1, r0) and not bool(trunc0.any())
        hit_limit = term1 & (r1 != -10.0) & (info["episode"]["l"] == limit)   # ended by the limit, not by a death
        assert torch.equal(trunc1, hit_limit)
        seen += int(trunc1.sum())
    assert seen > 100


def test_crypto_truncation_flag(pkg):
    n, limit = 2048, 6
    on = pkg.BatchedCryptoTradingEnv(n, None, "discrete", device=DEV, seed=2, max_steps=limit, time_limit_truncation=True)
    off = pkg.BatchedCryptoTradingEnv(n, None, "discrete", device=DEV, seed=2, max_steps=limit)
    on.reset(), off.reset()
    g = torch.Generator(device=DEV).manual_seed(0)
    for t in range(20):
        a = torch.randint(0, 5, (n,), device=DEV, generator=g)
        _, _, term1, trunc1, _ = on.step(a)
        _, _, term0, trunc0, _ = off.step(a)
        assert torch.equal(term1, term0) and not bool(trunc0.any())
        assert bool(trunc1.all()) == ((t + 1) % limit == 0) and (bool(trunc1.any()) == bool(trunc1.all()))
        assert torch.equal(on.obs, off.obs)


def test_traffic_truncation_flag(pkg):
    n, limit = 1024, 7
    on = pkg.BatchedTrafficManagementEnv(n, device=DEV, seed=2, max_timesteps=limit, time_limit_truncation=True)
    off = pkg.BatchedTrafficManagementEnv(n, device=DEV, seed=2, max_timesteps=limit)
    on.reset(), off.reset()
    g = torch.Generator(device=DEV).manual_seed(0)
    for t in range(22):
        a = torch.randint(0, 3, (n, 9), device=DEV, generator=g)
        _, _, term1, trunc1, _ = on.step(a)
        _, _, term0, trunc0, _ = off.step(a)
        assert torch.equal(term1, term0) and torch.equal(trunc1, term1) and not bool(trunc0.any())
        assert bool(term1.all()) == ((t + 1) % limit == 0)
        assert torch.equal(on.obs, off.obs)
