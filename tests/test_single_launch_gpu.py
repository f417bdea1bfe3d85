"""Every Batched*Env.step() with device-resident, aligned actions is exactly ONE kernel launch: no host-side tensor op
(mask, cast, copy) rides along inside the step, so the device-timed figures of bench.py are the step kernel's."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def pkg():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import custom_gymnasium_environments_b200 as p
    return p


def make(pkg, name, n):
    g = torch.Generator(device=DEV).manual_seed(3)
    if name == "snake":
        return pkg.BatchedSnakeEnv(n, device=DEV), torch.randint(0, 4, (n,), device=DEV, generator=g)
    if name == "crypto":
        return pkg.BatchedCryptoTradingEnv(n, None, "discrete", device=DEV), torch.randint(0, 5, (n,), device=DEV, generator=g)
    if name == "traffic":
        return pkg.BatchedTrafficManagementEnv(n, device=DEV), torch.randint(0, 3, (n, 9), device=DEV, generator=g)
    if name == "builder":
        return pkg.BatchedWorldBuilderEnv(n, device=DEV), torch.randint(0, 5, (n,), device=DEV, generator=g)
    return pkg.BatchedSmartClimateEnv(n, device=DEV), {
        "ac_temp": torch.rand(n, device=DEV, generator=g) * 16 + 16,
        "lights": torch.randint(0, 2, (n, 4), device=DEV, generator=g).to(torch.int8)}


@pytest.mark.parametrize("name", ["snake", "crypto", "traffic", "climate", "builder"])
def test_step_is_one_kernel(pkg, name):
    env, act = make(pkg, name, 8192)
    env.reset()
    lib = pkg._lib.load()
    for _ in range(3):
        env.step(act)
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        before = lib.beng_launch_count()
        for _ in range(5):
            env.step(act)
        torch.cuda.synchronize()
    assert lib.beng_launch_count() - before == 5
    kernels = [e.name for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    assert len(kernels) == 5 and len(set(kernels)) == 1, kernels


@pytest.mark.parametrize("name", ["snake", "crypto", "traffic", "climate", "builder"])
def test_step_host_from_pinned_tensors_matches_numpy(pkg, name):
    """step_host() uploads straight from a caller's pinned tensor (no staging copy); same results as the numpy path."""
    import numpy as np

    n = 2048 + 5
    env_a, act = make(pkg, name, n)
    env_b, _ = make(pkg, name, n)
    env_a.reset(); env_b.reset()
    for t in range(6):
        if isinstance(act, dict):
            dev = {k: (v + t) % 2 if v.dtype == torch.int8 else v for k, v in act.items()}
            host_np = {k: v.cpu().numpy() for k, v in dev.items()}
            host_pin = {k: v.cpu().pin_memory() for k, v in dev.items()}
        else:
            dev = (act + t) % (int(act.max()) + 1)
            host_np, host_pin = dev.cpu().numpy(), dev.cpu().pin_memory()
        out_a = env_a.step_host(host_np)
        out_b = env_b.step_host(host_pin)
        used = env_b._host_src
        pins = list(host_pin.values()) if isinstance(host_pin, dict) else [host_pin]
        used = list(used) if isinstance(used, tuple) else [used]
        assert {u.data_ptr() for u in used} == {p.data_ptr() for p in pins}  # no staging copy
        for x, y in zip(out_a[:4], out_b[:4]):
            if isinstance(x, dict):
                for k in x:
                    assert np.array_equal(np.asarray(x[k]), np.asarray(y[k])), (name, t, k)
            else:
                assert np.array_equal(np.asarray(x), np.asarray(y)), (name, t)


@pytest.mark.parametrize("name", ["traffic", "climate", "builder"])
def test_steps_replay_from_a_cuda_graph(pkg, name):
    """step() is a pure stream operation for the envs without host-side per-step state (no allocation, no sync, no host
    counter): a rollout's inner loop can be captured once in a CUDA graph and replayed, which is what a launch-bound
    batch (traffic at 65,536 envs: ~12 us of kernel per ~13 us of Python enqueue) wants.  Replays must walk the same
    trajectory as eager stepping."""
    n = 4096 + 3
    env_a, act = make(pkg, name, n)
    env_b, _ = make(pkg, name, n)
    env_a.reset(); env_b.reset()
    if isinstance(act, dict):
        tape = [{"ac_temp": act["ac_temp"] + k, "lights": (act["lights"] + k) % 2} for k in range(3)]
    else:
        hi = int(act.max()) + 1
        tape = [((act + k) % hi).contiguous() for k in range(3)]
    for a in tape[:2]:  # eager warm-up on both (lazy initialisation happens outside the capture)
        env_a.step(a); env_b.step(a)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for a in tape:
            env_b.step(a)
    for _ in range(2):
        g.replay()
        for a in tape:
            env_a.step(a)
    torch.cuda.synchronize()
    sa, sb = env_a.state_dict(), env_b.state_dict()
    for k in sa:
        if isinstance(sa[k], torch.Tensor):
            assert torch.equal(sa[k], sb[k]), (name, k)
    obs_a = env_a.grid if name == "builder" else env_a.obs
    obs_b = env_b.grid if name == "builder" else env_b.obs
    assert torch.equal(obs_a, obs_b) and torch.equal(env_a.reward, env_b.reward)
