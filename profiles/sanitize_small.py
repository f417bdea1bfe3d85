"""Small invocation of every kernel for compute-sanitizer (memcheck / racecheck): ragged env counts, short
episodes so that auto-reset paths run, all three auto-reset modes.  Run under gpurun:
    compute-sanitizer --tool memcheck python profiles/sanitize_small.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import custom_gymnasium_environments_b200 as pkg  # noqa: E402

dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
for mode in ("same_step", "next_step", "disabled"):
    for n, G in ((301, 20), (67, 15), (130, 8)):
        env = pkg.BatchedSnakeEnv(n, G, device=dev, seed=1, autoreset_mode=mode, max_steps=12)
        env.reset()
        for t in range(30):
            env.step(torch.randint(0, 4, (n,), device=dev, generator=g))
        env.step_host(torch.randint(0, 4, (n,), generator=torch.Generator().manual_seed(t)).numpy())
    for n, kind in ((101, "discrete"), (70, "continuous")):
        env = pkg.BatchedCryptoTradingEnv(n, None, kind, device=dev, seed=2, autoreset_mode=mode, max_steps=7)
        env.reset()
        for t in range(20):
            a = torch.randint(0, 5, (n,), device=dev, generator=g) if kind == "discrete" else \
                torch.rand((n, 2), device=dev, generator=g) * 2 - 1
            env.step(a)
    for n, kw in ((133, {}), (65, dict(grid_size=(3, 4), num_intersections=7, max_vehicles=20, spawn_rate=0.9))):
        env = pkg.BatchedTrafficManagementEnv(n, device=dev, seed=3, autoreset_mode=mode, max_timesteps=9, **kw)
        env.reset()
        for t in range(25):
            env.step(torch.randint(0, 3, (n, env.num_intersections), device=dev, generator=g))
torch.cuda.synchronize()
print("sanitize_small: done, launches =", pkg._lib.load().beng_launch_count())
