"""Smallest program that launches the step kernel of one env at its BASELINE size a few times (the ncu target).

    python profiles/prof_step.py crypto [steps] [n_envs]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import custom_gymnasium_environments_b200 as pkg  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "crypto"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
dev = torch.device("cuda:0")
sizes = {"snake": 1 << 20, "crypto": 1 << 18, "traffic": 1 << 16, "climate": 1 << 20, "builder": 1 << 20}
n = int(sys.argv[3]) if len(sys.argv) > 3 else sizes[name]
g = torch.Generator(device=dev).manual_seed(0)
if name == "crypto":
    env, acts = pkg.BatchedCryptoTradingEnv(n, None, "discrete", device=dev), torch.randint(0, 5, (8, n), device=dev, generator=g)
elif name == "snake":
    env, acts = pkg.BatchedSnakeEnv(n, device=dev), torch.randint(0, 4, (8, n), device=dev, generator=g)
elif name == "traffic":
    env, acts = pkg.BatchedTrafficManagementEnv(n, device=dev), torch.randint(0, 3, (8, n, 9), device=dev, generator=g)
elif name == "builder":
    env, acts = pkg.BatchedWorldBuilderEnv(n, device=dev), torch.randint(0, 5, (8, n), device=dev, generator=g)
else:
    env = pkg.BatchedSmartClimateEnv(n, device=dev)
    acts = [{"ac_temp": torch.rand(n, device=dev, generator=g) * 16 + 16,
             "lights": torch.randint(0, 2, (n, 4), device=dev, generator=g).to(torch.int8)} for _ in range(8)]
env.reset()
for t in range(steps):
    env.step(acts[t % 8])
torch.cuda.synchronize()
print("ok", name, n, steps)
