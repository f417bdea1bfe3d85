"""Same-box A/B of step-kernel variants selected by an environment variable the library reads at every launch
(BENG_CLIMATE_CFG, BENG_BUILDER_CFG, BENG_TRAFFIC_CFG): python profiles/cfg_probe.py <env> <VAR> <value|default>...
Each value is timed per step on its own events with a 256 MB fill in between (L2 flushed) and back-to-back (which is
CPU-enqueue-bound below ~25 us per step), in interleaved rounds so that box drift hits every variant alike."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import custom_gymnasium_environments_b200 as pkg

name, var = sys.argv[1], sys.argv[2]
cfgs = sys.argv[3:]
dev = torch.device("cuda:0")
n = 1 << 20
if name == "climate":
    env = pkg.BatchedSmartClimateEnv(n, device=dev)
    acts = [{"ac_temp": torch.rand(n, device=dev) * 16 + 16, "lights": torch.randint(0, 2, (n, 4), device=dev).to(torch.int8)} for _ in range(16)]
elif name == "builder":
    env = pkg.BatchedWorldBuilderEnv(n, device=dev); acts = torch.randint(0, 5, (16, n), device=dev)
else:
    n = 1 << 16
    env = pkg.BatchedTrafficManagementEnv(n, device=dev); acts = torch.randint(0, 3, (16, n, 9), device=dev)
os.environ.pop(var, None)
env.reset()
for t in range(100):
    env.step(acts[t % 16])
scratch = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {c: [] for c in cfgs}
K = 40
for rnd in range(3):
    for c in cfgs:
        os.environ.pop(var, None)
        if c != "default":
            os.environ[var] = c
        for t in range(5):
            env.step(acts[t % 16])
        pairs = []
        for t in range(K):
            scratch.fill_(t & 0xFF)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); env.step(acts[t % 16]); b.record()
            pairs.append((a, b))
        torch.cuda.synchronize()
        res[c] += [a.elapsed_time(b) * 1e3 for a, b in pairs]
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for t in range(K):
            env.step(acts[t % 16])
        b.record(); torch.cuda.synchronize()
        res[c].append(("warm", a.elapsed_time(b) * 1e3 / K))
for c in cfgs:
    cold = sorted(x for x in res[c] if not isinstance(x, tuple))
    warm = [x[1] for x in res[c] if isinstance(x, tuple)]
    print(f"{name} n={n} {var}={c:10s} flushed mean {sum(cold)/len(cold):7.2f} us  median {cold[len(cold)//2]:7.2f}  min {cold[0]:7.2f}   "
          f"back-to-back {min(warm):7.2f} us", flush=True)
