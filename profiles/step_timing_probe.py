"""Where does a step's time go?  CPU enqueue time per env.step() vs device time per step (events) vs back-to-back."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import custom_gymnasium_environments_b200 as pkg

name = sys.argv[1] if len(sys.argv) > 1 else "crypto"
dev = torch.device("cuda:0")
if name == "crypto":
    n = 1 << 18; env = pkg.BatchedCryptoTradingEnv(n, None, "discrete", device=dev); acts = torch.randint(0, 5, (64, n), device=dev)
elif name == "snake":
    n = 1 << 20; env = pkg.BatchedSnakeEnv(n, device=dev); acts = torch.randint(0, 4, (64, n), device=dev)
elif name == "climate":
    n = 1 << 20; env = pkg.BatchedSmartClimateEnv(n, device=dev)
    acts = [{"ac_temp": torch.rand(n, device=dev) * 16 + 16, "lights": torch.randint(0, 2, (n, 4), device=dev).to(torch.int8)} for _ in range(64)]
elif name == "builder":
    n = 1 << 20; env = pkg.BatchedWorldBuilderEnv(n, device=dev); acts = torch.randint(0, 5, (64, n), device=dev)
else:
    n = 1 << 16; env = pkg.BatchedTrafficManagementEnv(n, device=dev); acts = torch.randint(0, 3, (64, n, 9), device=dev)
env.reset()
for t in range(50): env.step(acts[t % 64])
torch.cuda.synchronize()
K = 300
# (a) CPU enqueue time only
t0 = time.perf_counter()
for t in range(K): env.step(acts[t % 64])
t_enq = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
# (b) per-step events, sync after each step
evs = []
for t in range(K):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); env.step(acts[t % 64]); b.record(); torch.cuda.synchronize(); evs.append(a.elapsed_time(b))
evs.sort()
# (c) back-to-back events
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for t in range(K): env.step(acts[t % 64])
b.record(); torch.cuda.synchronize()
print(f"{name}: cpu enqueue {1e6*t_enq/K:.1f} us/step; wall incl. drain {1e6*t_all/K:.1f} us/step; "
      f"isolated step median {1e3*evs[K//2]:.1f} us (min {1e3*evs[0]:.1f}, p90 {1e3*evs[int(K*.9)]:.1f}); "
      f"back-to-back {1e3*a.elapsed_time(b)/K:.1f} us/step")
