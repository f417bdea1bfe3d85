"""Same-box A/B of traffic step-kernel shapes: BENG_TRAFFIC_CFG values given on the command line ("ipw,maxreg";
"default" = the library's default shape), each timed per step on its own events with a 256 MB fill in between (L2 flushed), at
65,536 envs and, with --big, at 1,048,576 envs.  Interleaved rounds so box drift hits every config alike."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import custom_gymnasium_environments_b200 as pkg

cfgs = [c for c in sys.argv[1:] if not c.startswith("--")]
big = "--big" in sys.argv
dev = torch.device("cuda:0")
for n in ([1 << 16, 1 << 20] if big else [1 << 16]):
    env = pkg.BatchedTrafficManagementEnv(n, device=dev)
    acts = torch.randint(0, 3, (32, n, 9), device=dev)
    env.reset()
    os.environ.pop("BENG_TRAFFIC_CFG", None)
    for t in range(150):
        env.step(acts[t % 32])
    scratch = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = {c: [] for c in cfgs}
    K = 60 if n <= (1 << 16) else 20
    for rnd in range(3):
        for c in cfgs:
            os.environ.pop("BENG_TRAFFIC_CFG", None)
            if c != "default":
                os.environ["BENG_TRAFFIC_CFG"] = c
            for t in range(5):
                env.step(acts[t % 32])
            pairs = []
            for t in range(K):
                scratch.fill_(t & 0xFF)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); env.step(acts[t % 32]); b.record()
                pairs.append((a, b))
            torch.cuda.synchronize()
            res[c] += [a.elapsed_time(b) * 1e3 for a, b in pairs]
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for t in range(K):
                env.step(acts[t % 32])
            b.record(); torch.cuda.synchronize()
            res[c].append(("warm", a.elapsed_time(b) * 1e3 / K))
    for c in cfgs:
        cold = sorted(x for x in res[c] if not isinstance(x, tuple))
        warm = [x[1] for x in res[c] if isinstance(x, tuple)]
        print(f"n={n} cfg={c:12s} flushed mean {sum(cold)/len(cold):7.2f} us  median {cold[len(cold)//2]:7.2f}  min {cold[0]:7.2f}   "
              f"back-to-back {min(warm):7.2f} us", flush=True)
    del env, acts, scratch
