#!/usr/bin/env python
"""Mixed rollout (BASELINE.json configs[4]): batched SnakeEnv on the CUDA engine driving a small torch PPO policy,
one process per GPU, observations never leaving HBM.  Reports the env-step / policy-forward / PPO-update time split.

    python examples/ppo_snake.py --envs 262144 --horizon 16 --iters 3
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 examples/ppo_snake.py --envs 1048576

This replaces the RLlib env-runner pattern of the reference's training scripts
(smartclimate_rl-main/training/train.py:10-15, smart_parking_env/examples/training.py:28-48: CPU actors stepping 24
envs each) and its tabular Q-learner (snake_env_classic/train.py:6-111): the env writes each step's observation
straight into the (T, N, 20, 20) rollout buffer (`step(..., out_obs=buf[t])`), the policy reads it in place and its
sampled int64 actions are fed back as-is.  The only collectives are DDP's gradient all-reduce and one all-reduce of
the integer episode statistics per iteration.  A consumer of the hot path, not part of it.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import custom_gymnasium_environments_b200 as pkg  # noqa: E402
from custom_gymnasium_environments_b200.dist import all_reduce_episode_stats, init_process_group, summarize  # noqa: E402


class SnakePolicy(nn.Module):
    """Small CNN actor-critic over the int8 (20, 20) grid, one-hot encoded on the fly (empty / snake / food)."""

    def __init__(self, grid=20, hidden=128):
        super().__init__()
        self.c1 = nn.Conv2d(3, 16, 3, padding=1)
        self.c2 = nn.Conv2d(16, 32, 3, stride=2, padding=1)
        self.fc = nn.Linear(32 * (grid // 2) ** 2, hidden)
        self.pi = nn.Linear(hidden, 4)
        self.v = nn.Linear(hidden, 1)

    def forward(self, obs_i8):
        # one-hot straight from the int8 grid (no int64 copy of the observation): channel c = (cell == c)
        classes = torch.arange(3, dtype=torch.int8, device=obs_i8.device).view(1, 3, 1, 1)
        x = (obs_i8.unsqueeze(1) == classes).to(torch.bfloat16)
        x = F.relu(self.c1(x))
        x = F.relu(self.c2(x))
        x = F.relu(self.fc(x.flatten(1)))
        return self.pi(x).float(), self.v(x).float().squeeze(-1)


def timed(stream):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    return a, b


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 18, help="envs per GPU")
    ap.add_argument("--horizon", type=int, default=16)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--minibatch", type=int, default=1 << 16)
    ap.add_argument("--fwd-chunk", type=int, default=1 << 17, help="envs per policy-forward chunk (bounds activations)")
    ap.add_argument("--lr", type=float, default=3e-4)
    ap.add_argument("--out", default=None, help="also write the JSON line to this file (rank 0)")
    return ap.parse_args(argv)


def run(args, record=None):
    """One PPO run.  -> the result dict (rank 0; None elsewhere).  `record`: a list that receives, per iteration, clones
    of (actions (T, n), observations (T + 1, n, G, G)) -- what tests/test_ppo_example_gpu.py replays through a second env."""
    rank, local_rank, world = init_process_group()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    torch.manual_seed(1234 + rank)
    n, T, G = args.envs, args.horizon, 20
    env = pkg.BatchedSnakeEnv(n, G, device=dev, seed=0, env_id_base=rank * n)
    policy = SnakePolicy(G).to(dev).to(torch.bfloat16)
    model = nn.parallel.DistributedDataParallel(policy, device_ids=[local_rank]) if world > 1 else policy
    opt = torch.optim.Adam(model.parameters(), lr=args.lr)
    stream = torch.cuda.current_stream(dev)

    obs_buf = torch.zeros((T + 1, n, G, G), dtype=torch.int8, device=dev)   # the env writes into this directly
    act_buf = torch.zeros((T, n), dtype=torch.int64, device=dev)
    logp_buf = torch.zeros((T, n), device=dev)
    val_buf = torch.zeros((T + 1, n), device=dev)
    rew_buf = torch.zeros((T, n), device=dev)
    done_buf = torch.zeros((T, n), device=dev)

    def act(obs):
        logits, values = [], []
        for s in range(0, n, args.fwd_chunk):
            lg, v = policy(obs[s:s + args.fwd_chunk])
            logits.append(lg)
            values.append(v)
        return torch.cat(logits), torch.cat(values)

    obs, _ = env.reset()
    obs_buf[0].copy_(obs)
    split = {"env_step_ms": 0.0, "policy_forward_ms": 0.0, "update_ms": 0.0}
    t_wall = time.perf_counter()
    for it in range(args.iters):
        ev_env, ev_fwd = [], []
        with torch.no_grad():
            for t in range(T):
                a0, b0 = timed(stream)
                logits, val = act(obs_buf[t])
                dist_ = torch.distributions.Categorical(logits=logits)
                action = dist_.sample()
                b0.record(stream)
                ev_fwd.append((a0, b0))
                act_buf[t], logp_buf[t], val_buf[t] = action, dist_.log_prob(action), val
                a1, b1 = timed(stream)
                _, rew, term, _, _ = env.step(action, out_obs=obs_buf[t + 1])   # obs lands in the rollout buffer
                b1.record(stream)
                ev_env.append((a1, b1))
                rew_buf[t].copy_(rew)
                done_buf[t].copy_(term)
            _, val_buf[T] = act(obs_buf[T])
            # GAE(0.99, 0.95)
            adv = torch.zeros_like(rew_buf)
            last = torch.zeros(n, device=dev)
            for t in reversed(range(T)):
                nonterminal = 1.0 - done_buf[t]
                delta = rew_buf[t] + 0.99 * val_buf[t + 1] * nonterminal - val_buf[t]
                last = delta + 0.99 * 0.95 * nonterminal * last
                adv[t] = last
            ret = adv + val_buf[:T]
        a2, b2 = timed(stream)
        flat_obs = obs_buf[:T].reshape(T * n, G, G)
        flat = [x.reshape(T * n) for x in (act_buf, logp_buf, adv, ret)]
        perm = torch.randperm(T * n, device=dev)
        for s in range(0, T * n, args.minibatch):
            idx = perm[s:s + args.minibatch]
            lg, v = model(flat_obs[idx])
            d = torch.distributions.Categorical(logits=lg)
            ratio = (d.log_prob(flat[0][idx]) - flat[1][idx]).exp()
            a_mb = flat[2][idx]
            a_mb = (a_mb - a_mb.mean()) / (a_mb.std() + 1e-8)
            loss = -torch.min(ratio * a_mb, ratio.clamp(0.8, 1.2) * a_mb).mean() \
                + 0.5 * F.mse_loss(v, flat[3][idx]) - 0.01 * d.entropy().mean()
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
        b2.record(stream)
        torch.cuda.synchronize(dev)
        if record is not None:
            record.append((act_buf.clone(), obs_buf.clone()))
        obs_buf[0].copy_(obs_buf[T])
        if it > 0 or args.iters == 1:  # iteration 0 is warm-up (cudnn autotune, allocator)
            split["env_step_ms"] += sum(a.elapsed_time(b) for a, b in ev_env)
            split["policy_forward_ms"] += sum(a.elapsed_time(b) for a, b in ev_fwd)
            split["update_ms"] += a2.elapsed_time(b2)
    wall = time.perf_counter() - t_wall
    stats = summarize(all_reduce_episode_stats(env.stats))
    out = None
    if rank == 0:
        timed_iters = max(1, args.iters - 1) if args.iters > 1 else 1
        tot = sum(split.values())
        out = {"config": "mixed rollout: batched SnakeEnv + small torch PPO (CNN actor-critic, bf16)", "n_gpus": world,
               "envs_per_gpu": n, "horizon": T, "timed_iterations": timed_iters,
               "per_iteration_ms": {k: v / timed_iters for k, v in split.items()},
               "share": {k: v / tot for k, v in split.items()},
               "env_steps_per_s_end_to_end": world * n * T * args.iters / wall,
               "env_steps_per_s_env_only": world * n * T * timed_iters / (split["env_step_ms"] * 1e-3),
               "episodes": stats, "loss": float(loss.detach())}
    return out


def main(argv=None):
    args = parse_args(argv)
    out = run(args)
    if out is not None:
        line = json.dumps(out)
        print(line, flush=True)
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            with open(args.out, "w") as f:
                f.write(line + "\n")
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
