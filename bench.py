#!/usr/bin/env python
"""Benchmark of the batched env-step hot path (BASELINE.json: env-steps/sec, batched SnakeEnv, 1M envs/GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA engine
    python bench.py --impl reference [--gpus N] --steps K --warmup W   # CPU arm: the reference's step loop

A "step" is ONE launch of the fused step kernel over the whole batch (1,048,576 envs per GPU), inputs
resident in HBM; `value` = envs x steps x ranks / max-over-ranks device time (CUDA events).  `e2e` is the
same metric through the host-buffer C-ABI call (`beng_snake_step_host`: pinned host actions in, numpy
obs/reward/terminated out, copies inside the timed region).  `roofline` uses SURVEY.md 8(d)'s 450
algorithmic bytes per env-step against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
`cpu_baseline` times the pure-Python port of the reference's per-env step loop (oracle/snake_port.py; the
reference itself cannot travel to the GPU box) on the host cores, on rank 0 at N=1 only.

Prints exactly one JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"
SNAKE_BYTES_PER_ENV_STEP = 450  # obs 400 W + action 8 R + reward 4 W + flags 2 W + state 16 R/W + ring 2 R/W
FALLBACK_HBM_GBS = 6650.0       # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's per-env Python step loop (port), one env per process
# ------------------------------------------------------------------------------------------------
def _cpu_worker_loop(args):
    """Step one SnakePort env with random actions and reset-on-done for `n_steps` steps (the loop shape of
    crypto_trading_env/test_crypto_trading.py:410-416 applied to snake).  Returns (steps, seconds)."""
    worker, n_steps = args
    import random

    from oracle.snake_port import SnakePort

    rng = random.Random(1234 + worker)
    env = SnakePort(20, rng=rng)
    env.reset()
    randrange = rng.randrange
    t0 = time.perf_counter()
    for _ in range(n_steps):
        _, _, term, trunc, _ = env.step(randrange(4))
        if term or trunc:
            env.reset()
    return n_steps, time.perf_counter() - t0


def cpu_rate_single(seconds: float = 2.0) -> float:
    n, dt = _cpu_worker_loop((0, 20000))
    rate = n / dt
    n, dt = _cpu_worker_loop((0, max(20000, int(rate * seconds))))
    return n / dt


def cpu_rate_parallel(pool, procs: int, steps_per_proc: int):
    t0 = time.perf_counter()
    res = pool.map(_cpu_worker_loop, [(w, steps_per_proc) for w in range(procs)])
    wall = time.perf_counter() - t0
    return sum(r[0] for r in res) / wall, wall


def c_oracle_rate(seconds: float = 1.5) -> float:
    """Throughput of the plain-C oracle (1 thread), for context next to the Python port."""
    import numpy as np

    from oracle.c_oracle import SnakeOracle

    n = 1 << 15
    orc = SnakeOracle(n, seed=0)
    orc.reset()
    rng = np.random.default_rng(0)
    acts = rng.integers(0, 4, (16, n))
    orc.step(acts[0])
    t0 = time.perf_counter()
    k = 0
    while time.perf_counter() - t0 < seconds:
        orc.step(acts[k % 16])
        k += 1
    return n * k / (time.perf_counter() - t0)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    cores = os.cpu_count() or 1
    single = cpu_rate_single(1.0)
    budget_s = 60.0
    per_step = int(min(50000, max(200, single * budget_s / max(1, args.steps + args.warmup))))
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(args.warmup):
            cpu_rate_parallel(pool, cores, per_step)
        t0 = time.perf_counter()
        total = 0
        for _ in range(args.steps):
            res = pool.map(_cpu_worker_loop, [(w, per_step) for w in range(cores)])
            total += sum(r[0] for r in res)
        wall = time.perf_counter() - t0
    value = total / wall
    sample = (f"pure-Python port of SnakeEnvClassic (oracle/snake_port.py), G=20, 1 env per process x {cores} "
              f"processes, random actions, reset on done; each step = {per_step} env-steps per process")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, per_gpu=args.envs_per_gpu),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "single_core_value": single},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
    NOTED = {"sw_power_cap": 0x4, "hw_power_brake": 0x80, "sync_boost": 0x10}

    def __init__(self, index: int, period_s: float = 0.02):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self.period = period_s
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, repr(e)

    def _once(self):
        nv = self.nv
        self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        try:
            bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        except Exception:
            bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
        for name, bit in {**self.BAD, **self.NOTED}.items():
            if bits & bit:
                self.reasons.add(name)

    def _run(self):
        while not self._stop.is_set():
            try:
                self._once()
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        if self.nv:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread:
            self._stop.set()
            self._thread.join()

    def summary(self):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml_unavailable"]}
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
def workload_config(args, per_gpu):
    return {"workload": "batched SnakeEnvClassic, G=20, random actions, SAME_STEP auto-reset "
                        "(BASELINE.json configs[1])",
            "envs_per_gpu": per_gpu, "global_envs": per_gpu * args.gpus, "grid_size": 20, "max_steps": 1000,
            "actions": "i.i.d. uniform{0..3} int64, device-generated tape (Philox stream 1), resident in HBM",
            "l2_policy": "inputs larger than L2: each step streams 472 MB (400 MB obs write) through a 126 MB L2; "
                         "action tape cycles over >= 64 distinct 8 MB tensors",
            "parallelism": f"env-sharded x{args.gpus}, no data-path collective"}


def load_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic(n_envs):
    """Per-launch DRAM bytes of the step kernel from the committed ncu capture (profiles/), scaled to this batch."""
    try:
        with open(os.path.join(ROOT, "profiles", "snake_step_traffic.json")) as f:
            d = json.load(f)
        return d["dram_bytes_per_env_step"] * n_envs
    except Exception:
        return None


def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    import custom_gymnasium_environments_b200 as pkg
    from custom_gymnasium_environments_b200.dist import all_reduce_episode_stats, init_process_group, summarize

    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO") and not os.environ.get("BENG_KEEP_NCCL_DEBUG"):
        os.environ["NCCL_DEBUG"] = "WARN"  # NCCL's banner goes to stdout and would precede the JSON line
    rank, local_rank, world = init_process_group()
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = pkg._lib.load()
    n = args.envs_per_gpu
    seed = 0
    base = rank * n  # contiguous global env-id slice per rank (weak scaling)

    env = pkg.BatchedSnakeEnv(n, 20, device=dev, seed=seed, env_id_base=base)
    env.reset()
    stream = torch.cuda.current_stream(dev)

    # action tapes, resident in HBM before the timed region (pool of distinct steps, cycled)
    pool = max(64, min(args.steps + args.warmup, args.action_pool))
    tapes = torch.empty((pool, n), dtype=torch.int64, device=dev)
    for t in range(pool):
        pkg._lib.check(lib.beng_fill_random_actions(tapes[t].data_ptr(), n, 1, 4, t, base, seed, stream.cuda_stream),
                       "beng_fill_random_actions")
    torch.cuda.synchronize(dev)

    def barrier():
        if world > 1:
            dist.barrier()

    # ---- device-resident timing -----------------------------------------------------------------
    for t in range(args.warmup):
        env.step(tapes[t % pool])
    torch.cuda.synchronize(dev)
    barrier()
    sampler = ClockSampler(local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = lib.beng_launch_count()
    torch.cuda.synchronize(dev)
    sampler.start()
    ev0.record(stream)
    for t in range(args.steps):
        env.step(tapes[(args.warmup + t) % pool])
    ev1.record(stream)
    torch.cuda.synchronize(dev)
    sampler.stop()
    launches = lib.beng_launch_count() - launches0
    barrier()
    ms = ev0.elapsed_time(ev1)
    ms_t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_max = float(ms_t.item())
    value = n * world * args.steps / (ms_max * 1e-3)

    # ---- end-to-end through the host-buffer C-ABI call -------------------------------------------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    host_tapes = [tapes[t % pool].cpu().pin_memory() for t in range(min(e2e_steps + 2, 8))]
    G = 20
    h2d = n * 8
    d2h_full = n * (G * G + 4 + 1 + 1 + 4 + 4)
    d2h_lite = n * (4 + 1 + 1 + 4 + 4)

    def time_e2e(copy_obs):
        for t in range(2):
            env.step_host(host_tapes[t % len(host_tapes)], copy_obs=copy_obs)
        torch.cuda.synchronize(dev)
        barrier()
        t0 = time.perf_counter()
        for t in range(e2e_steps):
            # synchronous, as a caller of the reference's numpy API would see it: results are on the host on return
            env.step_host(host_tapes[t % len(host_tapes)], copy_obs=copy_obs)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        dt_t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt_t, op=dist.ReduceOp.MAX)
        return n * world * e2e_steps / float(dt_t.item())

    e2e_full = time_e2e(True)
    e2e_lite = time_e2e(False)

    # ---- episode statistics: the only collective, outside the step loop ---------------------------
    stats = all_reduce_episode_stats(env.stats)
    summary = summarize(stats)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peak, peak_src = load_peak()
    import ctypes as C
    t_, s_, c_ = C.c_int32(), C.c_int32(), C.c_int32()
    lib.beng_snake_launch_config(20, n, C.byref(t_), C.byref(s_), C.byref(c_))
    kernel_name = (f"beng::snake_kernel<T={t_.value},STAGES={s_.value},IS_RESET=false,OWNROW=true> "
                   f"grid={c_.value} CTAs/SM persistent")
    per_launch_ms = ms / max(1, args.steps)  # rank-0 kernel time; the timed region is back-to-back step launches
    achieved = n * SNAKE_BYTES_PER_ENV_STEP / (per_launch_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, per_gpu=n),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": load_traffic(n), "peak_source": peak_src,
                     "algorithmic_bytes_per_env_step": SNAKE_BYTES_PER_ENV_STEP,
                     "kernel": kernel_name, "kernel_ms": per_launch_ms},
        "e2e": {"value": e2e_full, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_full,
                "steps": e2e_steps, "api": "BatchedSnakeEnv.step_host -> beng_snake_step_host (pinned host buffers, "
                                           "synchronous per step, full observation copied back)"},
        "e2e_obs_on_device": {"value": e2e_lite, "unit": UNIT, "h2d_bytes_per_step": h2d,
                              "d2h_bytes_per_step": d2h_lite,
                              "note": "same call with obs_host=NULL: reward/terminated/info to host, observation "
                                      "stays in HBM for an on-device policy"},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
        "episodes": summary,
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        single = cpu_rate_single(2.0)
        ctx = mp.get_context("fork")
        with ctx.Pool(cores) as pool_:
            per_proc = int(single * 10.0)
            par, wall = cpu_rate_parallel(pool_, cores, per_proc)
        line["cpu_baseline"] = {
            "value": par, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"pure-Python port of the reference step loop (oracle/snake_port.py): 1 env per process x "
                      f"{cores} processes x {per_proc} random-action steps with reset-on-done ({wall:.1f} s wall)",
            "single_core_value": single, "c_oracle_single_thread_value": c_oracle_rate()}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--envs-per-gpu", type=int, default=1 << 20)
    ap.add_argument("--action-pool", type=int, default=256, help="distinct pre-generated action steps kept in HBM")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
