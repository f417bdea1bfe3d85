#!/usr/bin/env python
"""Copy what profiles/collect_r1.sh (all envs) or collect_r1b.sh (traffic, climate) left in gpurun_out/ into profiles/
(tracked) and regenerate the derived files: ncu summaries, launch shares of the refreshed envs, and the
<env>_step_traffic.json files bench.py reads.

    python profiles/refresh_r1.py [env ...]        # default: traffic climate
"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
SIZES = {"snake": 1 << 20, "crypto": 1 << 18, "traffic": 1 << 16, "climate": 1 << 20, "builder": 1 << 20}
ENVS = {e: SIZES[e] for e in (sys.argv[1:] or ["traffic", "climate"])}

for name in [f"bench_r1_{e}_n1.json" for e in ENVS] + ["bench_r1_traffic_n1_1m_envs.json", "bench_r1_snake_reference.json",
                                                         "pytest_gpu_r1.log", "smoke_r1.log"]:
    if os.path.exists(os.path.join(OUT, name)) and os.path.getmtime(os.path.join(OUT, name)) > time.time() - 6 * 3600:
        shutil.copy(os.path.join(OUT, name), os.path.join(PROF, name))
for env in ENVS:
    shutil.copy(os.path.join(OUT, f"launches_{env}_r1.csv"), os.path.join(PROF, f"{env}_r1_launches.csv"))
    with open(os.path.join(PROF, f"{env}_step_r1_ncu_summary.txt"), "w") as f:
        subprocess.run([sys.executable, os.path.join(PROF, "ncu_summary.py"),
                        os.path.join(OUT, f"prof_{env}_r1_final.ncu-rep")], stdout=f, check=True)


def launch_block(env):
    rows = [r for r in csv.reader(open(os.path.join(PROF, f"{env}_r1_launches.csv"))) if len(r) > 10 and r[0].isdigit()]
    d = collections.defaultdict(list)
    for r in rows:
        d[re.sub(r"\(.*", "", r[4])].append(float(r[-1]) / 1e3)
    tot, n = sum(sum(v) for v in d.values()), sum(len(v) for v in d.values())
    s = f"{env}: {n} launches captured inside the timed region (cold-cache, serialised)\n"
    for k, v in sorted(d.items(), key=lambda x: -sum(x[1])):
        s += f"   {100 * sum(v) / tot:5.1f}%  n={len(v):4d}  mean {sum(v) / len(v):8.1f} us   {k}\n"
    return s


path = os.path.join(PROF, "r1_launch_shares.txt")
blocks = [b for b in open(path).read().split("== ") if b.strip()]
blocks = [launch_block(b.split(":")[0]) if b.split(":")[0] in ENVS else b for b in blocks]
open(path, "w").write("".join("== " + b for b in blocks))

for env, n in ENVS.items():
    t = open(os.path.join(PROF, f"{env}_step_r1_ncu_summary.txt")).read()
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def metric(name):
        m = re.search(name + r"\s+([\d.]+) (\w+)", t)
        return float(m.group(1)) * unit[m.group(2)]

    rd, wr = metric("dram__bytes_read.sum"), metric("dram__bytes_write.sum")
    us = float(re.search(r"gpu__time_duration.sum\s+([\d.]+) us", t).group(1))
    p = os.path.join(PROF, f"{env}_step_traffic.json")
    j = json.load(open(p))
    j.update(dram_bytes_per_env_step=(rd + wr) / n, dram_bytes_read_per_launch=rd, dram_bytes_write_per_launch=wr,
             n_envs=n, kernel_us_under_ncu=us)
    json.dump(j, open(p, "w"), indent=1)
    print(env, round((rd + wr) / n, 1), "B/env-step DRAM,", us, "us under ncu")
print(open(path).read())
