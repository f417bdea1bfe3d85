#!/bin/bash
# Round-2 evidence collected on the GPU box (run under gpurun from the repo root):
#   1. the default bench command (snake headline + crypto/traffic secondary blocks), plain
#   2. the ncu launch list of the SAME command (cold-cache, serialised: compare shares, not absolutes)
cd ${GRAFT_REPO_ROOT:-.}
CMD="python bench.py --steps 40 --warmup 5 --e2e-steps 3 --no-cpu-baseline"
$CMD > gpurun_out/bench_r2_short.json 2> gpurun_out/bench_r2_short.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_launches.log; wc -l gpurun_out/launches_r2.csv
