#!/bin/bash
# Multi-GPU evidence (SURVEY section 8 d/e): weak scaling of the three north-star envs on N GPUs of one box.
# usage (under gpurun --gpus N): bash profiles/collect_r1_multi.sh N
set -u
N=$1; out=gpurun_out; port=29611
for env in snake crypto traffic; do
  steps=2000; [ $env = crypto ] && steps=1000
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $N --env $env --steps $steps --warmup 200 --no-cpu-baseline > $out/bench_r1_${env}_n$N.json 2> $out/bench_${env}_n$N.err
  echo "$env rc=$?"; port=$((port+1))
  python -c "
import json; d=json.loads(open('$out/bench_r1_${env}_n$N.json').read().strip().splitlines()[-1]); print('$env n=$N', round(d['value']/1e9,3),'G/s', round(d['ms_per_step']*1e3,1),'us frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value']/1e6,1),'M/s', d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
