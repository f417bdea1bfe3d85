#!/bin/bash
# Config sweep of the snake step kernel: BENG_SNAKE_CFG="tile,stages,ctas_per_sm".  Run under gpurun.
out=gpurun_out/sweep.txt
: > $out
for cfg in "$@"; do
  BENG_SNAKE_CFG=$cfg timeout 120 python bench.py --steps 300 --warmup 30 --no-cpu-baseline --e2e-steps 1 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$cfg', round(d['value']/1e9,3), 'Gsteps/s', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])" >> $out
done
cat $out
