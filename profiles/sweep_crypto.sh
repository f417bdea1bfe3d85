#!/bin/bash
# Tile sweep of the crypto step kernel: BENG_CRYPTO_TILE = envs (threads) per CTA.  Run under gpurun.
out=gpurun_out/sweep_crypto.txt
: > $out
for t in "$@"; do
  BENG_CRYPTO_TILE=$t timeout 120 python bench.py --env crypto --steps 300 --warmup 30 --no-cpu-baseline --e2e-steps 1 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tile=$t', round(d['value']/1e9,3), 'Gsteps/s', round(d['roofline']['kernel_ms']*1e3,1), 'us', round(d['roofline']['frac'],3))" >> $out
done
cat $out
