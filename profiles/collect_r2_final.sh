#!/bin/bash
# Final round-2 bench lines on one B200 (run under gpurun from the repo root): the default command (snake headline +
# secondary crypto/traffic blocks + reference-class CPU arm), the crypto and traffic workloads as primary lines with
# their own CPU arms, and the stand-alone reference arm.
cd ${GRAFT_REPO_ROOT:-.}
python bench.py > gpurun_out/bench_r2_snake_n1.json 2> gpurun_out/bench_r2_snake_n1.err; echo "snake rc=$?"
python bench.py --env crypto --steps 1000 --warmup 200 > gpurun_out/bench_r2_crypto_n1.json 2> gpurun_out/bench_r2_crypto_n1.err; echo "crypto rc=$?"
python bench.py --env traffic --steps 1000 --warmup 200 > gpurun_out/bench_r2_traffic_n1.json 2> gpurun_out/bench_r2_traffic_n1.err; echo "traffic rc=$?"
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_r2_snake_reference.json 2> /dev/null; echo "ref rc=$?"
python - <<'PY'
import json
for e in ('snake','crypto','traffic'):
    d=json.load(open(f'gpurun_out/bench_r2_{e}_n1.json'))
    c=d.get('cpu_baseline',{})
    print(e, round(d['value']/1e9,3), 'G', round(d['ms_per_step']*1e3,1), 'us frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value']/1e6,1), 'M pcie_frac', round(d['e2e']['pcie_frac'],3), '| cpu', c.get('kind'), c.get('cores'), round(c.get('value',0)), 'single', round(c.get('single_core_value',0)))
    for k,v in d.get('secondary',{}).items(): print('   secondary', k, round(v['value']/1e9,3), round(v['ms_per_step']*1e3,1), round(v['roofline']['frac'],3))
PY
