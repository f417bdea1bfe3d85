#!/bin/bash
# Round-1 evidence run (under gpurun): GPU tests, one ncu --set full capture + launch list per step kernel (each
# after the same command exited 0 without ncu), and the bench lines of all five envs.
set -u
out=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > $out/pytest_gpu_r1.log 2>&1; echo "pytest rc=$?"; tail -2 $out/pytest_gpu_r1.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_r1.log 2>&1; echo "smoke rc=$?"; tail -1 $out/smoke_r1.log
for env in snake crypto traffic climate builder; do
  kern=${env}_kernel; [ $env = crypto ] && kern=crypto2_kernel; [ $env = traffic ] && kern=traffic_wpi_kernel
  B="python bench.py --env $env --steps 120 --warmup 40 --no-cpu-baseline --e2e-steps 1 --no-l2-flush"
  $B > $out/plain_$env.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$kern -s 100 -c 1 -o $out/prof_${env}_r1_final $B > $out/ncu_full_$env.log 2>&1
  skip=200; [ $env = climate ] && skip=60   # the climate run builds its tapes with a handful of torch launches
  $B > $out/plain2_$env.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s $skip -c 100 --csv --log-file $out/launches_${env}_r1.csv $B > $out/ncu_list_$env.log 2>&1
done
python bench.py --env snake --steps 2000 --warmup 200 > $out/bench_r1_snake_n1.json 2> $out/bench_snake.err; tail -1 $out/bench_snake.err
python bench.py --env crypto --steps 1200 --warmup 100 > $out/bench_r1_crypto_n1.json 2> $out/bench_crypto.err; tail -1 $out/bench_crypto.err
python bench.py --env traffic --steps 2000 --warmup 200 > $out/bench_r1_traffic_n1.json 2> $out/bench_traffic.err; tail -1 $out/bench_traffic.err
python bench.py --env climate --steps 1500 --warmup 100 > $out/bench_r1_climate_n1.json 2> $out/bench_climate.err; tail -1 $out/bench_climate.err
python bench.py --env builder --steps 1500 --warmup 100 > $out/bench_r1_builder_n1.json 2> $out/bench_builder.err; tail -1 $out/bench_builder.err
python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_r1_snake_reference.json 2>/dev/null
for f in snake crypto traffic climate builder; do python -c "
import json; d=json.load(open('$out/bench_r1_${f}_n1.json')); print('$f', round(d['value']/1e9,3),'G/s', round(d['ms_per_step']*1e3,1),'us frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value']/1e6,1),'M/s cpu',round(d['cpu_baseline']['value']),d['clocks']['sm_mhz'],d['clocks']['reasons'],'launches',d['gpu_launches'])"; done
