#!/bin/bash
# Refresh of the round-1 evidence for the kernels changed after collect_r1.sh ran (usage: collect_r1b.sh [traffic] [climate],
# default both): GPU tests, one ncu --set full capture + launch list each, bench lines.
set -u
out=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > $out/pytest_gpu_r1.log 2>&1; echo "pytest rc=$?"; tail -2 $out/pytest_gpu_r1.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_r1.log 2>&1; echo "smoke rc=$?"; tail -1 $out/smoke_r1.log
ENVS=${*:-traffic climate}
for env in $ENVS; do
  kern=${env}_kernel; [ $env = traffic ] && kern=traffic_wpi_kernel
  B="python bench.py --env $env --steps 120 --warmup 40 --no-cpu-baseline --e2e-steps 1 --no-l2-flush"
  $B > $out/plain_$env.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$kern -s 100 -c 1 -o $out/prof_${env}_r1_final $B > $out/ncu_full_$env.log 2>&1
  skip=200; [ $env = climate ] && skip=60
  $B > $out/plain2_$env.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s $skip -c 100 --csv --log-file $out/launches_${env}_r1.csv $B > $out/ncu_list_$env.log 2>&1
done
python bench.py --env traffic --steps 2000 --warmup 200 > $out/bench_r1_traffic_n1.json 2> $out/bench_traffic.err; tail -1 $out/bench_traffic.err
python bench.py --env traffic --envs-per-gpu 1048576 --steps 400 --warmup 50 --no-cpu-baseline > $out/bench_r1_traffic_n1_1m_envs.json 2> $out/bench_traffic1m.err; tail -1 $out/bench_traffic1m.err
case " $ENVS " in *" climate "*) python bench.py --env climate --steps 1500 --warmup 100 > $out/bench_r1_climate_n1.json 2> $out/bench_climate.err; tail -1 $out/bench_climate.err;; esac
for f in traffic traffic_n1_1m_envs $(case " $ENVS " in *" climate "*) echo climate;; esac); do g=bench_r1_${f}_n1.json; [ $f = traffic_n1_1m_envs ] && g=bench_r1_traffic_n1_1m_envs.json; python -c "
import json; d=json.load(open('$out/$g')); print('$f', round(d['value']/1e9,3),'G/s', round(d['ms_per_step']*1e3,1),'us frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value']/1e6,1),'M/s', d['clocks']['sm_mhz'],d['clocks']['reasons'],'launches',d['gpu_launches'], 'warm', d.get('value_l2_warm'))"; done
