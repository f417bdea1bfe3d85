#!/bin/bash
# Round-2 final evidence after the traffic / builder / climate kernel rework (run under gpurun from the repo root):
# full GPU test suite, smoke(), the default bench command and the per-env primary lines, the ncu launch list of the
# default command, and one `ncu --set full` capture of the builder and climate step kernels.
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r2b.log 2>&1; echo "pytest rc=$?"; tail -1 $O/pytest_gpu_r2b.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_r2b.log 2>&1; echo "smoke rc=$?"
python bench.py > $O/bench_r2b_snake_n1.json 2> $O/bench_r2b_snake_n1.err; echo "snake rc=$?"
python bench.py --env traffic --steps 1000 --warmup 200 --no-secondary > $O/bench_r2b_traffic_n1.json 2> $O/bench_r2b_traffic_n1.err; echo "traffic rc=$?"
python bench.py --env traffic --steps 300 --warmup 100 --envs-per-gpu 1048576 --no-secondary --no-cpu-baseline > $O/bench_r2b_traffic_n1_1m_envs.json 2> $O/bench_r2b_traffic_1m.err; echo "traffic 1M rc=$?"
python bench.py --env builder --steps 1000 --warmup 200 --no-secondary --no-cpu-baseline > $O/bench_r2b_builder_n1.json 2> $O/bench_r2b_builder_n1.err; echo "builder rc=$?"
python bench.py --env climate --steps 1000 --warmup 200 --no-secondary --no-cpu-baseline > $O/bench_r2b_climate_n1.json 2> $O/bench_r2b_climate_n1.err; echo "climate rc=$?"
python - <<'PY'
import json
for f in ('snake_n1','traffic_n1','traffic_n1_1m_envs','builder_n1','climate_n1'):
    try:
        d=json.load(open(f'gpurun_out/bench_r2b_{f}.json'))
    except Exception as e:
        print(f, 'FAILED', e); continue
    print(f, round(d['value']/1e9,3), 'G', round(d['ms_per_step']*1e3,1), 'us frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value']/1e6,1), 'M', d['clocks'])
    for k,v in d.get('secondary',{}).items(): print('   secondary', k, round(v['value']/1e9,3), round(v['ms_per_step']*1e3,1), round(v['roofline']['frac'],3))
PY
CMD="python bench.py --steps 40 --warmup 5 --e2e-steps 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/launches_r2b.csv $CMD > $O/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:climate_step -s 20 -c 1 -o $O/prof_climate_r2b python profiles/prof_step.py climate 30 > $O/ncu.log 2>&1; echo rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:builder_kernel -s 20 -c 1 -o $O/prof_builder_r2b python profiles/prof_step.py builder 30 > $O/ncu2.log 2>&1; echo rc=$?
