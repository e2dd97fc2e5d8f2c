cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -2 gpurun_out/bench_r1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r1.json')); e=d['e2e']
print('value %.3e'%d['value'], 'frac', round(d['roofline']['frac'],3), 'e2e %.3e'%e['value'], round(e['ms_per_step'],3), 'plain %.3e'%e['plain_float32_transport']['value'], 'cpu %.3e'%d['cpu_baseline']['value'], d['cpu_baseline'].get('cgroup_cpu_quota'), 'streams2 %.3e'%d['concurrent_streams']['value'])
"
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_under_ncu.log 2>&1; tail -1 gpurun_out/launches_r1.csv | cut -c1-150
