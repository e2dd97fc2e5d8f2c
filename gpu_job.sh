cd $GRAFT_REPO_ROOT
python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -2 gpurun_out/bench_r1.err; cut -c1-300 gpurun_out/bench_r1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rw_main_kernel --launch-skip 6 --launch-count 2 -f -o gpurun_out/prof_r1_final2 python profiles/prof_step.py > gpurun_out/ncu_final2.log 2>&1; tail -2 gpurun_out/ncu_final2.log
python profiles/phase_timeline.py > gpurun_out/phase_timeline_tiny4ag.json 2>/dev/null
P='import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({k:d[k] for k in ("value","ms_per_step","roofline","obs_gather","config","e2e")}))'
for wl in tiny-2ag small-4ag medium-8ag-hard large-16ag-r2 tiny-4ag-image; do python bench.py --workload $wl --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 4 2>/dev/null | tail -1 | python -c "$P" >> gpurun_out/bench_other_workloads.jsonl; done
cat gpurun_out/bench_other_workloads.jsonl | cut -c1-200
