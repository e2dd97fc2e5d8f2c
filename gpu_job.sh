cd $GRAFT_REPO_ROOT
P='import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({k:d[k] for k in ("value","ms_per_step","roofline","obs_gather","config","e2e","concurrent_streams")}))'
rm -f gpurun_out/bench_other_workloads.jsonl
for wl in tiny-2ag small-4ag medium-8ag-hard large-16ag-r2 tiny-4ag-image tiny-4ag-imagedict; do python bench.py --workload $wl --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 6 2>/dev/null | tail -1 | python -c "$P" >> gpurun_out/bench_other_workloads.jsonl; done
wc -l gpurun_out/bench_other_workloads.jsonl
