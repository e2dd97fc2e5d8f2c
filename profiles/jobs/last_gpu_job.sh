cd $GRAFT_REPO_ROOT
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], "%.3e"%d["value"], round(d["ms_per_step"]*1000,2), "us obs", round(d["obs_gather"]["kernel_us"],2))'
python bench.py --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 2 2>/dev/null | tail -1 | python -c "$P" flat
python bench.py --workload tiny-4ag-image --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 2 2>/dev/null | tail -1 | python -c "$P" image
